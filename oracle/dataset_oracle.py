"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the dataset's per-step pixel sampling.

  getitem(...)          restates SceneDataset.__getitem__ (code/datasets/scene_hawp_dataset.py:148-194) in numpy: the
                        uv grid (:149-151), the masked-pixel list (:173), the subset `sampling_idx[perm[:R]]` (:176) and
                        the gathers (:179-186).  The permutation is an INPUT (the reference draws it with
                        torch.randperm on the CPU generator).
  pixel_permutation()   numpy mirror of the keyed bijection neat_b200 draws on the device when no permutation is given
                        (neat_b200/csrc/pixels.cuh) -- not reference behaviour, it pins the host/device implementations
                        against each other.

Parity status: PINNED -- oracle/make_golden_dataset.py runs the UNMODIFIED reference __getitem__ (on an instance filled
from tests/golden/hawp_abc.npz, no image folders needed) and stores its outputs in tests/golden/dataset_abc.npz;
tests/test_dataset_oracle.py checks getitem() against them.  Only tests/ may import this module."""
import numpy as np


def uv_grid(img_res):
    """[HW,2] float32: (column, row) of every pixel, row-major (scene_hawp_dataset.py:149-151)."""
    H, W = int(img_res[0]), int(img_res[1])
    pix = np.arange(H * W)
    return np.stack([pix % W, pix // W], axis=1).astype(np.float32)


def getitem(img_res, rgb_image, lines, mask, labels, att_points, perm=None, num_pixels=None):
    """perm: positions into mask.nonzero() (the reference's torch.randperm(n)); the first num_pixels are used.
    perm None -> the full-image branch (sampling_idx is None)."""
    uv = uv_grid(img_res)
    labels = np.asarray(labels).astype(np.int64)
    sample = {"uv": uv, "uv_proj": np.asarray(att_points, dtype=np.float32), "mask": np.asarray(mask).astype(bool),
              "labels": labels, "lines": np.asarray(lines, dtype=np.float32)[labels]}
    gt = {"rgb": np.asarray(rgb_image, dtype=np.float32)}
    if perm is None:
        return sample, gt
    idx = np.nonzero(np.asarray(mask).reshape(-1))[0][np.asarray(perm)[:num_pixels]]
    gt["rgb"] = gt["rgb"][idx]
    gt["lines2d"] = np.asarray(lines, dtype=np.float32)[labels[idx]]
    sample["lines"] = gt["lines2d"]
    sample["labels"] = labels[idx]
    sample["uv"] = uv[idx]
    sample["uv_proj"] = sample["uv_proj"][idx]
    sample["sampling_idx"] = idx
    return sample, gt


# ------------------------------------------------------------------------------------------------ device draw mirror
_M32 = np.uint64(0xFFFFFFFF)


def _hash32(x):
    x = np.asarray(x, dtype=np.uint64) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


def pixel_permutation(n, seed, step, first, count):
    """positions drawn for rays first .. first+count-1 (pixels.cuh: make_pixel_perm + px_permute)."""
    bits = 2
    while bits < 32 and (1 << bits) < n:
        bits += 1
    bits += bits & 1
    half = np.uint64(bits // 2)
    m = np.uint64((1 << int(half)) - 1)
    s0 = _hash32((seed & 0xFFFFFFFF) ^ int(_hash32(((seed >> 32) + 0x9E3779B9) & 0xFFFFFFFF)))
    s1 = _hash32((step & 0xFFFFFFFF) ^ int(_hash32(((step >> 32) + 0x85EBCA6B) & 0xFFFFFFFF)))
    keys = [_hash32(int(s0) ^ int(_hash32((int(s1) + k * 0x9E3779B9) & 0xFFFFFFFF))) for k in range(4)]
    x = np.arange(first, first + count, dtype=np.uint64)
    todo = np.ones(count, dtype=bool)
    while todo.any():
        v = x[todo]
        l, r = v >> half, v & m
        for k in range(4):
            l, r = r, l ^ (_hash32(r ^ keys[k]) & m)
        v = (l << half) | r
        x[todo] = v
        todo[todo] = v >= np.uint64(n)
    return x.astype(np.int64)
