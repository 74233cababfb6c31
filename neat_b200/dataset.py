"""Device-resident scene tables and the per-step pixel sampling (second half): the work of
`SceneDataset.__getitem__` / `change_sampling_idx` / `collate_fn` (code/datasets/scene_hawp_dataset.py:148-221) without
the per-step CPU `mask.nonzero()`, `randperm`, fancy-index gathers and host-to-device copies.

    scene = DeviceScene(img_res, device="cuda:0")
    scene.add_image(rgb[HW,3], lines[n,5], intrinsics[4,4], pose[4,4], wireframe, distance_threshold=5.0)
    scene.change_sampling_idx(1024)              # same name and meaning as the reference (-1: full image)
    idx, sample, ground_truth = scene[i]         # same keys as the reference; every tensor already on the device
    loader = DataLoader(scene, batch_size=1, shuffle=True, collate_fn=scene.collate_fn)   # volsdf_train.py:155-159

File loading (images, cameras.npz, the HAWP json files, `SceneDataset.__init__` :18-91) stays with the reference's loaders:
pass their arrays to `add_image`.  `rng = "reference"` (default) draws the subset with the same CPU-generator call as
the reference (`torch.randperm(n_masked)[:R]`, :176), so the same seed gives the same rays; `rng = "reference-numpy"` is
BlenderDataset's draw (`np.random.choice(sampling_idx, R)`, WITH replacement, numpy's global generator;
code/datasets/blender_hawp_dataset.py:43 -- the dataset class of abc-neat-a.conf); `rng = "device"` draws R distinct masked
pixels inside the gather kernel (keyed bijection, csrc/pixels.cuh) with no host work at all."""
import ctypes

import torch

from . import _lib, attraction

_P = ctypes.c_void_p


def _stream(dev):
    return _P(torch.cuda.current_stream(dev).cuda_stream)


def nonzero_mask(mask):
    """Ascending indices of the set entries of a bool/uint8 CUDA tensor (int32), like mask.nonzero().flatten()."""
    if not mask.is_cuda:
        raise _lib.NeatError("nonzero_mask: `mask` must be a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    m = mask.detach().reshape(-1).contiguous().view(torch.uint8) if mask.dtype == torch.bool else \
        mask.detach().reshape(-1).to(torch.uint8).contiguous()
    n, dev = m.numel(), m.device
    if n == 0:
        return torch.empty(0, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.neat_mask_compact_workspace_bytes(n), dtype=torch.uint8, device=dev)
    out = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.neat_mask_compact(_P(m.data_ptr()), n, _P(ws.data_ptr()), _P(out.data_ptr()), _P(cnt.data_ptr()),
                                     _stream(dev)))
    return out[:int(cnt.item())].clone()      # once per image, at load time


def pixel_permutation(n, seed, step, first, count):
    """Host evaluation of the bijection the kernel uses for rng="device": positions of rays first .. first+count-1."""
    lib = _lib.load()
    out = (ctypes.c_uint * count)()
    _lib.check(lib.neat_pixel_permutation(n, seed, step, first, count, out))
    return list(out)


def reference_numpy_positions(n, R):
    """Positions into the masked-pixel list that `np.random.choice(sampling_idx, R)` (blender_hawp_dataset.py:43) picks:
    the legacy generator draws `randint(0, n, R)` and indexes the array, so choosing positions consumes the same stream."""
    import numpy as np
    return np.random.choice(n, R).astype(np.int64)


class _Image:
    __slots__ = ("rgb", "lines", "mask", "labels", "att_points", "masked", "intrinsics", "pose", "wireframe")


class DeviceScene(torch.utils.data.Dataset):
    def __init__(self, img_res, device="cuda:0", rng="reference", seed=0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NeatError("DeviceScene needs a CUDA device (no CPU fallback)")
        _lib.load()
        self.img_res = [int(img_res[0]), int(img_res[1])]
        self.total_pixels = self.img_res[0] * self.img_res[1]
        self.sampling_size = None          # the reference keeps a dummy index tensor; only its length is ever used
        self.rng = rng
        self.seed = int(seed)
        self.draws = 0                      # rng="device": one bijection per __getitem__ call
        self.images = []

    # ------------------------------------------------------------------ construction
    def add_image(self, rgb, lines, intrinsics, pose, wireframe=None, distance_threshold=5.0, tables=None):
        """rgb [HW,3] in the reference's layout (rend_util.load_rgb + reshape(3,-1).T, :64-66), lines [n,5]
        (`wireframe.line_segments(score_threshold)`, :75).  The attraction tables (mask, labels, att_points) are
        computed on the device (`attraction.compute_point_line_attraction`, :86-90) unless given as `tables`."""
        im = _Image()
        dev = self.device
        im.rgb = torch.as_tensor(rgb, dtype=torch.float32).to(dev).reshape(-1, 3).contiguous()
        im.lines = torch.as_tensor(lines, dtype=torch.float32).to(dev).contiguous()
        if im.rgb.shape[0] != self.total_pixels or im.lines.dim() != 2 or im.lines.shape[1] != 5 or im.lines.shape[0] == 0:
            raise _lib.NeatError("add_image: rgb must be [H*W,3] and lines [n>0,5]")
        if tables is None:
            tables = attraction.compute_point_line_attraction(im.lines, self.img_res, distance_threshold)
        mask, labels, att = tables
        im.mask = torch.as_tensor(mask).to(dev).reshape(-1).bool().contiguous()
        im.labels = torch.as_tensor(labels).to(dev).reshape(-1).long().contiguous()
        im.att_points = torch.as_tensor(att, dtype=torch.float32).to(dev).reshape(-1, 2).contiguous()
        if not (im.mask.numel() == im.labels.numel() == im.att_points.shape[0] == self.total_pixels):
            raise _lib.NeatError("add_image: attraction tables do not match img_res")
        im.masked = nonzero_mask(im.mask)
        im.intrinsics = torch.as_tensor(intrinsics, dtype=torch.float32).to(dev)
        im.pose = torch.as_tensor(pose, dtype=torch.float32).to(dev)
        im.wireframe = wireframe
        self.images.append(im)
        return len(self.images) - 1

    def __len__(self):
        return len(self.images)

    def change_sampling_idx(self, sampling_size):
        """scene_hawp_dataset.py:216-220.  The reference also draws a `randperm(total_pixels)` here that nothing reads
        (only its length is used, :176); with rng="reference" the draw is repeated so the generator stays in step."""
        if sampling_size == -1:
            self.sampling_size = None
        else:
            self.sampling_size = int(sampling_size)
            if self.rng in ("reference", "reference-numpy"):
                torch.randperm(self.total_pixels)

    # ------------------------------------------------------------------ one item
    def _gather(self, im, R, first=0, perm=None, draw=None, with_lines2d=True):
        lib = _lib.load()
        dev = self.device
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        uv, uvp, rgb = f(R, 2), f(R, 2), f(R, 3)
        l2d = f(R, 5) if with_lines2d else None
        lab = torch.empty(R, dtype=torch.int64, device=dev)
        idx = torch.empty(R, dtype=torch.int64, device=dev)
        masked = perm is not None or draw is not None
        a = _lib.PixelArgs(R=R, W=self.img_res[1], first=first,
                           masked=_P(im.masked.data_ptr()) if masked else None, n_masked=int(im.masked.numel()),
                           perm=_P(perm.data_ptr()) if perm is not None else None,
                           seed=draw[0] if draw else 0, step=draw[1] if draw else 0,
                           rgb_image=_P(im.rgb.data_ptr()), labels=_P(im.labels.data_ptr()),
                           att_points=_P(im.att_points.data_ptr()), lines=_P(im.lines.data_ptr()),
                           n_lines=int(im.lines.shape[0]), uv=_P(uv.data_ptr()), uv_proj=_P(uvp.data_ptr()),
                           rgb=_P(rgb.data_ptr()), lines2d=_P(l2d.data_ptr()) if with_lines2d else None,
                           labels_out=_P(lab.data_ptr()), index_out=_P(idx.data_ptr()))
        _lib.check(lib.neat_sample_pixels(ctypes.byref(a), _stream(dev)))
        return uv, uvp, rgb, l2d, lab, idx

    def __getitem__(self, i):
        im = self.images[i]
        sample = {"intrinsics": im.intrinsics, "pose": im.pose, "wireframe": im.wireframe, "mask": im.mask,
                  "lines_uniq": im.lines}
        if im.wireframe is not None and hasattr(im.wireframe, "vertices"):
            sample["juncs2d"] = im.wireframe.vertices
        if self.sampling_size is None:        # full image (plots / evaluation): uv grid, every table as is
            uv, uvp, rgb, l2d, lab, _ = self._gather(im, self.total_pixels)
            sample.update(uv=uv, uv_proj=uvp, labels=lab, lines=l2d)
            return i, sample, {"rgb": rgb}
        R, n = self.sampling_size, int(im.masked.numel())
        if self.rng == "reference-numpy":
            if n == 0:
                raise _lib.NeatError("image %d has no masked pixel" % i)
            perm = torch.from_numpy(reference_numpy_positions(n, R)).pin_memory().to(self.device, non_blocking=True)
            uv, uvp, rgb, l2d, lab, idx = self._gather(im, R, perm=perm)
            sample.update(uv=uv, uv_proj=uvp, labels=lab, lines=l2d, sampling_idx=idx)
            return i, sample, {"rgb": rgb, "lines2d": l2d}
        if R > n:
            raise _lib.NeatError("image %d has %d masked pixels, fewer than the %d requested" % (i, n, R))
        if self.rng == "reference":
            perm = torch.randperm(n)[:R].pin_memory().to(self.device, non_blocking=True)
            uv, uvp, rgb, l2d, lab, idx = self._gather(im, R, perm=perm)
        elif self.rng == "device":
            self.draws += 1
            uv, uvp, rgb, l2d, lab, idx = self._gather(im, R, draw=(self.seed, self.draws))
        else:
            raise _lib.NeatError("rng must be 'reference', 'reference-numpy' or 'device'")
        sample.update(uv=uv, uv_proj=uvp, labels=lab, lines=l2d, sampling_idx=idx)
        return i, sample, {"rgb": rgb, "lines2d": l2d}

    def full_image_chunk(self, i, first, count):
        """Pixels first .. first+count-1 of image i (what utils.split_input, code/utils/general.py:56-74, slices out of
        the full-image item) without materialising the whole image's inputs."""
        if first < 0 or count <= 0 or first + count > self.total_pixels:
            raise _lib.NeatError("chunk outside the image")
        uv, uvp, rgb, l2d, lab, _ = self._gather(self.images[i], count, first=first)
        return {"uv": uv, "uv_proj": uvp, "labels": lab, "lines": l2d}, {"rgb": rgb}

    @staticmethod
    def collate_fn(batch_list):
        """scene_hawp_dataset.py:196-214: dict entries are stacked per key (tensors) or listed (everything else)."""
        out = []
        for entry in zip(*batch_list):
            if isinstance(entry[0], dict):
                out.append({k: torch.stack([o[k] for o in entry]) if isinstance(entry[0][k], torch.Tensor)
                            else [o[k] for o in entry] for k in entry[0]})
            else:
                out.append(torch.LongTensor(entry))
        return tuple(out)
