"""TEST INFRASTRUCTURE ONLY -- golden vectors for the per-step pixel sampling from the UNMODIFIED reference method
SceneDataset.__getitem__ (code/datasets/scene_hawp_dataset.py:148-194), run in the build container on an instance whose
per-image tables come from tests/golden/hawp_abc.npz (themselves outputs of the reference's
compute_point_line_attraction, see make_golden_hawp.py) plus a seeded random image.   python oracle/make_golden_dataset.py"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

R = 256
SEED = 42


def load_reference_dataset_module():
    ref_shim.install()
    hawp, base, C = types.ModuleType("hawp"), types.ModuleType("hawp.base"), types.ModuleType("hawp.base._C")
    base._C = C
    hawp.base = base
    sys.modules.update({"hawp": hawp, "hawp.base": base, "hawp.base._C": C})
    spec = importlib.util.spec_from_file_location("ref_scene_hawp_dataset",
                                                  os.path.join(ref_shim.REF_CODE, "datasets", "scene_hawp_dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def tables(dist=5.0):
    g = np.load(os.path.join(ROOT, "tests", "golden", "hawp_abc.npz"))
    H, W = (int(v) for v in g["img_res"])
    mask = np.unpackbits(g["mask_%g" % dist])[:H * W].astype(bool)
    labels = g["labels_%g" % dist].astype(np.int64)
    att = np.zeros((H * W, 2), dtype=np.float32)
    att[g["proj_idx_%g" % dist]] = g["proj_val_%g" % dist]
    rgb = np.random.default_rng(7).random((H * W, 3), dtype=np.float32)
    return (H, W), g["lines"].astype(np.float32), mask, labels, att, rgb


def main():
    mod = load_reference_dataset_module()
    (H, W), lines, mask, labels, att, rgb = tables()

    class Wire:  # stands in for WireframeGraph: __getitem__ only reads .vertices
        vertices = torch.zeros(8, 2)

    ds = mod.SceneDataset.__new__(mod.SceneDataset)       # no __init__: it reads image folders
    ds.img_res, ds.total_pixels = [H, W], H * W
    ds.lines, ds.masks, ds.labels = [torch.from_numpy(lines)], [torch.from_numpy(mask)], [torch.from_numpy(labels)]
    ds.att_points, ds.rgb_images, ds.wireframes = [torch.from_numpy(att)], [torch.from_numpy(rgb)], [Wire()]
    ds.intrinsics_all, ds.pose_all = [torch.eye(4)], [torch.eye(4)]
    ds.sampling_idx = None
    gold = {"R": np.array(R), "seed": np.array(SEED)}
    _, s_full, gt_full = ds[0]
    gold["full_uv_head"] = s_full["uv"][:2 * W + 3].numpy()          # enough rows to pin the (column, row) order
    gold["full_lines_head"] = s_full["lines"][:2 * W + 3].numpy()
    assert gt_full["rgb"].shape[0] == H * W and "lines2d" not in gt_full
    torch.manual_seed(SEED)
    ds.change_sampling_idx(R)                                      # consumes one randperm(total_pixels), as in training
    _, s, gt = ds[0]
    # the permutation the reference drew, replayed
    torch.manual_seed(SEED)
    torch.randperm(H * W)
    n = int(mask.sum())
    gold["perm"] = torch.randperm(n)[:R].numpy()
    for k in ("uv", "uv_proj", "labels", "lines"):
        gold["s_" + k] = s[k].numpy()
    gold["gt_rgb"], gold["gt_lines2d"] = gt["rgb"].numpy(), gt["lines2d"].numpy()
    out = os.path.join(ROOT, "tests", "golden", "dataset_abc.npz")
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out), "masked pixels", n)


if __name__ == "__main__":
    main()
