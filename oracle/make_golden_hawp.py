"""TEST INFRASTRUCTURE ONLY -- golden vectors for the attraction precompute from the UNMODIFIED reference method
SceneDataset.compute_point_line_attraction (code/datasets/scene_hawp_dataset.py:92-146), run in the build container
with `hawp.base._C.encodels` (a CUDA extension that cannot be built here) replaced by oracle/hawp_oracle.encodels.
Input: the in-repo fixture data/abc/00075213/hawp/image_0000.json.   python oracle/make_golden_hawp.py"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import hawp_oracle, ref_shim  # noqa: E402


def main():
    ref_shim.install()
    # stub `hawp.base._C` with the CPU restatement
    hawp, base, C = types.ModuleType("hawp"), types.ModuleType("hawp.base"), types.ModuleType("hawp.base._C")

    def encodels(lines, ih, iw, h, w, n):
        m, l, t = hawp_oracle.encodels(lines.cpu().numpy(), ih, iw, h, w, n)
        return torch.from_numpy(m), torch.from_numpy(l), torch.from_numpy(t)

    C.encodels = encodels
    base._C = C
    hawp.base = base
    sys.modules.update({"hawp": hawp, "hawp.base": base, "hawp.base._C": C})
    spec = importlib.util.spec_from_file_location("ref_scene_hawp_dataset",
                                                  os.path.join(ref_shim.REF_CODE, "datasets", "scene_hawp_dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    wf = json.load(open(os.path.join(ref_shim.REF_ROOT, "data/abc/00075213/hawp/image_0000.json")))
    v = np.asarray(wf["vertices"], dtype=np.float32)
    e = np.asarray(wf["edges"], dtype=np.int64)
    w = np.asarray(wf["edges-weights"], dtype=np.float32)
    lines = np.concatenate([v[e[:, 0]], v[e[:, 1]], w[:, None]], axis=1).astype(np.float32)
    H, W = int(wf["height"]), int(wf["width"])
    gold = {"lines": lines, "img_res": np.array([H, W])}
    for dist in (5.0, 20.0):
        ds = mod.SceneDataset.__new__(mod.SceneDataset)       # no __init__: it reads image folders
        ds.img_res, ds.distance = [H, W], dist
        mask, labels, proj = mod.SceneDataset.compute_point_line_attraction(ds, torch.from_numpy(lines))
        gold["mask_%g" % dist] = np.packbits(mask.numpy())
        gold["labels_%g" % dist] = labels.numpy().astype(np.int16)
        idx = np.nonzero(mask.numpy())[0]
        gold["proj_idx_%g" % dist] = idx.astype(np.int32)
        gold["proj_val_%g" % dist] = proj.numpy()[idx]
        print("distance", dist, "masked pixels", int(mask.sum()))
    out = os.path.join(ROOT, "tests", "golden", "hawp_abc.npz")
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
