"""TEST INFRASTRUCTURE ONLY -- golden vectors for the dense-grid SDF evaluation from the UNMODIFIED
reference functions `get_grid_uniform` and `get_surface_trace` (code/utils/plots.py:101-108, 318-329): the grid points
(their float arithmetic and their ORDER) and the raveled `z` array the reference feeds to marching cubes, for an analytic
sdf.  plotly / skimage / trimesh are stubbed (never called: the analytic sdf has no zero crossing, so the marching-cubes
branch is skipped).      python oracle/make_golden_grid.py"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def load_plots():
    ref_shim.install()
    plotly = sys.modules["plotly"]
    plotly.__path__ = []
    for name in ("plotly.graph_objs", "plotly.offline", "plotly.subplots"):
        m = types.ModuleType(name)
        sys.modules[name] = m
        setattr(plotly, name.split(".")[1], m)
    sys.modules["plotly.subplots"].make_subplots = None
    sk = sys.modules["skimage"]
    sk.__path__ = []
    sk.measure = types.ModuleType("skimage.measure")
    sys.modules["skimage.measure"] = sk.measure
    if "torchvision" not in sys.modules:
        try:
            import torchvision  # noqa: F401
        except Exception:
            sys.modules["torchvision"] = types.ModuleType("torchvision")
    return importlib.import_module("utils.plots")


def analytic_sdf(x):
    return (x * x).sum(-1, keepdim=True).sqrt() + 3.0 + 0.25 * torch.sin(3.0 * x[:, :1]) * x[:, 1:2]   # > 0 everywhere


def main():
    plots = load_plots()
    gold = {}
    for res, bound in ((7, [-1.3, 1.7]), (100, [-1.5, 1.5])):
        g = plots.get_grid_uniform(res, bound)
        pts = g["grid_points"]
        if res == 7:
            gold["points_7"] = pts.numpy()
        else:   # 10^6 points: a strided sample and a checksum are enough to pin arithmetic and order
            gold["points_100_stride997"] = pts.numpy()[::997]
            gold["points_100_sum"] = pts.double().sum(0).numpy()
        seen = []
        plots.get_surface_trace(None, 0, lambda p: (seen.append(p.shape[0]), analytic_sdf(p)[:, 0])[1], resolution=res,
                                grid_boundary=bound)
        gold["chunks_%d" % res] = np.asarray(seen)
    z = analytic_sdf(torch.from_numpy(gold["points_7"]))[:, 0]
    gold["z_7"] = z.numpy()
    out = os.path.join(ROOT, "tests", "golden", "grid.npz")
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out), {k: v.shape for k, v in gold.items()})


if __name__ == "__main__":
    main()
