"""Dense-grid SDF evaluation for mesh extraction: what code/utils/plots.py:101-108 does with
`sdf = lambda x: implicit_network(x)[:, 0]` over `get_grid_uniform(resolution, grid_boundary)['grid_points']` in chunks of
100,000 points, as ONE launch of the fused SDF query kernel that generates the grid points itself."""
import ctypes

import numpy as np
import torch

from . import _lib

_P = ctypes.c_void_p


def sdf_grid(model, resolution=100, grid_boundary=(-2.0, 2.0), clamp=False):
    """-> float32 CUDA tensor [resolution**3] in the order of the reference's raveled np.meshgrid(x, y, x) grid, i.e.
    `z` of plots.get_surface_trace before its reshape (ny, nx, nz)."""
    n = (int(resolution),) * 3 if np.isscalar(resolution) else tuple(int(r) for r in resolution)
    lo = (ctypes.c_double * 3)(*([float(grid_boundary[0])] * 3))
    hi = (ctypes.c_double * 3)(*([float(grid_boundary[1])] * 3))
    nn = (ctypes.c_int * 3)(*n)
    rn = model._sync_weights()
    ctx = rn.ctx
    out = torch.empty(n[0] * n[1] * n[2], device=ctx.device)
    with torch.cuda.device(ctx.device):
        _lib.check(ctx.lib.neat_sdf_grid(ctx._h, lo, hi, nn, int(bool(clamp)), _P(out.data_ptr()), ctx._stream()))
    return out


def grid_points(resolution=100, grid_boundary=(-2.0, 2.0)):
    """The reference's get_grid_uniform(...)['grid_points'] (plots.py:318-324), for callers that want the coordinates."""
    x = np.linspace(grid_boundary[0], grid_boundary[1], resolution)
    xx, yy, zz = np.meshgrid(x, x, x)
    return torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)
