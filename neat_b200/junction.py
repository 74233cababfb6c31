"""Host step of the junction block (neat_wfr_rend_a.py:466-484, loss_wfr.py:104-108) through the C ABI
(csrc/junction.cpp): numpy arrays in, numpy arrays out; no device work."""
import ctypes

import numpy as np

from . import _lib

_P = ctypes.c_void_p


def _p(a):
    return _P(a.ctypes.data)


def linear_sum_assignment(cost):
    """Drop-in for scipy.optimize.linear_sum_assignment (minimisation): (row_ind, col_ind), rows ascending."""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    if cost.ndim != 2:
        raise ValueError("expected a matrix (2-D array), got a %r array" % (cost.shape,))
    nr, nc = cost.shape
    n = min(nr, nc)
    rows, cols = np.empty(n, np.int32), np.empty(n, np.int32)
    got = _lib.load().neat_linear_sum_assignment(_p(cost), nr, nc, _p(rows), _p(cols))
    if got != n:
        raise ValueError("cost matrix is infeasible or contains invalid numeric entries")
    return rows.astype(np.int64), cols.astype(np.int64)


def junction_match(centroids, gt_vertices, pose, intrinsics, global_junctions, use_median=False):
    """-> (local [n,7] = xyz | uv | uv_calib, rows [n], cols [n], n_close, median)."""
    cent = np.ascontiguousarray(centroids, dtype=np.float32).reshape(-1, 3)
    gt = np.ascontiguousarray(gt_vertices, dtype=np.float32).reshape(-1, 2)
    pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(16)
    K = np.ascontiguousarray(intrinsics, dtype=np.float32).reshape(16)
    glob = np.ascontiguousarray(global_junctions, dtype=np.float32).reshape(-1, 3)
    cap = max(1, min(cent.shape[0], gt.shape[0]))
    local = np.empty((cap, 7), np.float32)
    rows, cols = np.empty(cap, np.int32), np.empty(cap, np.int32)
    n, close = ctypes.c_int(0), ctypes.c_int(0)
    med = ctypes.c_float(10.0)
    _lib.check(_lib.load().neat_junction_match(_p(cent), cent.shape[0], _p(gt), gt.shape[0], _p(pose), _p(K), _p(glob),
                                                glob.shape[0], int(bool(use_median)), _p(local), ctypes.byref(n),
                                                _p(rows), _p(cols), ctypes.byref(close), ctypes.byref(med)))
    k = min(n.value, glob.shape[0])
    return local[:n.value], rows[:k].astype(np.int64), cols[:k].astype(np.int64), close.value, med.value
