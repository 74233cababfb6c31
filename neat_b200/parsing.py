"""Wireframe finalisation helpers: the per-image voting step of code/neat-final-parsing.py's
initial_recon (:226-271) on the GPU.  The forward over the image's masked pixels is the plugin's eval-mode forward
(neat_b200.model.VolSDFNetwork); this module turns its outputs into per-ground-truth-line 3D line candidates."""
import ctypes

import numpy as np
import torch

from . import _lib, junction

_P = ctypes.c_void_p


def vote_lines(lines2d, lines3d, points3d, gt_lines, line_dis_threshold=10.0):
    """lines2d [N,4] | [N,2,2], lines3d [N,2,3], points3d [N,3] (out['l3d']), gt_lines [G,4]: CUDA fp32 tensors.
    Every prediction votes (in both end-point orders) for its nearest ground-truth 2D line; votes with squared distance
    < line_dis_threshold are kept (neat-final-parsing.py:226-260).  Returns (labels [K] ascending, lines3d_mean [K,2,3],
    scores [K], counts [K]) for the ground-truth lines that received votes, like the reference's per-label loop."""
    lib = _lib.load()
    dev = lines2d.device
    if dev.type != "cuda":
        raise _lib.NeatError("neat_b200.parsing runs on CUDA tensors only (no CPU path)")
    f = lambda t, w: t.detach().to(dev, torch.float32).reshape(-1, w).contiguous()
    l2, l3, p3, gt = f(lines2d, 4), f(lines3d, 6), f(points3d, 3), f(gt_lines, 4)
    N, G = l2.shape[0], gt.shape[0]
    if N == 0 or G == 0:
        z = torch.zeros
        return z(0, dtype=torch.long, device=dev), z(0, 2, 3, device=dev), z(0, device=dev), z(0, dtype=torch.long, device=dev)
    if l3.shape[0] != N or p3.shape[0] != N:
        raise ValueError("lines2d, lines3d and points3d must describe the same N predictions")
    ws = torch.empty(int(lib.neat_line_vote_workspace_bytes(N, G)), dtype=torch.uint8, device=dev)
    mean = torch.empty(G, 6, device=dev)
    scores = torch.empty(G, device=dev)
    counts = torch.empty(G, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.neat_line_vote(_P(l2.data_ptr()), _P(l3.data_ptr()), _P(p3.data_ptr()), N, _P(gt.data_ptr()), G,
                                      float(line_dis_threshold), _P(ws.data_ptr()), _P(mean.data_ptr()),
                                      _P(scores.data_ptr()), _P(counts.data_ptr()),
                                      _P(torch.cuda.current_stream(dev).cuda_stream)))
    labels = (counts > 0).nonzero().flatten()
    return labels, mean[labels].view(-1, 2, 3), scores[labels], counts[labels].long()


def match_endpoints(global_junctions, lines3d, junc_match_threshold=0.05):
    """neat-final-parsing.py:262-268: optimal assignment of the voted 3D end points to the global junctions (native host
    solver, csrc/junction.cpp).  Returns [(junction index, end-point index)] closer than the threshold."""
    endpoints = lines3d.reshape(-1, 3)
    if endpoints.shape[0] == 0 or global_junctions.shape[0] == 0:
        return []
    cd = torch.cdist(global_junctions.float(), endpoints.float()).cpu().numpy()
    ai, aj = junction.linear_sum_assignment(cd)
    return [(int(a), int(b)) for a, b in zip(ai, aj) if cd[a, b] < junc_match_threshold]


@torch.no_grad()
def refine_global_junctions(model, sdf_threshold=0.05):
    """neat-final-parsing.py:171-184: one SDF Newton step of ffn(latents), sorted by |sdf| order of the reference.
    Returns (global_junctions [J,3], sdf [J], is_valid [J])."""
    gj = model.ffn(model.latents).detach()
    sdf, _, grad = model.implicit_network.get_outputs(gj)
    gj = (gj - sdf.reshape(-1, 1) * grad).detach()
    s = model.implicit_network.get_sdf_vals(gj).flatten()
    order = torch.argsort(s)
    gj, s = gj[order], s[order]
    return gj, s, s.abs() < sdf_threshold


def line_visibility(lines3d, pose, intrinsics, gt_lines, mindis_th=25.0):
    """One view of visibility_checking (neat-final-parsing.py:314-335): lines3d [L,2,3], pose [4,4] camera-to-world,
    intrinsics [4,4] | [3,3], gt_lines [G,4] -> (visible [L] bool, mindis [L]).  The caller ORs / counts over views
    (`lines3d_visibility.sum(dim=1) >= min_visible_views`, :336)."""
    lib = _lib.load()
    dev = lines3d.device
    if dev.type != "cuda":
        raise _lib.NeatError("neat_b200.parsing runs on CUDA tensors only (no CPU path)")
    l3 = lines3d.detach().to(dev, torch.float32).reshape(-1, 6).contiguous()
    gt = gt_lines.detach().to(dev, torch.float32).reshape(-1, 4).contiguous()
    K = intrinsics.detach().to(dev, torch.float32).contiguous()
    pose_inv = torch.linalg.inv(pose.detach().to(dev, torch.float32)).contiguous()   # pose.inverse(), :320
    L, Gn = l3.shape[0], gt.shape[0]
    vis = torch.zeros(L, dtype=torch.uint8, device=dev)
    mind = torch.full((L,), float("inf"), device=dev)
    if L and Gn:
        with torch.cuda.device(dev):
            _lib.check(lib.neat_line_visibility(_P(l3.data_ptr()), L, _P(pose_inv.data_ptr()), _P(K.data_ptr()), K.shape[-1],
                                                _P(gt.data_ptr()), Gn, float(mindis_th), _P(vis.data_ptr()),
                                                _P(mind.data_ptr()), _P(torch.cuda.current_stream(dev).cuda_stream)))
    return vis.bool(), mind
