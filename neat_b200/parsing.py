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


def wireframe_from_lines_and_junctions(lines, junctions, rel_matching_distance_threshold=0.01):
    """get_wireframe_from_lines_and_junctions (neat-final-parsing.py:128-157) on the GPU: lines [N,2,3], junctions [J,3]
    -> (graph [J,J] float, lines3d_wf [E,2,3] = junctions[graph.triu().nonzero()])."""
    from . import dataset
    lib = _lib.load()
    dev = junctions.device
    if dev.type != "cuda":
        raise _lib.NeatError("neat_b200.parsing runs on CUDA tensors only (no CPU path)")
    l3 = lines.detach().to(dev, torch.float32).reshape(-1, 6).contiguous()
    jn = junctions.detach().to(dev, torch.float32).reshape(-1, 3).contiguous()
    N, J = l3.shape[0], jn.shape[0]
    if J == 0:
        return torch.zeros(0, 0, device=dev), torch.zeros(0, 2, 3, device=dev)
    midx = torch.empty(max(N, 1), 2, dtype=torch.int32, device=dev)
    matched = torch.empty(max(N, 1), dtype=torch.uint8, device=dev)
    graph = torch.empty(J, J, device=dev)
    upper = torch.empty(J, J, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.neat_line_junction_graph(_P(l3.data_ptr()), N, _P(jn.data_ptr()), J,
                                                float(rel_matching_distance_threshold), _P(midx.data_ptr()),
                                                _P(matched.data_ptr()), _P(graph.data_ptr()), _P(upper.data_ptr()),
                                                _P(torch.cuda.current_stream(dev).cuda_stream)))
        pairs = dataset.nonzero_mask(upper).long()            # row-major, like graph.triu().nonzero()
    return graph, jn[torch.stack((pairs // J, pairs % J), dim=1)]


@torch.no_grad()
def initial_recon(model, eval_dataloader, chunksize, *, line_dis_threshold=10, line_score_threshold=0.01,
                  junc_match_threshold=0.05, sdf_junction_refine=True, device="cuda", **kwargs):
    """initial_recon of code/neat-final-parsing.py:159-295, same arguments and result dict.  `eval_dataloader` yields the
    reference's (indices, model_input, ground_truth) items (full image + `mask`; a neat_b200.dataset.DeviceScene with
    change_sampling_idx(-1) works); the forward over the masked pixels runs in chunks of `chunksize` through `model`
    (the plugin in eval mode), voting / scoring / graph construction in the kernels of csrc/parsing.cuh, the end-point /
    junction assignment in the native host solver."""
    model.eval()
    dev = torch.device(device)
    if sdf_junction_refine:
        global_junctions, _, _ = refine_global_junctions(model)
    else:
        global_junctions = model.ffn(model.latents).detach()
    global_junctions = global_junctions.to(dev)
    votes = {}                                  # junction index -> number of matched end points, in first-seen order
    lines3d_all, scores_all = [], []
    for indices, model_input, ground_truth in eval_dataloader:
        mask = model_input["mask"][0].to(dev)
        uv = model_input["uv"].to(dev)[:, mask]                                        # :203
        uv_proj = model_input["uv_proj"].to(dev)[:, mask]
        l3, l2, p3 = [], [], []
        for a in range(0, uv.shape[1], chunksize):                                     # utils.split_input, :210
            s = dict(model_input)
            s["intrinsics"], s["pose"] = model_input["intrinsics"].to(dev), model_input["pose"].to(dev)
            s["uv"], s["uv_proj"] = uv[:, a:a + chunksize], uv_proj[:, a:a + chunksize]
            out = model(s)
            l3.append(out["lines3d"].detach().reshape(-1, 2, 3))
            l2.append(out["lines2d"].detach().reshape(-1, 4))
            p3.append(out["l3d"].detach().reshape(-1, 3))
        if not l3:
            continue
        gt_lines = model_input["wireframe"][0].line_segments(0.01).to(dev)[:, :-1]        # :232
        _, lines3d, scores, _ = vote_lines(torch.cat(l2), torch.cat(l3), torch.cat(p3), gt_lines, line_dis_threshold)
        if lines3d.shape[0] > 0:
            for ai, _ in match_endpoints(global_junctions, lines3d, junc_match_threshold):   # :262-268
                votes[ai] = votes.get(ai, 0) + 1
            lines3d_all.append(lines3d)
            scores_all.append(scores)
    if not lines3d_all:
        raise _lib.NeatError("initial_recon: no view produced a line vote")
    lines3d_all, scores_all = torch.cat(lines3d_all, dim=0), torch.cat(scores_all, dim=0)
    lines3d_all = lines3d_all[scores_all < line_score_threshold]                        # :276
    keep = [k for k, n in votes.items() if n > 1]                                        # :288
    if not keep:
        raise _lib.NeatError("initial_recon: no global junction received two end-point votes")
    junctions3d_initial = global_junctions[torch.tensor(keep, device=dev)]
    graph_initial, lines3d_wfi = wireframe_from_lines_and_junctions(lines3d_all, junctions3d_initial,
                                                                    rel_matching_distance_threshold=0)
    return {"junctions3d_initial": junctions3d_initial, "lines3d_all": lines3d_all, "graph_initial": graph_initial,
            "lines3d_wfi": lines3d_wfi}


@torch.no_grad()
def visibility_checking(lines3d_all, eval_dataloader, model=None, *, mindis_th=25, min_visible_views=1, device="cuda"):
    """visibility_checking of code/neat-final-parsing.py:305-337: keep the 3D lines whose projection lies within
    `mindis_th` (squared pixels, either end-point order) of a detected 2D line in at least `min_visible_views` views.
    One fused launch per view; `model` is accepted for signature compatibility (the projection is in the kernel)."""
    dev = torch.device(device)
    lines3d_all = lines3d_all.to(dev)
    seen = torch.zeros(lines3d_all.shape[0], dtype=torch.int32, device=dev)
    for indices, model_input, ground_truth in eval_dataloader:
        gt = model_input["wireframe"][0].line_segments(0.05).to(dev)[:, :4]              # :311
        vis, _ = line_visibility(lines3d_all, model_input["pose"][0].to(dev), model_input["intrinsics"][0].to(dev), gt, mindis_th)
        seen += vis.int()
    return lines3d_all[seen >= min_visible_views]
