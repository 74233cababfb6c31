"""TEST INFRASTRUCTURE ONLY (imported by tests/): CPU restatement of the reference's wireframe finalisation,
code/neat-final-parsing.py: junction refinement (:171-184), the per-image voting block of initial_recon (:226-271), its
assembly (:274-295), get_wireframe_from_lines_and_junctions (:128-157) and visibility_checking (:305-337), statement by
statement with the same torch ops in the same order.

Parity status: PINNED -- oracle/make_golden_parsing.py runs the UNMODIFIED reference functions on a synthetic scene (a
replay model returns stored per-pixel eval outputs) and stores their results in tests/golden/parsing_synth.npz;
tests/test_parsing_oracle.py checks initial_recon() / visibility_checking() below against them."""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment


def _legacy_cross(a, b):
    """torch.cross(a, b) WITHOUT dim, as the reference calls it (:256): the deprecated default is the first dimension of
    size 3 -- the last one for [n,3] inputs unless n == 3, where it is dimension 0 (a reference quirk for labels with
    exactly three votes; the CUDA path reproduces it, csrc/parsing.cuh line_vote_finish_kernel)."""
    dim = next(i for i, n in enumerate(a.shape) if n == 3)
    return torch.cross(a, b, dim=dim)


def vote_lines(lines2d, lines3d, points3d, gt_lines, line_dis_threshold=10.0):
    """lines2d [N,4], lines3d [N,2,3], points3d [N,3] (out['l3d']), gt_lines [G,4]  (neat-final-parsing.py:226-260).
    Returns (labels [K] sorted, lines3d_mean [K,2,3], scores [K], counts [K])."""
    lines3d = torch.cat((lines3d, lines3d[:, [1, 0]]), dim=0)                      # :226
    lines2d = torch.cat((lines2d, lines2d[:, [2, 3, 0, 1]]), dim=0)                # :228
    points3d = torch.cat([points3d, points3d])                                    # :231
    dis = torch.sum((lines2d[:, None] - gt_lines[None]) ** 2, dim=-1)             # :234
    mindis, minidx = dis.min(dim=1)                                               # :236
    keep = mindis < line_dis_threshold
    labels = minidx[keep].unique()                                                # :238
    lines3d_valid, points3d_valid, assignment = lines3d[keep], points3d[keep], minidx[keep]
    out_l, out_s, out_c = [], [], []
    for label in labels:                                                          # :246-258
        idx = (assignment == label).nonzero().flatten()
        if idx.numel() == 0:
            continue
        val = lines3d_valid[idx].mean(dim=0)
        support_pts = points3d_valid[idx]
        support_dis = torch.norm(_legacy_cross(support_pts - val[:1], support_pts - val[1:]), dim=-1) / \
            torch.norm(val[1] - val[0]).clamp_min(1e-6)
        out_l.append(val)
        out_s.append(support_dis.mean())
        out_c.append(idx.numel())
    if not out_l:
        return labels, torch.zeros(0, 2, 3), torch.zeros(0), torch.zeros(0, dtype=torch.long)
    return labels, torch.stack(out_l, 0), torch.stack(out_s), torch.tensor(out_c)


def match_endpoints(global_junctions, lines3d, junc_match_threshold=0.05):
    """neat-final-parsing.py:262-268: Hungarian matching of the voted 3D line end points to the global junctions.
    Returns [(junction index, end-point index)] of the matches closer than the threshold."""
    endpoints = lines3d.reshape(-1, 3)
    cdist = torch.cdist(global_junctions, endpoints)
    ai, aj = linear_sum_assignment(cdist.cpu().numpy())
    return [(int(a), int(b)) for a, b in zip(ai, aj) if cdist[a, b] < junc_match_threshold]


def line_visibility(lines3d, pose, K3, gt_lines, mindis_th=25.0):
    """One view of visibility_checking (neat-final-parsing.py:314-335).  lines3d [L,2,3], pose [4,4], K3 [3,3],
    gt_lines [G,4] -> (visible [L] bool, mindis [L])."""
    from oracle import neat_oracle as O
    proj_mat = pose.inverse()[:3]                                                  # :320
    R, T = proj_mat[:, :3], proj_mat[:, 3:]
    lines2d_all = O.project2d(K3, R, T, lines3d).reshape(-1, 4)                    # :324
    dis1 = torch.sum((lines2d_all[:, None] - gt_lines[None][:, :, [0, 1, 2, 3]]) ** 2, dim=-1)   # :329
    dis2 = torch.sum((lines2d_all[:, None] - gt_lines[None][:, :, [2, 3, 0, 1]]) ** 2, dim=-1)   # :330
    dis = torch.min(dis1, dis2)
    mindis, _ = dis.min(dim=1)                                                    # :333
    return mindis < mindis_th, mindis


def refine_global_junctions(gj, get_outputs, get_sdf_vals, sdf_threshold=0.05):
    """neat-final-parsing.py:171-184.  get_outputs(x) -> (sdf [J,1], feat, grad [J,3]); get_sdf_vals(x) -> [J,1]."""
    glj_sdf, _, glj_grad = get_outputs(gj)
    gj = (gj - glj_sdf * glj_grad).detach()
    glj_sdf = get_sdf_vals(gj).flatten()
    argsort = torch.argsort(glj_sdf)
    gj, glj_sdf = gj[argsort], glj_sdf[argsort]
    return gj, glj_sdf, glj_sdf.abs() < sdf_threshold


def wireframe_from_lines_and_junctions(lines, junctions, rel_matching_distance_threshold=0.01):
    """get_wireframe_from_lines_and_junctions (:128-157): lines [N,2,3], junctions [J,3] -> (graph [J,J], lines3d_wf)."""
    ep1, ep2 = lines[:, 0], lines[:, 1]
    cost1, cost2 = torch.cdist(ep1, junctions), torch.cdist(ep2, junctions)
    mcost1, midx1 = cost1.min(dim=1)
    mcost2, midx2 = cost2.min(dim=1)
    is_matched = torch.max(mcost1, mcost2) < torch.norm(ep1 - ep2, dim=-1)
    if rel_matching_distance_threshold > 0:
        is_matched = is_matched * (is_matched < rel_matching_distance_threshold)   # :140 as written: clears every match
    graph = torch.zeros((junctions.shape[0], junctions.shape[0]))
    if is_matched.sum() > 0:
        pair = torch.stack([torch.min(midx1, midx2), torch.max(midx1, midx2)], dim=1)[is_matched]
        graph[pair[:, 0], pair[:, 1]] = 1
        graph[pair[:, 1], pair[:, 0]] = 1
    return graph, junctions[graph.triu().nonzero()]


def initial_recon(views, global_junctions, line_dis_threshold=10, line_score_threshold=0.01, junc_match_threshold=0.05):
    """The assembly of initial_recon (:186-295) from per-view eval outputs.  views: iterable of
    (lines2d [N,4], lines3d [N,2,3], l3d [N,3], gt_lines [G,4]); global_junctions: the (refined) [J,3]."""
    from collections import defaultdict
    gjc, lines3d_all, scores_all = defaultdict(list), [], []
    for lines2d, lines3d, l3d, gt_lines in views:
        _, l3, sc, _ = vote_lines(lines2d, lines3d, l3d, gt_lines, line_dis_threshold)
        if l3.shape[0] > 0:
            endpoints = l3.reshape(-1, 3)
            for ai, aj in match_endpoints(global_junctions, l3, junc_match_threshold):
                gjc[ai].append(endpoints[aj])
            lines3d_all.append(l3)
            scores_all.append(sc)
    lines3d_all, scores_all = torch.cat(lines3d_all, dim=0), torch.cat(scores_all, dim=0)
    lines3d_all = lines3d_all[scores_all < line_score_threshold]
    junctions3d_initial = torch.stack([global_junctions[k] for k, v in gjc.items() if len(v) > 1])
    graph_initial, lines3d_wfi = wireframe_from_lines_and_junctions(lines3d_all, junctions3d_initial, 0)
    return {"junctions3d_initial": junctions3d_initial, "lines3d_all": lines3d_all, "graph_initial": graph_initial,
            "lines3d_wfi": lines3d_wfi}


def visibility_checking(lines3d_all, views, mindis_th=25, min_visible_views=1):
    """:305-337.  views: iterable of (pose [4,4], K3 [3,3], gt_lines [G,4])."""
    vis = torch.stack([line_visibility(lines3d_all, pose, K3, gt, mindis_th)[0] for pose, K3, gt in views], dim=1)
    return lines3d_all[vis.sum(dim=1) >= min_visible_views]
