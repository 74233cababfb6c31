"""TEST INFRASTRUCTURE ONLY (imported by tests/): CPU restatement of the per-image voting block of the reference's
wireframe finalisation, code/neat-final-parsing.py:226-271 (initial_recon).  Parity unpinned by reference outputs: the
block is inline code of a function that needs the dataset classes (hawp CUDA extension at import) and cannot run in the
build container; it is restated statement by statement with the same torch ops, in the same order."""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment


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
        support_dis = torch.norm(torch.cross(support_pts - val[:1], support_pts - val[1:], dim=-1), dim=-1) / \
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
