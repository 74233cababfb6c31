"""CPU: the finalisation oracle (oracle/parsing_oracle.py) against a brute-force evaluation of the same definition."""
import numpy as np
import torch

from oracle import parsing_oracle as PO


def test_vote_lines_oracle_matches_bruteforce():
    g = torch.Generator().manual_seed(0)
    N, Gn = 200, 17
    gt = torch.rand(Gn, 4, generator=g) * 100
    pick = torch.randint(0, Gn, (N,), generator=g)
    l2 = gt[pick] + torch.randn(N, 4, generator=g)
    l2[::3] = l2[::3][:, [2, 3, 0, 1]]
    l2[::7] += 80
    l3 = torch.randn(N, 2, 3, generator=g)
    p3 = torch.randn(N, 3, generator=g)
    labels, mean, scores, counts = PO.vote_lines(l2, l3, p3, gt, 10.0)
    acc = {}
    for i in range(N):
        for sw in (False, True):
            a = l2[i][[2, 3, 0, 1]] if sw else l2[i]
            d = ((a[None] - gt) ** 2).sum(-1)
            k = int(d.argmin())
            if float(d[k]) < 10.0:
                acc.setdefault(k, []).append((l3[i][[1, 0]] if sw else l3[i], p3[i]))
    assert sorted(acc) == labels.tolist()
    for j, k in enumerate(labels.tolist()):
        m = torch.stack([v[0] for v in acc[k]]).mean(0)
        assert torch.allclose(m, mean[j], atol=1e-6)
        assert counts[j] == len(acc[k])
        dis = [float(torch.linalg.norm(torch.linalg.cross(p - m[0], p - m[1])) / torch.linalg.norm(m[1] - m[0]).clamp_min(1e-6))
               for _, p in acc[k]]
        if len(acc[k]) == 3:      # reference quirk: torch.cross without dim runs along dimension 0 of a [3,3] input
            pts = torch.stack([p for _, p in acc[k]])
            dis = (torch.linalg.norm(torch.cross(pts - m[:1], pts - m[1:], dim=0), dim=-1) /
                   torch.linalg.norm(m[1] - m[0]).clamp_min(1e-6)).tolist()
        assert abs(np.mean(dis) - float(scores[j])) < 1e-5


def test_match_endpoints_oracle():
    gj = torch.tensor([[0.0, 0, 0], [1, 0, 0], [5, 5, 5]])
    lines = torch.tensor([[[0.01, 0, 0], [1.0, 0.02, 0]]])
    assert PO.match_endpoints(gj, lines, 0.05) == [(0, 0), (1, 1)]


# ---------------------------------------------------------------------------- pinned against the unmodified reference
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "parsing_synth.npz")


def fake_sdf(x):                                  # the analytic stand-ins of oracle/make_golden_parsing.py
    return 0.02 * torch.sin(7.0 * x[:, :1])


def fake_grad(x):
    g = x + 0.1
    return g / g.norm(dim=-1, keepdim=True)


def load_scene():
    """The synthetic finalisation scene of the golden file: per view (pose, gt_lines [G,5], mask_idx, labels, outputs)."""
    g = np.load(GOLD)
    t = torch.from_numpy
    edges, weights = t(g["edges"]), t(g["weights"])
    views, v = [], 0
    while "v%d_pose" % v in g:
        verts = t(g["v%d_verts2d" % v])
        gt5 = torch.cat((verts[edges[:, 0]], verts[edges[:, 1]], weights[:, None]), dim=-1)
        views.append({"pose": t(g["v%d_pose" % v]), "verts2d": verts, "gt5": gt5, "mask_idx": t(g["v%d_mask_idx" % v]),
                      "labels": t(g["v%d_labels" % v]), "lines3d": t(g["v%d_lines3d" % v]), "lines2d": t(g["v%d_lines2d" % v]),
                      "l3d": t(g["v%d_l3d" % v])})
        v += 1
    return g, views, edges, weights


def test_initial_recon_and_visibility_oracle_vs_reference():
    g, views, _, _ = load_scene()
    gj, _, _ = PO.refine_global_junctions(torch.from_numpy(g["gj"]), lambda x: (fake_sdf(x), None, fake_grad(x)), fake_sdf)
    res = PO.initial_recon([(d["lines2d"].reshape(-1, 4), d["lines3d"], d["l3d"], d["gt5"][:, :4]) for d in views], gj,
                           line_dis_threshold=10, line_score_threshold=0.01, junc_match_threshold=0.05)
    assert np.array_equal(res["junctions3d_initial"].numpy(), g["r_junctions3d_initial"])
    assert np.allclose(res["lines3d_all"].numpy(), g["r_lines3d_all"], rtol=0, atol=1e-7)
    assert np.array_equal(res["graph_initial"].numpy().astype(np.uint8), g["r_graph_initial"])
    assert np.array_equal(res["lines3d_wfi"].numpy(), g["r_lines3d_wfi"])
    K3 = torch.from_numpy(g["K"])[:3, :3]
    cams = [(d["pose"], K3, d["gt5"][:, :4]) for d in views]
    for nv, th in ((1, 25.0), (2, 8.0), (4, 4.0)):
        got = PO.visibility_checking(res["lines3d_wfi"], cams, mindis_th=th, min_visible_views=nv)
        assert np.array_equal(got.numpy(), g["r_checked_%d_%g" % (nv, th)]), (nv, th)


def test_wireframe_graph_oracle_threshold_quirk_and_self_edges():
    """rel_matching_distance_threshold > 0 clears every match (:140 as written); a line whose end points both snap to
    the same junction gives a diagonal entry and a degenerate wireframe line (triu keeps the diagonal)."""
    J = torch.tensor([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]])
    lines = torch.tensor([[[0.02, 0, 0], [0.98, 0.01, 0]], [[0.0, 0.9, 0], [0.9, 0.05, 0]], [[0.1, 0, 0], [-0.1, 0.3, 0]]])
    graph, wf = PO.wireframe_from_lines_and_junctions(lines, J, 0)
    assert graph.tolist() == [[1, 1, 0], [1, 0, 1], [0, 1, 0]]
    assert wf.shape == (3, 2, 3) and torch.equal(wf[0, 0], wf[0, 1])
    graph, wf = PO.wireframe_from_lines_and_junctions(lines, J, 0.01)
    assert float(graph.sum()) == 0 and wf.shape[0] == 0


# ------------------------------------------------------------------------------------------------------------ GPU
import pytest


class _ReplayModel:
    """CUDA twin of oracle/make_golden_parsing.py's stand-in: replays the stored eval outputs chunk by chunk."""

    def __init__(self, g, views):
        self.views, self.view, self.cursor = views, 0, 0
        self.latents = torch.zeros(1, device="cuda")
        gj = torch.from_numpy(g["gj"]).cuda()
        self.ffn = lambda _: gj.clone()
        outer = self

        class Implicit:
            def get_outputs(self, x):
                return fake_sdf(x), None, fake_grad(x)

            def get_sdf_vals(self, x):
                return fake_sdf(x)

        self.implicit_network = Implicit()

    def eval(self):
        return self

    def __call__(self, s):
        n = s["uv"].shape[1]
        d = self.views[self.view]
        a, self.cursor = self.cursor, self.cursor + n
        return {k: d[k][a:a + n].cuda() for k in ("lines3d", "lines2d", "l3d")}


class _Loader:
    def __init__(self, g, views, edges, weights, model):
        from neat_b200 import trainer as TR
        H, W = int(g["H"]), int(g["W"])
        pix = torch.arange(H * W)
        uv = torch.stack((pix % W, pix // W), dim=1).float()
        self.items, self.model = [], model
        for v, d in enumerate(views):
            mask = torch.zeros(H * W, dtype=torch.bool)
            mask[d["mask_idx"]] = True
            mi = {"mask": mask[None], "intrinsics": torch.from_numpy(g["K"])[None], "uv": uv[None], "uv_proj": uv[None],
                  "pose": d["pose"][None], "wireframe": [TR.Wireframe(d["verts2d"], edges, weights)]}
            self.items.append((torch.LongTensor([v]), mi, {}))

    def __len__(self):
        return len(self.items)

    def __iter__(self):
        for v, it in enumerate(self.items):
            self.model.view, self.model.cursor = v, 0
            yield it


@pytest.mark.gpu
def test_gpu_finalisation_vs_reference():
    """neat_b200.parsing.initial_recon / visibility_checking (CUDA kernels + native assignment) against the results of
    the unmodified reference functions on the same scene."""
    from neat_b200 import parsing as P
    g, views, edges, weights = load_scene()
    model = _ReplayModel(g, views)
    loader = _Loader(g, views, edges, weights, model)
    res = P.initial_recon(model, loader, int(g["chunk"]), line_dis_threshold=10, line_score_threshold=0.01,
                          junc_match_threshold=0.05, sdf_junction_refine=True)
    assert np.allclose(res["junctions3d_initial"].cpu().numpy(), g["r_junctions3d_initial"], rtol=0, atol=1e-6)
    assert res["lines3d_all"].shape == g["r_lines3d_all"].shape
    assert np.allclose(res["lines3d_all"].cpu().numpy(), g["r_lines3d_all"], rtol=0, atol=1e-6)
    assert np.array_equal(res["graph_initial"].cpu().numpy().astype(np.uint8), g["r_graph_initial"])
    assert np.allclose(res["lines3d_wfi"].cpu().numpy(), g["r_lines3d_wfi"], rtol=0, atol=1e-6)
    wfi = torch.from_numpy(g["r_lines3d_wfi"]).cuda()
    for nv, th in ((1, 25.0), (2, 8.0), (4, 4.0)):
        got = P.visibility_checking(wfi, loader, model, mindis_th=th, min_visible_views=nv)
        assert np.array_equal(got.cpu().numpy(), g["r_checked_%d_%g" % (nv, th)]), (nv, th)


@pytest.mark.gpu
@pytest.mark.parametrize("N,J,thr", [(300, 40, 0.0), (2000, 1500, 0.0), (1, 1, 0.0), (50, 7, 0.01), (0, 5, 0.0)])
def test_gpu_wireframe_graph_vs_oracle(N, J, thr):
    """incl. more junctions than one shared-memory tile, a single junction (self edge), the positive-threshold quirk
    and an empty line set."""
    from neat_b200 import parsing as P
    gen = torch.Generator().manual_seed(N + J)
    jn = torch.rand(J, 3, generator=gen)
    a, b = torch.randint(0, J, (N,), generator=gen), torch.randint(0, J, (N,), generator=gen)
    lines = torch.stack((jn[a], jn[b]), dim=1) + 0.002 * torch.randn(N, 2, 3, generator=gen)
    lines[::5] += 0.3 * torch.randn(lines[::5].shape, generator=gen)          # some far from every junction
    graph, wf = P.wireframe_from_lines_and_junctions(lines.cuda(), jn.cuda(), thr)
    og, owf = PO.wireframe_from_lines_and_junctions(lines, jn, thr)
    assert torch.equal(graph.cpu(), og) and torch.equal(wf.cpu(), owf)
