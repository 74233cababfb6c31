"""TEST INFRASTRUCTURE ONLY -- golden vectors for the wireframe finalisation from the UNMODIFIED reference functions
`initial_recon`, `get_wireframe_from_lines_and_junctions` and `visibility_checking` (code/neat-final-parsing.py:128-337),
run in the build container on a synthetic scene: the model is a stand-in that replays stored per-pixel outputs
(lines3d / lines2d / l3d, i.e. what the eval forward returns) and an analytic SDF for the junction refinement, the
dataloader a list of hand-built items.  Everything the functions compute from those outputs -- voting, scoring,
end-point / junction assignment, graph construction, visibility -- is the reference's own code.
    python oracle/make_golden_parsing.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from neat_b200 import synth  # noqa: E402  (pure-numpy camera helper)
from oracle import ref_shim  # noqa: E402

H, W, FOCAL = 120, 160, 200.0
N_VIEWS, N_PIX, CHUNK = 4, 500, 192


def fake_sdf(x):
    return 0.02 * torch.sin(7.0 * x[:, :1])


def fake_grad(x):
    g = x + 0.1
    return g / g.norm(dim=-1, keepdim=True)


def build_scene(seed=3):
    rs = np.random.RandomState(seed)
    J3 = rs.uniform(-0.4, 0.4, size=(10, 3)).astype(np.float32)
    pairs = [(a, b) for a in range(10) for b in range(a + 1, 10)]
    edges = np.asarray([pairs[i] for i in rs.permutation(len(pairs))[:14]], dtype=np.int64)
    weights = rs.uniform(0.3, 1.0, len(edges)).astype(np.float32)
    K = np.eye(4, dtype=np.float32)
    K[0, 0] = K[1, 1] = FOCAL
    K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    views = []
    for v in range(N_VIEWS):
        a = 0.4 + 1.5 * v
        pose = np.asarray(synth.look_at_pose((2.0 * np.cos(a), 2.0 * np.sin(a), 0.7)), dtype=np.float32)
        pinv = np.linalg.inv(pose.astype(np.float64))

        def project(X):
            x = (K[:3, :3].astype(np.float64) @ (pinv[:3, :3] @ X.reshape(-1, 3).T.astype(np.float64) + pinv[:3, 3:])).T
            return (x[:, :2] / x[:, 2:]).astype(np.float32)

        verts2d = project(J3)
        mask_idx = np.sort(rs.permutation(H * W)[:N_PIX])
        labels = rs.randint(0, len(edges), H * W).astype(np.int64)
        e = labels[mask_idx]
        L3 = J3[edges[e]].copy()                                   # [N,2,3]
        swap = rs.rand(N_PIX) < 0.5
        L3[swap] = L3[swap][:, ::-1]
        L3 += rs.normal(0, 0.004, L3.shape).astype(np.float32)
        l2 = project(L3).reshape(N_PIX, 2, 2) + rs.normal(0, 0.8, (N_PIX, 2, 2)).astype(np.float32)
        out = rs.rand(N_PIX) < 0.1
        l2[out] += rs.normal(0, 15.0, (int(out.sum()), 2, 2)).astype(np.float32)
        t = rs.rand(N_PIX, 1).astype(np.float32)
        l3d = L3[:, 0] + t * (L3[:, 1] - L3[:, 0]) + rs.normal(0, 0.003, (N_PIX, 3)).astype(np.float32)
        views.append(dict(pose=pose, verts2d=verts2d, mask_idx=mask_idx, labels=labels, lines3d=L3.astype(np.float32),
                          lines2d=l2.astype(np.float32), l3d=l3d.astype(np.float32)))
    gj = np.concatenate([J3 + rs.normal(0, 0.01, J3.shape), rs.uniform(-0.5, 0.5, (6, 3))]).astype(np.float32)
    gj = gj[rs.permutation(len(gj))]
    return dict(J3=J3, edges=edges, weights=weights, K=K, views=views, gj=gj)


class ReplayModel:
    """Stand-in for VolSDFNetwork in eval mode: replays the stored outputs chunk by chunk, in call order."""

    def __init__(self, scene, project2D):
        self.scene, self.cursor, self.view = scene, 0, 0
        self.latents = torch.zeros(1)
        self.ffn = lambda _: torch.from_numpy(scene["gj"]).clone()
        self.project2D = project2D
        outer = self

        class Implicit:
            def get_outputs(self, x):
                return fake_sdf(x), None, fake_grad(x)

            def get_sdf_vals(self, x):
                return fake_sdf(x)

        self.implicit_network = Implicit()

    def eval(self):
        return self

    def start_view(self, v):
        self.view, self.cursor = v, 0

    def __call__(self, s):
        n = s["uv"].shape[1]
        d = self.scene["views"][self.view]
        a, b = self.cursor, self.cursor + n
        self.cursor = b
        t = lambda k: torch.from_numpy(d[k][a:b])
        return {"lines3d": t("lines3d"), "lines2d": t("lines2d"), "l3d": t("l3d")}


def loader_items(scene, model):
    """What the DataLoader + collate_fn of scene_hawp_dataset.py:148-214 yields, built by hand (batch of one)."""
    uv = np.stack([np.arange(H * W) % W, np.arange(H * W) // W], 1).astype(np.float32)
    for v, d in enumerate(scene["views"]):
        wf = ref_shim.Wireframe(d["verts2d"], scene["edges"], scene["weights"])
        mask = np.zeros(H * W, dtype=bool)
        mask[d["mask_idx"]] = True
        lines = wf.line_segments(0.0)[torch.from_numpy(d["labels"])]
        mi = {"mask": torch.from_numpy(mask)[None], "intrinsics": torch.from_numpy(scene["K"])[None],
              "uv": torch.from_numpy(uv)[None], "uv_proj": torch.from_numpy(uv)[None], "lines": lines[None],
              "labels": torch.from_numpy(d["labels"])[None], "pose": torch.from_numpy(d["pose"])[None], "wireframe": [wf]}

        yield v, (torch.LongTensor([v]), mi, {})


class Loader(list):
    """A list of items that tells the replay model which view is being consumed."""

    def __init__(self, scene, model):
        super().__init__(it for _, it in loader_items(scene, model))
        self.model = model

    def __iter__(self):
        for v, it in enumerate(list.__iter__(self)):
            self.model.start_view(v)
            yield it


def main():
    ref_shim.install()
    RefNet = ref_shim.load_classes()[0]
    spec = importlib.util.spec_from_file_location("ref_final_parsing", os.path.join(ref_shim.REF_CODE, "neat-final-parsing.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.tqdm = lambda x: x
    scene = build_scene()
    model = ReplayModel(scene, lambda K, R, T, p: RefNet.project2D(None, K, R, T, p))
    loader = Loader(scene, model)
    torch.manual_seed(0)
    res = mod.initial_recon(model, loader, CHUNK, line_dis_threshold=10, line_score_threshold=0.01, junc_match_threshold=0.05,
                            sdf_junction_refine=True)
    checked = {}
    for views, th in ((1, 25.0), (2, 8.0), (4, 4.0)):
        checked[(views, th)] = mod.visibility_checking(res["lines3d_wfi"], loader, model, mindis_th=th, min_visible_views=views,
                                                       device="cpu")
    gold = {"H": np.array(H), "W": np.array(W), "chunk": np.array(CHUNK), "J3": scene["J3"], "edges": scene["edges"],
            "weights": scene["weights"], "K": scene["K"], "gj": scene["gj"]}
    for v, d in enumerate(scene["views"]):
        for k, a in d.items():
            gold["v%d_%s" % (v, k)] = a
    gold["r_junctions3d_initial"] = res["junctions3d_initial"].numpy()
    gold["r_lines3d_all"] = res["lines3d_all"].numpy()
    gold["r_graph_initial"] = res["graph_initial"].numpy().astype(np.uint8)
    gold["r_lines3d_wfi"] = res["lines3d_wfi"].numpy()
    for (views, th), val in checked.items():
        gold["r_checked_%d_%g" % (views, th)] = val.numpy()
    out = os.path.join(ROOT, "tests", "golden", "parsing_synth.npz")
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out))
    print({k: v.shape for k, v in gold.items() if k.startswith("r_")})


if __name__ == "__main__":
    main()
