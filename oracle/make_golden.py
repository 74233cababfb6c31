"""TEST INFRASTRUCTURE ONLY -- generate golden vectors from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/*.npz

For each config (toy = BASELINE.json configs[0]; dtu = dtu.conf nets at a small ray count)
the reference ``VolSDFNetwork`` / ``VolSDFLoss`` are imported through ``oracle/ref_shim.py``,
loaded with the deterministic synthetic state dict of ``neat_b200.synth`` and run
  (a) in eval mode (bit-deterministic, no RNG), and
  (b) in training mode with forward + loss + backward, recording every CPU-generator draw
       so the oracle / CUDA path can replay them.
Weights are NOT stored (they are regenerated from the seed); inputs, the recorded randoms,
outputs, loss terms and parameter gradients are.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from neat_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class DrawRecorder:
    """Wraps torch.rand / randint / randperm / Tensor.uniform_ to record what the reference draws."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._rand, self._randint, self._randperm = torch.rand, torch.randint, torch.randperm
        self._uniform = torch.Tensor.uniform_
        rec = self

        def rand(*a, **k):
            t = rec._rand(*a, **k)
            rec.draws.append(("rand", t.clone()))
            return t

        def randint(*a, **k):
            t = rec._randint(*a, **k)
            rec.draws.append(("randint", t.clone()))
            return t

        def randperm(*a, **k):
            t = rec._randperm(*a, **k)
            rec.draws.append(("randperm", t.clone()))
            return t

        def uniform_(self_, *a, **k):
            t = rec._uniform(self_, *a, **k)
            rec.draws.append(("uniform_", t.clone()))
            return t

        torch.rand, torch.randint, torch.randperm = rand, randint, randperm
        torch.Tensor.uniform_ = uniform_
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randint, torch.randperm = self._rand, self._randint, self._randperm
        torch.Tensor.uniform_ = self._uniform


def build_reference(conf, sd_np):
    Net, Loss, _, _ = ref_shim.load_classes()
    model = Net(conf=ref_shim.to_config(conf))
    sd = {k: torch.from_numpy(v.copy()) for k, v in sd_np.items()}
    missing = model.load_state_dict(sd, strict=True)
    loss = Loss(**synth.loss_conf())
    return model, loss


def clustered_uv(R, seed, W, H, n_clusters):
    """Pixels in tight groups so that DBSCAN(eps=0.01) finds clusters of attraction endpoints."""
    rs = np.random.RandomState(seed)
    centers = np.stack([rs.uniform(0.3 * W, 0.7 * W, n_clusters), rs.uniform(0.3 * H, 0.7 * H, n_clusters)], -1)
    idx = np.arange(R) % n_clusters
    return (centers[idx] + rs.uniform(-0.25, 0.25, size=(R, 2))).astype(np.float32)


def model_input(batch, wf):
    t = lambda a: torch.from_numpy(np.asarray(a))
    return {"intrinsics": t(batch["intrinsics"]), "uv": t(batch["uv"]), "pose": t(batch["pose"]),
            "uv_proj": t(batch["uv_proj"]), "wireframe": [wf]}


def to_np(v):
    return v.detach().cpu().numpy()


def run_case(name, conf, R, seed_w, beta, cam, perturb=0.15):
    sd_np = synth.make_state_dict(conf, seed=seed_w, perturb=perturb, beta=beta)
    model, loss_fn = build_reference(conf, sd_np)
    H, W, focal, pose = cam
    K = np.eye(4, dtype=np.float32)
    K[0, 0] = K[1, 1] = focal
    K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    batch = synth.make_batch(R, seed=3, img_res=(H, W), K=K, pose=pose, n_junctions=12, n_edges=20)
    batch["uv"] = clustered_uv(R, 5, W, H, max(R // 8, 1))[None]
    batch["uv_proj"] = (batch["uv"] + np.random.RandomState(6).normal(size=batch["uv"].shape)).astype(np.float32)
    wf = ref_shim.Wireframe(batch["wf_vertices"], batch["wf_edges"], batch["wf_weights"])
    gold = {"R": R, "seed_w": seed_w, "beta": beta, "perturb": perturb}
    for k in ("intrinsics", "pose", "uv", "uv_proj", "rgb", "lines2d"):
        gold["in_" + k] = batch[k]

    # ---- (a) eval mode ------------------------------------------------------------
    model.eval()
    out = model(model_input(batch, wf))
    for k in ("rgb_values", "depth", "xyz", "points3d", "lines3d", "lines2d", "lines2d_calib", "l3d",
              "sdf", "normal_map", "points"):
        gold["eval_" + k] = to_np(out[k])
    gold["eval_z_vals"] = to_np((out["points"][:, :, :] - torch.from_numpy(batch["pose"][0, :3, 3]))
                                .norm(dim=-1))
    # per-stage outputs of the submodules on the eval points (kernel-level goldens)
    pts = out["points"].reshape(-1, 3).detach()
    n_pts = min(pts.shape[0], 768)
    sel = torch.from_numpy(np.random.RandomState(7).choice(pts.shape[0], n_pts, replace=False))
    p_sel = pts[sel].clone()
    sdf, feat, grad = model.implicit_network.get_outputs(p_sel)
    dirs = torch.from_numpy(np.random.RandomState(8).normal(size=(n_pts, 3)).astype(np.float32))
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    rgb = model.rendering_network(p_sel, grad, dirs, feat)
    l3 = model.attraction_network(p_sel, grad, dirs, feat)
    gold.update(stage_points=to_np(p_sel), stage_dirs=to_np(dirs), stage_sdf=to_np(sdf),
                stage_feat=to_np(feat), stage_grad=to_np(grad), stage_rgb=to_np(rgb),
                stage_lines3d=to_np(l3),
                stage_sdf_vals=to_np(model.implicit_network.get_sdf_vals(p_sel.detach())))

    # ---- (b) training mode: pass 1 finds cluster projections to place GT junctions ----
    model.train()
    torch.manual_seed(42)
    out1 = model(model_input(batch, wf))
    with torch.no_grad():
        import importlib
        cent = model.cluster_dbscan(out1["lines3d"].detach().cpu().numpy().reshape(-1, 3), eps=0.01, min_samples=2)
        proj = torch.linalg.inv(torch.from_numpy(batch["pose"][0]))[:3]
        j2d = model.project2D(torch.from_numpy(batch["intrinsics"][0, :3, :3]), proj[:, :3], proj[:, 3:], cent)
    rs = np.random.RandomState(9)
    nj = j2d.shape[0]
    keep = rs.permutation(nj)[: max(nj // 2, min(nj, 4))]
    verts = np.concatenate([to_np(j2d)[keep] + rs.normal(scale=1.0, size=(len(keep), 2)),
                            rs.uniform(0, W, size=(3, 2))], 0).astype(np.float32)
    edges = np.stack([np.arange(len(verts)), (np.arange(len(verts)) + 1) % len(verts)], -1)
    wf = ref_shim.Wireframe(verts, edges, np.ones(len(edges), np.float32))
    gold["wf_vertices"] = verts
    gold["n_clusters_pass1"] = nj

    model.zero_grad()
    torch.manual_seed(42)
    with DrawRecorder() as rec:
        out = model(model_input(batch, wf))
    gt = {"rgb": torch.from_numpy(batch["rgb"]), "lines2d": torch.from_numpy(batch["lines2d"])}
    lo = loss_fn(out, gt)
    lo["loss"].backward()
    kinds = [k for k, _ in rec.draws]
    assert kinds == ["rand", "randint", "rand", "randperm", "randint", "uniform_"], kinds
    d = [t for _, t in rec.draws]
    gold.update(rnd_t_rand=to_np(d[0]), rnd_u_final=to_np(d[2]),
                rnd_extra_idx=to_np(d[3][: conf["ray_sampler"]["N_samples_extra"]]),
                rnd_perm_len=int(d[3].shape[0]), rnd_eik_idx=to_np(d[4]), rnd_eik_uniform=to_np(d[5]))
    for k in ("rgb_values", "depth", "xyz", "points3d", "lines3d", "lines2d", "lines2d_calib", "l3d", "sdf",
              "j2d_local", "j3d_local", "j3d_global", "j2d_global", "j2d_local_calib", "j2d_global_calib",
              "grad_theta", "points"):
        gold["train_" + k] = to_np(out[k])
    if "median" in out:
        gold["train_median"] = to_np(out["median"])
    for k, v in lo.items():
        gold["loss_" + k] = np.asarray(to_np(v), dtype=np.float64)
    # parameter gradients: full for small tensors, (sum, abs-sum, 256 sampled entries) for all
    rs = np.random.RandomState(11)
    for n, p in model.named_parameters():
        g = p.grad
        if g is None:
            continue
        g = to_np(g).astype(np.float64).ravel()
        idx = rs.choice(g.size, min(256, g.size), replace=False)
        gold["gstat_" + n] = np.array([g.sum(), np.abs(g).sum(), np.sqrt((g * g).sum())])
        gold["gidx_" + n] = idx
        gold["gval_" + n] = g[idx]
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **gold)
    print(name, "k_train_perm_len", gold["rnd_perm_len"], "n_local", out["j3d_local"].shape[0],
          "loss", float(lo["loss"]), "size", os.path.getsize(path))


def main():
    torch.set_num_threads(8)
    cams = np.load(os.path.join(ref_shim.REF_ROOT, "data/abc/00075213/cameras.npz"))
    abc_pose = cams["extrinsics"][0].astype(np.float32)
    only = sys.argv[1:]
    if "abc" in only or not only:
        # abc-neat-a.conf's model block (no DBSCAN: all end points are junction candidates; median match filter;
        # 64 global junctions), ABC camera 0
        run_case("abc_beta0.1", synth.abc_conf(), 128, seed_w=3, beta=0.1, cam=(512, 512, 560.0, abc_pose))
        if only:
            return
    if "l3d" in only or not only:
        run_case("toy_l3d", synth.toy_l3d_conf(), 128, seed_w=5, beta=0.1, cam=(512, 512, 560.0, abc_pose))
        if only:
            return
    if "white" in only or not only:
        # the model class's optional branches no shipped conf selects: white_bkgd (+ no sphere clamp) and junction_eikonal
        run_case("toy_white_jeik", synth.toy_white_conf(), 128, seed_w=4, beta=0.1, cam=(512, 512, 560.0, abc_pose))
        if only:
            return
    # toy: ABC camera 0 (f=560, 512x512), 256 rays x 64 samples, 4x128 nets
    run_case("toy_beta0.1", synth.toy_conf(), 256, seed_w=0, beta=0.1, cam=(512, 512, 560.0, abc_pose))
    # DTU nets (8x256 / 4x256), 98 samples, small ray count; two density settings (k=2.. and k=5)
    pose = synth.look_at_pose((1.6, 1.5, 1.1))
    run_case("dtu_beta0.1", synth.dtu_conf(), 128, seed_w=1, beta=0.1, cam=(1200, 1600, 2900.0, pose))
    run_case("dtu_beta0.01", synth.dtu_conf(), 128, seed_w=2, beta=0.01, cam=(1200, 1600, 2900.0, pose))


if __name__ == "__main__":
    main()
