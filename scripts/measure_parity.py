"""Measured parity numbers behind the tolerances in tests/ (run on the GPU box; writes gpurun_out/parity_measured.json,
copied to profiles/ per round).  For every case: worst output / loss-term error and the per-tensor gradient metric of
tests/parity_util.py against autograd through the oracle at identical samples; for the eval forward: the per-ray error
split into rays whose sampler decisions agree with the oracle's and rays where they flipped."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import golden_io as G
import parity_util as PU
from neat_b200 import synth
from oracle import neat_oracle as O

T = PU.T
res = {}


def train_case(tag, conf, R, beta, seed_w=5, seed_b=4, cap=0, split=(0, 0), img=None):
    sd_np = synth.make_state_dict(conf, seed=seed_w, perturb=0.15, beta=beta)
    model = PU.make_model(conf, sd_np)
    rn = model._get_renderer()
    rn.ctx.debug_grid_cap(cap)
    rn.ctx.debug_wgrad_split(*split)
    b = synth.make_batch(R, seed=seed_b, **(img or {}))
    t0 = time.time()
    out, lo = PU.gpu_step(model, b)
    st = model.last_step
    oo, ol, leaves = PU.oracle_step(conf, sd_np, b, st)                         # float64: the ground truth
    o32, l32, leaves32 = PU.oracle_step(conf, sd_np, b, st, dtype=torch.float32)  # what the fp32 reference computes
    ref = {n: v.grad for n, v in leaves.items()}
    table = PU.grad_errors({n: p.grad for n, p in model.named_parameters()}, ref)
    table32 = PU.grad_errors({n: v.grad for n, v in leaves32.items()}, ref)
    res[tag] = {"rays": R, "beta": beta, "sampler_k": int(st.n_iters.item()), "outputs_vs_f64": PU.output_errors(out, oo),
                "f32_oracle_outputs_vs_f64": PU.output_errors(o32, oo),
                "loss_terms_vs_f64": PU.loss_errors(lo, ol), "grad_worst_vs_f64": PU.worst(table),
                "f32_oracle_grad_worst_vs_f64": PU.worst(table32),
                "grad_table_vs_f64": {n: [float("%.3e" % v) for v in t] for n, t in table.items()},
                "f32_oracle_grad_table_vs_f64": {n: [float("%.3e" % v) for v in t] for n, t in table32.items()},
                "seconds": time.time() - t0}
    print(tag, json.dumps({k: res[tag][k] for k in ("sampler_k", "outputs_vs_f64", "f32_oracle_outputs_vs_f64", "loss_terms_vs_f64",
                                                    "grad_worst_vs_f64", "f32_oracle_grad_worst_vs_f64", "seconds")}), flush=True)
    rn.ctx.debug_grid_cap(0)
    rn.ctx.debug_wgrad_split(0, 0)


def eval_case(tag, name=None, conf=None, R=1024, beta=0.01, seed_b=9):
    from neat_b200.context import Context
    from neat_b200.render import Renderer
    if name is not None:
        g, conf, sd_np = G.load(name)
        uv, pose, K, uvp = (T(g["in_uv"][0]), T(g["in_pose"][0]), T(g["in_intrinsics"][0]), T(g["in_uv_proj"][0]))
    else:
        sd_np = synth.make_state_dict(conf, seed=5, perturb=0.15, beta=beta)
        b = synth.make_batch(R, seed=seed_b)
        uv, pose, K, uvp = T(b["uv"][0]), T(b["pose"][0]), T(b["intrinsics"][0]), T(b["uv_proj"][0])
    ctx = Context(conf)
    sd = {k: torch.from_numpy(v).cuda() for k, v in sd_np.items()}
    ctx.pack_weights(ctx.flatten_state_dict(sd))
    rn = Renderer(ctx, conf)
    out = rn.forward_eval(uv.cuda(), pose.cuda(), K.cuda(), uvp.cuda().contiguous(), sd["density.beta"].reshape(1))
    refs = {}
    for dt in (torch.float64, torch.float32):     # float64 = the truth; float32 = what the reference computes
        P, _ = G.oracle_params(conf, sd_np, dtype=dt)
        refs[dt] = O.neat_forward(P, G.sampler_conf(conf), K.to(dt), pose.to(dt), uv.to(dt), uvp.to(dt), training=False)
    ref, r32 = refs[torch.float64], refs[torch.float32]
    n = uv.shape[0]

    def per_ray(a, key):
        want = ref[key].numpy().astype(np.float64)
        return np.abs(np.asarray(a, dtype=np.float64) - want).reshape(n, -1).max(axis=1) / np.abs(want).max()

    def summary(get):
        o = {}
        for key in ("rgb_values", "depth", "points3d", "lines3d", "lines2d", "lines2d_calib", "normal_map", "l3d"):
            e = per_ray(get(key), key)
            q = np.quantile(e, [0.5, 0.9, 0.99, 1.0])
            o[key] = {"median": float(q[0]), "q90": float(q[1]), "q99": float(q[2]), "max": float(q[3]),
                      "rays_over_1e-4": int((e > 1e-4).sum())}
        return o

    dz = (out["z_vals"].cpu().double() - ref["z_vals"]).abs()
    dz32 = (r32["z_vals"].double() - ref["z_vals"]).abs()
    r = {"rays": int(n), "k_gpu": int(out["n_sampler_iters"].item()), "k_oracle": int(ref["n_sampler_iters"]),
         "rays_with_same_samples_1e-4": {"gpu": int((dz.max(1).values <= 1e-4).sum()), "f32_oracle": int((dz32.max(1).values <= 1e-4).sum())},
         "samples_moved_over_2e-4": {"gpu": float((dz > 2e-4).float().mean()), "f32_oracle": float((dz32 > 2e-4).float().mean())},
         "gpu_vs_f64": summary(lambda k: out[k].cpu().numpy()), "f32_oracle_vs_f64": summary(lambda k: r32[k].numpy())}
    res[tag] = r
    print(tag, json.dumps(r), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["all"]
    dtu = synth.dtu_conf()
    if "all" in which or "train" in which:
        train_case("dtu_1024_beta0.1", dtu, 1024, 0.1)
        train_case("dtu_1024_beta0.01", dtu, 1024, 0.01)
        train_case("dtu_130_cap8_split16", dtu, 130, 0.1, seed_w=6, seed_b=8, cap=8, split=(16, 4))
        train_case("toy_256", synth.toy_conf(), 256, 0.1, img=dict(img_res=(512, 512), focal=560.0))
    if "all" in which or "eval" in which:
        eval_case("eval_dtu_1024_beta0.01", conf=dtu)
        for name in ("abc_beta0.1", "dtu_beta0.1", "dtu_beta0.01", "toy_beta0.1"):
            eval_case("eval_golden_" + name, name=name)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "parity_measured.json"), "w"), indent=1)
