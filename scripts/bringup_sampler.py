"""GPU bring-up: error-bound sampler vs the oracle (eval + training with recorded randoms) + query timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from neat_b200 import synth, _lib
from neat_b200.context import Context, ErrorBoundSampler
from oracle import neat_oracle as O
import golden_io as G
T = lambda a: torch.from_numpy(np.asarray(a))

def run(name):
    g, conf, sd_np = G.load(name)
    P, _ = G.oracle_params(conf, sd_np)
    ctx = Context(conf)
    sd = {k: torch.from_numpy(v).cuda() for k, v in sd_np.items()}
    ctx.pack_weights(ctx.flatten_state_dict(sd))
    smp = ErrorBoundSampler(ctx, conf)
    dirs, cam = O.camera_rays(T(g["in_uv"][0]), T(g["in_pose"][0]), T(g["in_intrinsics"][0]))
    R = dirs.shape[0]
    beta = sd["density.beta"].reshape(1)
    for training in (False, True):
        rnd = G.train_randoms(g).sampler if training else None
        zo, ze, k = O.error_bound_sampler(P, G.sampler_conf(conf), dirs, cam[None].expand(R, 3), training=training, rnd=rnd)
        randoms = dict(t_rand=rnd.t_rand, u_final=rnd.u_final, extra_idx=rnd.extra_idx, eik_idx=rnd.eik_idx) if training else None
        z, zeik, nit = smp.get_z_vals(cam.cuda(), dirs.cuda().contiguous(), beta, training=training, randoms=randoms)
        torch.cuda.synchronize()
        dz = (z.cpu() - zo).abs()
        print(name, "training" if training else "eval", "k oracle", k, "k gpu", int(nit.item()), "max dz", float(dz.max()),
              "frac>1e-4", float((dz > 1e-4).float().mean()), "zeik err", float((zeik.cpu() - ze).abs().max()), flush=True)

for name in ("toy_beta0.1", "dtu_beta0.1", "dtu_beta0.01"):
    run(name)

# timing of the query kernel at sampler scale
conf = synth.dtu_conf()
ctx = Context(conf)
sd = {k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(conf, seed=1).items()}
ctx.pack_weights(ctx.flatten_state_dict(sd))
for R in (1024, 8192):
    x = (torch.rand(R * 128, 3, device="cuda") - 0.5) * 3
    for _ in range(3): ctx.sdf_points(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx.sdf_points(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = R * 128 * 1049088.0
    print("query R=%d: %.3f ms, %.1f Mpts/s, %.1f TFLOP/s algorithmic (x3 bf16 MMAs issued)" % (R, ms, R * 128 / ms / 1e3, fl / ms / 1e9))
