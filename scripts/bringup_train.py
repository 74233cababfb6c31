"""GPU bring-up: training forward + loss + hand-written backward vs the goldens of the unmodified reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from neat_b200 import synth
from neat_b200.model import VolSDFNetwork
from neat_b200.loss import VolSDFLoss
import golden_io as G
T = lambda a: torch.from_numpy(np.asarray(a))

class WF:
    def __init__(self, v): self.vertices = torch.as_tensor(v, dtype=torch.float32)

def run(name):
    g, conf, sd_np = G.load(name)
    model = VolSDFNetwork(conf)
    missing = model.load_state_dict({k: T(v.copy()) for k, v in sd_np.items()}, strict=True)
    model = model.cuda().train()
    loss_fn = VolSDFLoss(**synth.loss_conf())
    rnd = G.train_randoms(g)
    model.replay = dict(sampler=dict(t_rand=rnd.sampler.t_rand, u_final=rnd.sampler.u_final, extra_idx=rnd.sampler.extra_idx,
                                     eik_idx=rnd.sampler.eik_idx), eik_uniform=rnd.eik_uniform)
    inp = {"intrinsics": T(g["in_intrinsics"]).cuda(), "uv": T(g["in_uv"]).cuda(), "pose": T(g["in_pose"]).cuda(),
           "uv_proj": T(g["in_uv_proj"]).cuda(), "wireframe": [WF(g["wf_vertices"])]}
    out = model(inp)
    torch.cuda.synchronize()
    for k in ("rgb_values", "lines3d", "lines2d", "lines2d_calib", "l3d", "grad_theta", "j3d_local", "j3d_global", "j2d_local_calib"):
        ok = tuple(out[k].shape) == tuple(g["train_" + k].shape)
        print("  fwd", k, G.rel_err(out[k].detach().cpu(), g["train_" + k]) if ok else ("SHAPE", tuple(out[k].shape), g["train_" + k].shape), flush=True)
    lo = loss_fn(out, {"rgb": T(g["in_rgb"]), "lines2d": T(g["in_lines2d"])})
    for k in ("loss", "rgb_loss", "eikonal_loss", "line_loss", "l2d_loss", "j3d_loss", "j2d_loss"):
        print("  loss", k, float(lo[k]), float(g["loss_" + k]))
    lo["loss"].backward()
    torch.cuda.synchronize()
    worst = []
    for n, p in model.named_parameters():
        if "gstat_" + n not in g or p.grad is None:
            print("  no grad for", n); continue
        gr = p.grad.detach().cpu().numpy().astype(np.float64).ravel()
        ref_norm = g["gstat_" + n][2]
        err = np.abs(gr[g["gidx_" + n]] - g["gval_" + n]).max() / max(ref_norm, 1e-12)
        nerr = abs(np.sqrt((gr * gr).sum()) - ref_norm) / max(ref_norm, 1e-12)
        worst.append((err, nerr, n, ref_norm))
    worst.sort(reverse=True)
    for e in worst[:12]:
        print("  grad %-45s maxerr/norm %.3e  norm relerr %.3e  refnorm %.3e" % (e[2], e[0], e[1], e[3]))
    print(name, "n params compared", len(worst), "worst", worst[0][0], flush=True)

for name in sys.argv[1:] or ("toy_beta0.1", "dtu_beta0.1", "dtu_beta0.01"):
    print("====", name)
    run(name)
