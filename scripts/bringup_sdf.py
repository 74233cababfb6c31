"""GPU bring-up: SDF query kernel vs the oracle, for both matrix-descriptor field orders."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from neat_b200 import synth, _lib
from neat_b200.context import Context
from oracle import neat_oracle as O
import golden_io as G

def run(name, swap):
    g, conf, sd_np = G.load(name)
    P, _ = G.oracle_params(conf, sd_np)
    ctx = Context(conf)
    _lib.check(ctx.lib.neat_debug_set_desc_swap(swap))
    sd = {k: torch.from_numpy(v).cuda() for k, v in sd_np.items()}
    ctx.pack_weights(ctx.flatten_state_dict(sd))
    x = torch.from_numpy(g["stage_points"]).cuda()
    torch.cuda.synchronize()
    try:
        s = ctx.sdf_points(x); torch.cuda.synchronize()
    except Exception as e:
        print(name, "swap", swap, "FAILED", e); return
    ref = O.sdf_vals(P, torch.from_numpy(g["stage_points"]))[:, 0].numpy()
    err = np.abs(s.cpu().numpy() - ref)
    print(name, "swap", swap, "max abs err", err.max(), "ref range", ref.min(), ref.max(), "nan", np.isnan(s.cpu().numpy()).sum(), flush=True)
    print("   first 6 got", s[:6].cpu().numpy(), "\n   first 6 ref", ref[:6])

for swap in (0, 1):
    for name in ("toy_beta0.1", "dtu_beta0.1"):
        run(name, swap)
