"""GPU bring-up: stage outputs (get_outputs, heads) and the full eval forward vs the goldens/oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from neat_b200 import synth, _lib
from neat_b200.context import Context
from neat_b200.render import Renderer
import golden_io as G
T = lambda a: torch.from_numpy(np.asarray(a)).cuda().contiguous()

def run(name):
    g, conf, sd_np = G.load(name)
    ctx = Context(conf)
    sd = {k: torch.from_numpy(v).cuda() for k, v in sd_np.items()}
    ctx.pack_weights(ctx.flatten_state_dict(sd))
    rn = Renderer(ctx, conf)
    x, d = T(g["stage_points"]), T(g["stage_dirs"])
    M = x.shape[0]
    pts = rn.explicit_points(x, d)
    sdf, grad, _, feat, _ = rn.sdf_outputs(pts, M)
    torch.cuda.synchronize()
    print(name, "stage sdf", G.rel_err(sdf.cpu(), g["stage_sdf"][:, 0]), "grad", G.rel_err(grad.cpu(), g["stage_grad"]), flush=True)
    gr = T(g["stage_grad"])
    rgb, _ = rn.head_forward(0, pts, M, gr, feat)
    l3, _ = rn.head_forward(1, pts, M, gr, feat)
    torch.cuda.synchronize()
    print(name, "stage rgb", G.rel_err(rgb.cpu(), g["stage_rgb"]), "lines3d", G.rel_err(l3.cpu().view(-1, 2, 3), g["stage_lines3d"]), flush=True)
    out = rn.forward_eval(T(g["in_uv"][0]), T(g["in_pose"][0]), T(g["in_intrinsics"][0]), T(g["in_uv_proj"][0]), sd["density.beta"].reshape(1))
    torch.cuda.synchronize()
    for k in ("rgb_values", "depth", "points3d", "lines3d", "lines2d", "lines2d_calib", "l3d", "normal_map"):
        print("   eval", k, G.rel_err(out[k].cpu(), g["eval_" + k]))
    print("   eval sdf abs", float(np.abs(out["sdf"].cpu().numpy() - g["eval_sdf"]).max()), "k", int(out["n_sampler_iters"].item()))

for name in ("toy_beta0.1", "dtu_beta0.1", "dtu_beta0.01"):
    run(name)
