import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from neat_b200.context import Context
from neat_b200.render import Renderer
from oracle import neat_oracle as O
import golden_io as G
T = lambda a: torch.from_numpy(np.asarray(a)).cuda().contiguous()
name = "toy_beta0.1"
g, conf, sd_np = G.load(name)
P, _ = G.oracle_params(conf, sd_np)
ctx = Context(conf)
sd = {k: torch.from_numpy(v).cuda() for k, v in sd_np.items()}
ctx.pack_weights(ctx.flatten_state_dict(sd))
rn = Renderer(ctx, conf)
out = rn.forward_eval(T(g["in_uv"][0]), T(g["in_pose"][0]), T(g["in_intrinsics"][0]), T(g["in_uv_proj"][0]), sd["density.beta"].reshape(1))
torch.cuda.synchronize()
dirs, cam = O.camera_rays(torch.from_numpy(g["in_uv"][0]), torch.from_numpy(g["in_pose"][0]), torch.from_numpy(g["in_intrinsics"][0]))
R = dirs.shape[0]
z = out["z_vals"].cpu()
print("z vs golden", float((z - torch.from_numpy(g["eval_z_vals"])).abs().max()))
rr = O.render_rays(P, dirs, cam[None].expand(R, 3), z)
for k, kk in (("sdf_pts", "sdf_pts"), ("grad_pts", "grad"), ("rgb_pts", "rgb_pts"), ("lines_pts", "lines3d_pts"), ("weights", "weights"),
              ("rgb_values", "rgb_values"), ("depth", "depth")):
    a, b = out[k].cpu(), rr[kk]
    d = (a - b.reshape(a.shape)).abs()
    print(k, "max abs", float(d.max()), "argmax", int(d.argmax()), "of", d.numel())
d = (out["sdf_pts"].cpu() - rr["sdf_pts"]).abs()
bad = (d > 1e-3).nonzero()
print("bad sdf count", len(bad), bad[:10].tolist())
d = (out["rgb_pts"].cpu() - rr["rgb_pts"]).abs().amax(-1)
bad = (d > 1e-3).nonzero()
print("bad rgb count", len(bad), bad[:10].tolist())
