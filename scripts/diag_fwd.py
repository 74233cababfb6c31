"""Diagnostic: forward error budget vs the float64 oracle: SDF query, get_outputs (sdf / normal / features), heads,
compositing weights at identical inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_io as G, parity_util as PU
from neat_b200 import synth
from oracle import neat_oracle as O
beta = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
conf = synth.dtu_conf()
sd_np = synth.make_state_dict(conf, seed=5, perturb=0.15, beta=beta)
model = PU.make_model(conf, sd_np).eval()
rn = model._sync_weights()
P64, _ = G.oracle_params(conf, sd_np, dtype=torch.float64)
P32, _ = G.oracle_params(conf, sd_np, dtype=torch.float32)
g = torch.Generator().manual_seed(0)
x = (torch.rand(20000, 3, generator=g) * 2 - 1) * 1.2
d = torch.nn.functional.normalize(torch.randn(20000, 3, generator=g), dim=1)
def e(a, ref):
    a, ref = a.double().cpu(), ref.double()
    return "abs %.2e rel(max) %.2e" % (float((a - ref).abs().max()), float((a - ref).abs().max() / ref.abs().max()))
s64 = O.sdf_vals(P64, x.double())[:, 0]
print("sdf_points      gpu:", e(rn.ctx.sdf_points(x.cuda()), s64), "| f32 oracle:", e(O.sdf_vals(P32, x)[:, 0], s64))
sdf64, feat64, grad64, _ = O.sdf_outputs(P64, x.double())
sdf32, feat32, grad32, _ = O.sdf_outputs(P32, x)
sdf, grad, _, feat, _ = rn.sdf_outputs(rn.explicit_points(x.cuda()), x.shape[0])
print("get_outputs sdf gpu:", e(sdf, sdf64.reshape(-1)), "| f32:", e(sdf32.reshape(-1), sdf64.reshape(-1)))
print("           grad gpu:", e(grad, grad64), "| f32:", e(grad32, grad64))
print("           feat gpu:", e(rn.unpack_features(feat, x.shape[0]), feat64), "| f32:", e(feat32, feat64))
rgb64 = O.rendering_forward(P64, x.double(), grad64, d.double(), feat64)
rgb, _ = rn.head_forward(0, rn.explicit_points(x.cuda(), d.cuda()), x.shape[0], grad64.float().cuda().contiguous(),
                         rn.pack_features(feat64.float().cuda()))
print("rendering head  gpu:", e(rgb, rgb64), "| f32:", e(O.rendering_forward(P32, x, grad64.float(), d, feat64.float()), rgb64))
l64 = O.attraction_forward(P64, x.double(), grad64, d.double(), feat64).reshape(-1, 6)
l3, _ = rn.head_forward(1, rn.explicit_points(x.cuda(), d.cuda()), x.shape[0], grad64.float().cuda().contiguous(),
                        rn.pack_features(feat64.float().cuda()))
print("attraction head gpu:", e(l3, l64), "| f32:", e(O.attraction_forward(P32, x, grad64.float(), d, feat64.float()).reshape(-1, 6), l64))
# compositing at identical sdf: rays through the surface
R, S = 512, 98
z = torch.sort(torch.rand(R, S, generator=g) * 4 + 0.5, dim=1).values
sd = (torch.rand(R, 1, generator=g) * 2 + 1.5) - z + 0.02 * torch.randn(R, S, generator=g)   # crosses zero along the ray
w64 = O.volume_weights(z.double(), sd.double(), P64.beta())
w = model.volume_rendering(z.cuda(), sd.reshape(-1, 1).cuda())
print("weights (beta %.3g) gpu:" % beta, e(w, w64), "| f32:", e(O.volume_weights(z, sd, P32.beta()), w64))
print("  depth from weights gpu: %.2e | f32: %.2e" % (float(((w.double().cpu() - w64) * z.double()).sum(1).abs().max()),
      float(((O.volume_weights(z, sd, P32.beta()).double() - w64) * z.double()).sum(1).abs().max())))
