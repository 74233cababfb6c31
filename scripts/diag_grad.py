"""Diagnostic: where does the gradient error of the rendering head come from?  GPU vs the fp32 oracle vs the fp64 oracle
(ground truth) for the rgb-loss path alone, at identical samples; plus the intermediate adjoints."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_io as G, parity_util as PU
from neat_b200 import synth
from oracle import neat_oracle as O
T = PU.T
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
conf = synth.dtu_conf()
sd_np = synth.make_state_dict(conf, seed=5, perturb=0.15, beta=beta)
model = PU.make_model(conf, sd_np)
b = synth.make_batch(R, seed=4)
inp, gt = PU.device_inputs(b)
model.seed_draws(7)
out = model(inp)
st = model.last_step
rgb_gt = gt["rgb"].reshape(-1, 3)
bar = torch.sign(out["rgb_values"].detach() - rgb_gt) / (3 * R)
for p in model.parameters(): p.grad = None
torch.autograd.backward([out["rgb_values"]], [bar])
torch.cuda.synchronize()
gpu = {n: p.grad.detach().cpu().double() for n, p in model.named_parameters() if p.grad is not None}
dbg = {k: v.detach().cpu().double() for k, v in st.debug.items() if k != "feat_bar"}
z, z_eik = PU.step_samples(st)

def oracle(dtype):
    P, leaves = G.oracle_params(conf, sd_np, dtype=dtype, track=True)
    dirs, cam = O.camera_rays(T(b["uv"][0]).to(dtype), T(b["pose"][0]).to(dtype), T(b["intrinsics"][0]).to(dtype))
    rr = O.render_rays(P, dirs, cam[None].expand(R, 3), z.to(dtype))
    rr["rgb_pts"].retain_grad()
    L = (rr["rgb_values"] * bar.cpu().to(dtype)).sum()
    L.backward()
    g = {n: v.grad.detach().double() for n, v in leaves.items() if v.grad is not None}
    inter = dict(rgb_pre_bar=(rr["rgb_pts"].grad * rr["rgb_pts"] * (1 - rr["rgb_pts"])).detach().double().reshape(-1, 3))
    return g, inter, rr

g64, i64, rr64 = oracle(torch.float64)
g32, i32, rr32 = oracle(torch.float32)
def errs(a, ref):
    d = a - ref
    return float(d.norm() / ref.norm().clamp_min(1e-300)), float(d.abs().max() / ref.abs().max().clamp_min(1e-300))
print("forward rgb_values: gpu vs f64 %.2e ; f32 vs f64 %.2e" % (errs(out["rgb_values"].detach().cpu().double(), rr64["rgb_values"].detach())[1],
      errs(rr32["rgb_values"].detach().double(), rr64["rgb_values"].detach())[1]))
for k in ("rgb_pre_bar",):
    a = dbg[k].reshape(i64[k].shape)
    if k == "sdf_bar":
        act = st.act.detach().cpu().double()
        print("  (sdf_bar: GPU value is masked by act)")
        ref64, ref32 = i64[k] * act, i32[k] * act
    else:
        ref64, ref32 = i64[k], i32[k]
    print("%-12s gpu vs f64: l2 %.2e max %.2e | f32 vs f64: l2 %.2e max %.2e" % ((k,) + errs(a, ref64) + errs(ref32, ref64)))
rows = []
for n in sorted(g64):
    if n not in gpu: continue
    e_g, e_o = errs(gpu[n], g64[n]), errs(g32[n], g64[n])
    rows.append((n, e_g, e_o))
rows.sort(key=lambda r: -r[1][1])
print("%-44s %-22s %-22s" % ("tensor", "GPU vs f64 (l2,max)", "f32 oracle vs f64"))
for n, e_g, e_o in rows[:25]:
    print("%-44s %.2e %.2e    %.2e %.2e" % (n, e_g[0], e_g[1], e_o[0], e_o[1]))
