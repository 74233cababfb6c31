"""EVIDENCE SCRIPT (checker side): the FIRST training step on the benchmark's synthetic DTU batch (1024 rays), the plugin with
model.rng = "reference" (the reference's own CPU-generator draws under the same seed) against the unmodified reference on
the same GPU: every loss term should agree, since weights, batch and draws are identical."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from neat_b200 import synth
from neat_b200 import trainer as TR

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
hb = TR.host_batch(R, seed=1)
inp, gt = TR.to_device(hb, dev)

ps = TR.TrainStep(synth.dtu_conf(), device=dev, seed=42, beta=0.1, rng="reference")   # torch.manual_seed(42) + constructor
out = ps.model(inp)
lo = ps.loss_fn(out, gt)
ours = {k: float(v) for k, v in lo.items() if torch.is_tensor(v) and v.numel() == 1}
ours["n_local"] = int(out["j3d_local"].shape[0])
print("ours     ", {k: round(v, 6) for k, v in ours.items()})

from oracle import ref_bench, ref_shim
ref_shim.install(); ref_shim.force_cpu(False)
model, loss_fn = ref_bench._build(synth.dtu_conf(), 42, 0.1, "cuda:0")
model.train()
wf = ref_shim.Wireframe(hb["wireframe"][0].vertices.numpy(), hb["wireframe"][0].edges.numpy(), hb["wireframe"][0].weights.numpy())
mi = {k: hb[k].to(dev) for k in ("intrinsics", "uv", "pose", "uv_proj")}
mi["wireframe"] = [wf]
o2 = model(mi)
l2 = loss_fn(o2, {"rgb": hb["rgb"], "lines2d": hb["lines2d"]})
ref = {k: float(v) for k, v in l2.items() if torch.is_tensor(v) and v.numel() == 1}
ref["n_local"] = int(o2["j3d_local"].shape[0])
print("reference", {k: round(v, 6) for k, v in ref.items()})
print("max |d| over shared loss terms:", max(abs(ours[k] - ref[k]) for k in ours if k in ref))
for k in ("rgb_values", "lines3d", "lines2d_calib", "grad_theta"):
    a, b = out[k].detach().float().cpu(), o2[k].detach().float().cpu()
    print(k, "rel err %.2e" % float((a - b).abs().max() / b.abs().max()))

pa, pb = out["points"].detach().float().cpu(), o2["points"].detach().float().cpu()
dz = (pa - pb).norm(dim=-1)                       # [R,S] distance between corresponding sample points
print("rays whose samples differ by > 1e-4 somewhere: %d of %d; samples moved: %.3f %%" % (
    int((dz.max(dim=1).values > 1e-4).sum()), dz.shape[0], 100.0 * float((dz > 1e-4).float().mean())))
# ---- parameter gradients of that step, both implementations (sampler decisions of a few rays differ, see grad_theta)
ps.bucket.zero()
lo["loss"].backward()
model.zero_grad()
l2["loss"].backward()
torch.cuda.synchronize()
mine = dict(ps.model.named_parameters())
worst = {}
for n, p in model.named_parameters():
    if p.grad is None or mine[n].grad is None:
        continue
    a, b = mine[n].grad.detach().double().cpu(), p.grad.detach().double().cpu()
    if float(b.norm()) == 0.0:
        continue
    net = n.split(".")[0]
    l2e, mxe = float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())
    w = worst.setdefault(net, [0.0, 0.0, ""])
    if l2e > w[0]:
        w[0], w[2] = l2e, n
    w[1] = max(w[1], mxe)
for net, (l2e, mxe, n) in worst.items():
    print("grad vs reference  %-20s worst rel-L2 %.2e (%s)  worst max-rel %.2e" % (net, l2e, n, mxe))
