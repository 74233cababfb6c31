"""BASELINE configs[2] names "bf16 tensor-core MLPs": the opt-in one-MMA-per-MAC mode (neat_set_precision(ctx, 1): hi
planes only, 11-bit fp16 forward operands / 8-bit bf16 backward operands) -- its speed at 8192 rays/step and its measured
distance from the x3 parity mode.    python scripts/fast_mode.py [rays]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from neat_b200 import _lib, synth
from neat_b200 import trainer as TR
import golden_io as G

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
res = {}
outs = {}
for fast in (0, 1):
    ts = TR.FusedTrainStep(synth.dtu_conf(), device=dev, seed=42, beta=0.1)
    inp, gt = TR.to_device(TR.host_batch(R, seed=1), dev)
    rn = ts.model._get_renderer()
    _lib.check(rn.ctx.lib.neat_set_precision(rn.ctx._h, fast))
    ts.model.seed_draws(7)
    lo = ts.step(inp, gt)                       # first step: same weights, same draws in both modes
    torch.cuda.synchronize()
    outs[fast] = {k: ts.out[k].detach().clone() for k in ("rgb_values", "lines3d", "grad_theta")}
    outs[fast]["loss"] = float(lo["loss"])
    outs[fast]["grads"] = ts.bucket.flat.detach().clone() if False else None
    for _ in range(5):
        ts.step(inp, gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        ts.step(inp, gt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res["fast" if fast else "x3"] = {"ms_per_step": round(ms, 3), "rays_per_s": round(R / ms * 1e3), "first_step_loss": outs[fast]["loss"]}
    del ts
    import gc; gc.collect(); torch.cuda.empty_cache()
res["fast_vs_x3_first_step"] = {k: G.rel_err(outs[1][k].cpu(), outs[0][k].cpu()) for k in ("rgb_values", "lines3d", "grad_theta")}
res["rays"] = R
print(json.dumps(res))
