"""Workload for compute-sanitizer (racecheck / synccheck / memcheck): one small DTU training step in which every persistent
CTA runs SEVERAL tiles (grid capped to 4 CTAs, 33 rays = 26 tiles -> 6-7 tiles per CTA: the mbarrier phases wrap), the
weight-gradient GEMMs are split, and an eval forward follows.
    compute-sanitizer --tool racecheck python scripts/sanitize_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from neat_b200 import synth
import parity_util as PU

conf = synth.dtu_conf()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 33
sd_np = synth.make_state_dict(conf, seed=6, perturb=0.15, beta=0.1)
model = PU.make_model(conf, sd_np)
rn = model._get_renderer()
rn.ctx.debug_grid_cap(4)
rn.ctx.debug_wgrad_split(4, 2)
b = synth.make_batch(R, seed=8)
out, lo = PU.gpu_step(model, b)
print("train step: loss", float(lo["loss"]), "k", int(model.last_step.n_iters.item()))
model.eval()
inp, _ = PU.device_inputs(b)
with torch.no_grad():
    o = model(inp)
torch.cuda.synchronize()
print("eval forward: rgb mean", float(o["rgb_values"].mean()))
