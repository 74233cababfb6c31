"""Is the residual error of the rendering-head gradients the round-toward-zero accumulation of the weight-gradient MMAs?
Same step, different numbers of accumulations per TMEM accumulator (forced tile-range splits).  GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import golden_io as G, parity_util as PU
from neat_b200 import synth
from oracle import neat_oracle as O

def case(tag, conf, R, beta, sw, sb, img=None):
    sd_np = synth.make_state_dict(conf, seed=sw, perturb=0.15, beta=beta)
    model = PU.make_model(conf, sd_np)
    rn = model._get_renderer()
    b = synth.make_batch(R, seed=sb, **(img or {}))
    ref = None
    for split in ((0, 0), (32, 12), (64, 4), (256, 1)):
        rn.ctx.debug_wgrad_split(*split)
        out, lo = PU.gpu_step(model, b)
        if ref is None:
            oo, ol, leaves = PU.oracle_step(conf, sd_np, b, model.last_step)
            ref = {n: v.grad for n, v in leaves.items()}
        t = PU.grad_errors({n: p.grad for n, p in model.named_parameters()}, ref)
        pick = lambda pre: (max(v[0] for n, v in t.items() if n.startswith(pre)), max(v[1] for n, v in t.items() if n.startswith(pre)))
        print("%-14s split %-9s sdf l2 %.2e max %.2e | rend l2 %.2e max %.2e | att l2 %.2e max %.2e" % (
            tag, split, *pick("implicit"), *pick("rendering"), *pick("attraction")), flush=True)
    rn.ctx.debug_wgrad_split(0, 0)

case("dtu_130", synth.dtu_conf(), 130, 0.1, 6, 8)
case("dtu_1024", synth.dtu_conf(), 1024, 0.1, 5, 4)
case("toy_256", synth.toy_conf(), 256, 0.1, 5, 4, dict(img_res=(512, 512), focal=560.0))
