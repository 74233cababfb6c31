"""Measurement-only what-if runs behind DESIGN.md section 7c (neat_debug_set_flags, umma.cuh g_dbg):
  timing : per-kernel CUDA-event times of the 1024-ray training step with parts of the save-record traffic switched off
           (results are wrong with those flags; only the times are read)
  parity : gradient error against the float64 oracle with reduced-precision variants of the backward
    python scripts/whatif.py timing|parity [rays]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from neat_b200 import _lib, synth
from neat_b200 import trainer as TR

lib = _lib.load()
dev = torch.device("cuda", 0)


def set_flags(f):
    torch.cuda.synchronize()
    _lib.check(lib.neat_debug_set_flags(int(f)))


def timing(R, flag_sets, beta=0.1, n=10):
    ps = TR.TrainStep(synth.dtu_conf(), device=dev, seed=42, beta=beta)
    rn = ps.model._get_renderer()
    if os.environ.get("NEAT_GRID_CAP"):
        rn.ctx.debug_grid_cap(int(os.environ["NEAT_GRID_CAP"]))
    inp, gt = TR.to_device(TR.host_batch(R, seed=1), dev)
    for _ in range(3):
        ps.step(inp, gt)
    out = {}
    for name, f in flag_sets:
        set_flags(f)
        ps.step(inp, gt)
        torch.cuda.synchronize()
        rn.timers = {}
        for _ in range(n):
            ps.step(inp, gt)
        torch.cuda.synchronize()
        tm = {k: round(v[1] / n, 4) for k, v in rn.timer_ms().items()}
        rn.timers = None
        out[name] = tm
        print("%-34s" % name, json.dumps(tm), flush=True)
    set_flags(0)
    return out


def parity(flag_sets, cases):
    import parity_util as PU
    res = {}
    for tag, conf, R, beta, sw, sb, img in cases:
        sd_np = synth.make_state_dict(conf, seed=sw, perturb=0.15, beta=beta)
        model = PU.make_model(conf, sd_np)
        b = synth.make_batch(R, seed=sb, **(img or {}))
        ref = None
        for name, f in flag_sets:
            set_flags(f)
            out, lo = PU.gpu_step(model, b)
            if ref is None:   # same seed -> same samples for every flag set
                oo, ol, leaves = PU.oracle_step(conf, sd_np, b, model.last_step)
                ref = {n: v.grad for n, v in leaves.items()}
            table = PU.grad_errors({n: p.grad for n, p in model.named_parameters()}, ref)
            sdf = {n: t for n, t in table.items() if n.startswith("implicit")}
            heads = {n: t for n, t in table.items() if n.startswith(("rendering", "attraction"))}
            w = lambda tb: (max(t[0] for t in tb.values()), max(t[1] for t in tb.values()))
            res[(tag, name)] = dict(sdf=w(sdf), heads=w(heads), grad_theta=PU.output_errors(out, oo)["grad_theta"],
                                    lines3d=PU.output_errors(out, oo)["lines3d"])
            print("%-20s %-28s sdf l2 %.2e max %.2e | heads l2 %.2e max %.2e | grad_theta %.2e lines3d %.2e" % (
                tag, name, *w(sdf), *w(heads), res[(tag, name)]["grad_theta"], res[(tag, name)]["lines3d"]), flush=True)
        set_flags(0)
    return res


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "timing"
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    if mode == "base":
        timing(R, [("current build", 0)], beta=float(sys.argv[3]) if len(sys.argv) > 3 else 0.1)
    elif mode == "timing":
        timing(R, [("baseline", 0), ("wgrad 2 MMAs (no x_lo*y_hi)", 2), ("wgrad 1 MMA", 3),
                   ("sdf_bwd no loads", 16), ("sdf_bwd no stores", 32), ("sdf_bwd no loads/stores", 48),
                   ("sdf_render no u/a/feat saves", 64), ("sdf_render no d1 loads", 128), ("sdf_render neither", 192),
                   ("baseline again", 0)])
    else:
        dtu = synth.dtu_conf()
        cases = [("dtu_1024_b0.1", dtu, 1024, 0.1, 5, 4, None), ("dtu_1024_b0.01", dtu, 1024, 0.01, 5, 4, None),
                 ("dtu_130", dtu, 130, 0.1, 6, 8, None),
                 ("toy_256", synth.toy_conf(), 256, 0.1, 5, 4, dict(img_res=(512, 512), focal=560.0))]
        parity([("x3 (product)", 0), ("wgrad: no x_hi*y_lo", 1), ("wgrad: no x_lo*y_hi", 2), ("wgrad: hi*hi only", 3),
                ("sigma' 16-bit fixed", 4), ("zhat from a_hi", 8), ("all of them", 15)], cases)
