"""Which entry separates the fused (graph) path from the plugin path after 6 steps: index, values, its gradient per step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from neat_b200 import synth
from neat_b200 import trainer as TR

NAME = "implicit_network.lin3.bias"

def run(kind, steps=6, R=256):
    dev = torch.device("cuda:0")
    conf = synth.dtu_conf()
    ts = TR.TrainStep(conf, device=dev, seed=11, beta=0.1) if kind == "plugin" else \
        TR.FusedTrainStep(conf, device=dev, seed=11, beta=0.1, graphs=(kind == "graphs"))
    ts.model.seed_draws(123)
    inp, gt = TR.to_device(TR.host_batch(R, seed=3), dev)
    p = dict(ts.model.named_parameters())[NAME]
    grads, vals = [], []
    for _ in range(steps):
        ts.step(inp, gt)
        torch.cuda.synchronize()
        grads.append(p.grad.detach().clone().cpu())
        vals.append(p.detach().clone().cpu())
    return grads, vals

base_g, base_v = run("plugin")
print("|b| =", float(base_v[-1].norm()))
for i in range(14):
    kind = "graphs" if i % 2 == 0 else "plugin"
    g, v = run(kind)
    d = (v[-1] - base_v[-1]).abs()
    j = int(d.argmax())
    print(i, kind, "max |d| %.3e at %d" % (float(d[j]), j), "val base %.6f this %.6f" % (float(base_v[-1][j]), float(v[-1][j])))
    if float(d[j]) > 2e-4:
        print("   grad base:", ["%.3e" % float(x[j]) for x in base_g])
        print("   grad this:", ["%.3e" % float(x[j]) for x in g])
        print("   vals base:", ["%.6f" % float(x[j]) for x in base_v])
        print("   vals this:", ["%.6f" % float(x[j]) for x in v])
        print("   typical |grad| of the tensor: %.3e" % float(base_g[0].abs().median()), flush=True)
