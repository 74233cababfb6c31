"""Where does the host time of a train step go?  (synchronised wall-clock per phase)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neat_b200 import synth
from neat_b200 import trainer as TR

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ts = TR.TrainStep(synth.dtu_conf(), device="cuda:0", seed=42, beta=0.1)
hb = TR.host_batch(R, seed=1)
inp, gt = TR.to_device(hb, "cuda:0")
for _ in range(3): ts.step(inp, gt)
torch.cuda.synchronize()
import cProfile, pstats
def phase(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
acc = {}
N = 10
for _ in range(N):
    out, t = phase(lambda: ts.model(inp)); acc["forward(model)"] = acc.get("forward(model)", 0) + t
    lo, t = phase(lambda: ts.loss_fn(out, gt)); acc["loss"] = acc.get("loss", 0) + t
    _, t = phase(lambda: ts.bucket.zero()); acc["zero"] = acc.get("zero", 0) + t
    _, t = phase(lambda: lo["loss"].backward()); acc["backward"] = acc.get("backward", 0) + t
    _, t = phase(lambda: ts.opt.step()); acc["adam"] = acc.get("adam", 0) + t
print({k: round(v / N, 3) for k, v in acc.items()}, "sum", round(sum(acc.values()) / N, 3))
pr = cProfile.Profile(); pr.enable()
for _ in range(5): ts.step(inp, gt)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
