"""Where does the HOST time of a train step go?  cProfile over free-running steps (the step's only synchronisation is the
junction hand-over), sorted by own time and by cumulative time."""
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
for _ in range(5): ts.step(inp, gt)
torch.cuda.synchronize()
N = 40
t0 = time.perf_counter()
for _ in range(N): ts.step(inp, gt)
torch.cuda.synchronize()
print("free-running: %.3f ms/step" % ((time.perf_counter() - t0) * 1e3 / N))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(N): ts.step(inp, gt)
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr)
print("==== by own time (per step = /%d)" % N)
st.sort_stats("tottime").print_stats(38)
print("==== by cumulative time")
st.sort_stats("cumulative").print_stats(45)
