"""A few FusedTrainStep steps (two CUDA-graph replays + the junction hand-over per step): the target of the ncu launch
list of the benchmarked path.    python scripts/profile_fused.py [rays] [steps] [beta]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neat_b200 import synth
from neat_b200 import trainer as TR

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
beta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
dev = torch.device("cuda", 0)
ts = TR.FusedTrainStep(synth.dtu_conf(), device=dev, seed=42, beta=beta)
inp, gt = TR.to_device(TR.host_batch(R, seed=1), dev)
for i in range(n):
    lo = ts.step(inp, gt)
torch.cuda.synchronize()
print("steps", n, "rays", R, "loss", float(lo["loss"]), "launches/step", ts.launches_per_step)
