import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neat_b200 import synth
from neat_b200.context import Context
conf = synth.dtu_conf()
ctx = Context(conf)
sd = {k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(conf, seed=1).items()}
ctx.pack_weights(ctx.flatten_state_dict(sd))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
fast = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx.lib.neat_set_precision(ctx._h, fast)
x = (torch.rand(R * 128, 3, device="cuda") - 0.5) * 3
for _ in range(3): ctx.sdf_points(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): ctx.sdf_points(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("fast=%d" % fast, "query R=%d: %.3f ms, %.1f Mpts/s, %.1f TFLOP/s algorithmic" % (R, ms, R * 128 / ms / 1e3, R * 128 * 1049088.0 / ms / 1e9))
