import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_fused import _run
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
runs = {}
for name, kind in (("plugin_a", "plugin"), ("plugin_b", "plugin"), ("eager", "eager"), ("graphs", "graphs")):
    ts, l, p = _run(kind, steps)
    runs[name] = (l, p)
    print(name, ["%.6f" % x for x in l])
def worst(a, b):
    w = max(a, key=lambda n: float((a[n] - b[n]).abs().max() / b[n].abs().max().clamp_min(1e-12)))
    return w, float((a[w] - b[w]).abs().max() / b[w].abs().max()), float((a[w] - b[w]).abs().max())
for x in ("plugin_b", "eager", "graphs"):
    print(x, "vs plugin_a: worst tensor", worst(runs[x][1], runs["plugin_a"][1]))
