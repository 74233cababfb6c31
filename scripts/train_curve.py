"""EVIDENCE SCRIPT (checker side, not the product path): the loss trajectory of N training steps on one fixed synthetic
DTU-shaped batch -- the UNMODIFIED reference (eager PyTorch on this GPU, oracle/ref_bench.py's loop = volsdf_train.py:361-374)
next to neat_b200.trainer.FusedTrainStep, same initial weights (seed 42), same batch, each with its own random draws.
    python scripts/train_curve.py [steps] [rays]  ->  gpurun_out/train_curve.json"""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from neat_b200 import synth
from neat_b200 import trainer as TR

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
every = 25
dev = torch.device("cuda", 0)
res = {"steps": steps, "rays": R, "every": every}

# ---- ours
ts = TR.FusedTrainStep(synth.dtu_conf(), device=dev, seed=42, beta=0.1)
inp, gt = TR.to_device(TR.host_batch(R, seed=1), dev)
ours, ks = [], []
t0 = time.perf_counter()
for i in range(steps):
    lo = ts.step(inp, gt)
    if i % every == 0 or i == steps - 1:
        ours.append((i, float(lo["loss"]), float(lo["rgb_loss"]), float(lo["eikonal_loss"])))
        ks.append(int(ts.st.n_iters.item()))
torch.cuda.synchronize()
res["ours"] = {"loss": ours, "sampler_k": ks, "seconds": time.perf_counter() - t0,
               "beta_end": float(ts.model.density.beta.detach().abs()) }
del ts
torch.cuda.empty_cache()

# ---- the unmodified reference on the same GPU
try:
    from oracle import ref_bench, ref_shim
    ref_shim.install()
    ref_shim.force_cpu(False)
    conf = synth.dtu_conf()
    model, loss_fn = ref_bench._build(conf, 42, 0.1, "cuda:0")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=5.0e-4)
    a = 0.3 + 0.7 * 1
    pose = synth.look_at_pose((2.5 * math.cos(a) * 0.9, 2.5 * math.sin(a) * 0.9, 2.5 * 0.436))
    b = synth.make_batch(R, seed=1, pose=pose)
    t = lambda x: torch.from_numpy(np.asarray(x))
    wf = ref_shim.Wireframe(b["wf_vertices"], b["wf_edges"], b["wf_weights"])
    host_in = {"intrinsics": t(b["intrinsics"]), "uv": t(b["uv"]), "pose": t(b["pose"]), "uv_proj": t(b["uv_proj"])}
    gtr = {"rgb": t(b["rgb"]), "lines2d": t(b["lines2d"])}
    ref = []
    t0 = time.perf_counter()
    for i in range(steps):
        mi = {k: v.to(dev) for k, v in host_in.items()}
        mi["wireframe"] = [wf]
        lo = loss_fn(model(mi), gtr)
        opt.zero_grad()
        lo["loss"].backward()
        opt.step()
        if i % every == 0 or i == steps - 1:
            ref.append((i, float(lo["loss"]), float(lo["rgb_loss"]), float(lo["eikonal_loss"])))
    torch.cuda.synchronize()
    res["reference_eager_gpu"] = {"loss": ref, "seconds": time.perf_counter() - t0,
                                  "beta_end": float(model.density.beta.detach().abs())}
except Exception as e:
    res["reference_eager_gpu"] = {"failed": "%s: %s" % (type(e).__name__, e)}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "train_curve.json"), "w"), indent=1)
print("step   ours.loss   ref.loss      (rgb, eikonal: ours | ref)")
rl = {i: (l, r, e) for i, l, r, e in res.get("reference_eager_gpu", {}).get("loss", [])}
for i, l, r, e in ours:
    x = rl.get(i)
    print("%4d   %.5f    %s     %.4f %.4f | %s" % (i, l, "%.5f" % x[0] if x else "-", r, e, "%.4f %.4f" % (x[1], x[2]) if x else "-"))
print("seconds: ours %.1f, reference %.1f" % (res["ours"]["seconds"], res.get("reference_eager_gpu", {}).get("seconds", float("nan"))))
