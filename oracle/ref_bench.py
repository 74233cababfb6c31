"""TEST / BENCHMARK INFRASTRUCTURE ONLY -- times the UNMODIFIED reference (cherubicXN/neat) on the benchmark's workload.

The loop is code/training/volsdf_train.py:361-374 verbatim in structure: inputs `.cuda()`ed, `model(model_input)`,
`loss(model_outputs, ground_truth)`, `optimizer.zero_grad()`, `loss.backward()`, `optimizer.step()` with
`torch.optim.Adam(lr)` (:178).  The classes are the reference's own `model.networks.neat_wfr_rend_a.VolSDFNetwork` and
`model.networks.loss_wfr.VolSDFLoss`, imported by oracle/ref_shim.py from /root/reference/code (build container) or from
oracle/_ref/neat_ref_code.zip (GPU box; the same files archived unmodified by oracle/build_ref.py).

Used only by bench.py: `--impl reference` / `cpu_baseline` (device="cpu": the reference's CPU path on the host cores,
`.cuda()` neutralised by the shim) and `gpu_eager_baseline` (device="cuda": the reference as a NEAT user runs it today,
fp32 eager PyTorch + cuBLAS on the same B200).  Never imported by the product path."""
import os
import time

import numpy as np
import torch

from neat_b200 import synth
from oracle import ref_shim


def available():
    return ref_shim.available()


def _build(conf, seed, beta, device):
    Net, Loss, _, _ = ref_shim.load_classes()
    torch.manual_seed(seed)
    model = Net(conf=ref_shim.to_config(conf))
    if beta is not None:
        with torch.no_grad():
            model.density.beta.fill_(beta)
    loss = Loss(**synth.loss_conf())
    if device != "cpu":
        model = model.to(device)
    return model, loss


def train_steps(R, steps, warmup, beta=0.1, device="cpu", threads=None, budget_s=None, seed=42, batch_seed=1,
                camera_seed=1):
    """Runs warmup + steps training steps of the unmodified reference on a synthetic DTU-shaped batch of R rays
    (the same generator, seeds and camera as bench.py's own arm).  budget_s: if the first step predicts that
    warmup + steps would exceed it, fewer timed steps are run (never fewer than 1; the numbers say how many).
    Returns dict(rays_per_s, ms_per_step, steps, warmup, k, loss, threads, device)."""
    import math
    on_gpu = device != "cpu"
    ref_shim.install()
    ref_shim.force_cpu(not on_gpu)
    if threads:
        torch.set_num_threads(threads)
    conf = synth.dtu_conf()
    model, loss_fn = _build(conf, seed, beta, device)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=5.0e-4)           # volsdf_train.py:178
    a = 0.3 + 0.7 * camera_seed
    pose = synth.look_at_pose((2.5 * math.cos(a) * 0.9, 2.5 * math.sin(a) * 0.9, 2.5 * 0.436))
    b = synth.make_batch(R, seed=batch_seed, pose=pose)
    t = lambda x: torch.from_numpy(np.asarray(x))
    wf = ref_shim.Wireframe(b["wf_vertices"], b["wf_edges"], b["wf_weights"])
    host_in = {"intrinsics": t(b["intrinsics"]), "uv": t(b["uv"]), "pose": t(b["pose"]), "uv_proj": t(b["uv_proj"])}
    gt = {"rgb": t(b["rgb"]), "lines2d": t(b["lines2d"])}
    dev = torch.device(device)

    def one_step():
        # volsdf_train.py:362-374 (the dataset already holds uv_proj on the device: scene_hawp_dataset.py)
        mi = {k: (v.to(dev) if on_gpu else v) for k, v in host_in.items()}
        mi["wireframe"] = [wf]
        out = model(mi)
        lo = loss_fn(out, gt)
        opt.zero_grad()
        lo["loss"].backward()
        opt.step()
        return lo["loss"]

    def sync():
        if on_gpu:
            torch.cuda.synchronize(dev)

    sync()
    t0 = time.perf_counter()
    one_step()
    sync()
    first = time.perf_counter() - t0
    done_warm = 1
    if budget_s is not None and first * (warmup + steps) > budget_s:
        warmup = 1
        steps = max(1, min(steps, int(budget_s / first) - 1))
    while done_warm < warmup:
        one_step()
        done_warm += 1
    sync()
    t0 = time.perf_counter()
    loss = None
    for _ in range(steps):
        loss = one_step()
    loss = float(loss)          # the read a trainer's logging does; a device sync on the GPU
    sync()
    sec = (time.perf_counter() - t0) / steps
    ref_shim.force_cpu(not torch.cuda.is_available())
    return {"rays_per_s": R / sec, "ms_per_step": sec * 1e3, "steps": steps, "warmup": max(warmup, 1), "rays": R,
            "loss": loss, "threads": torch.get_num_threads(), "device": device, "first_step_s": first}
