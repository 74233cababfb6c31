#!/usr/bin/env python
"""Benchmark of the NEAT attraction-field training step (BASELINE.json: train-step rays/s, DTU-shaped batch,
98 samples/ray, 8x256 SDF + 4x256 rendering / attraction nets).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--beta B] [--impl ours|reference]

A "step" is what code/training/volsdf_train.py:361-374 does for one image: model(input) -> loss -> zero_grad ->
backward -> (all-reduce of the flat gradient bucket for N>1) -> Adam step.  Rank 0 prints ONE JSON line.
  value   : whole-job rays/s with the batch already resident in HBM (device-timed, max over ranks); BASELINE configs[1]
  e2e     : the same step driven from HOST buffers (pinned H2D of the batch and a D2H read of the loss every step)
  roofline: the dominant kernel, timed live with CUDA events on the launching stream; `frac` is the TENSOR fraction
            (SURVEY 8d names the tensor cores as the bound of the MLP kernels), `hbm_frac` sits beside it
  cpu_baseline       : the UNMODIFIED reference classes on this box's host cores (bounded: 1 + 1 steps of the full batch)
  gpu_eager_baseline : the UNMODIFIED reference classes on this B200 (fp32 eager PyTorch + cuBLAS) -- what a NEAT user
                       runs today; `vs_gpu_eager` = value / that
  configs : the other BASELINE configs measured in the same run (steady-state beta = 0.01, 8192 rays/GPU, eval chunks)
  dp_check: (N > 1) all-reduced bucket == mean over ranks of the per-rank gradients, through the real plugin
`--impl reference` runs the reference arm alone: the unmodified reference's training loop on the host CPU (all threads)
at the FULL ray count, from oracle/_ref/neat_ref_code.zip (oracle/build_ref.py archives the reference's own files there;
/root/reference does not exist on the GPU box).  If that archive is missing the oracle port is timed and labelled "port"."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_SDF, F_REND, F_ATT = 1049088.0, 542720.0, 531968.0  # FLOP per point
S = 98
# dram__bytes_read.sum + dram__bytes_write.sum of one launch at 1024 rays (ncu --set full; profiles/*_ncu_full_summary.csv)
NCU_DRAM_BYTES_1024 = json.load(open(os.path.join(ROOT, "profiles", "ncu_dram_bytes_1024.json"))) \
    if os.path.exists(os.path.join(ROOT, "profiles", "ncu_dram_bytes_1024.json")) else \
    {"wgrad": 6.165e9, "sdf_bwd": 5.809e9, "sdf_render": 3.073e9, "source": "profiles/r01_v5_ncu_full_summary.csv"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference legs
def cpu_port(R_cpu, beta, steps, threads):
    """Fallback only (no reference archive): the CPU port (oracle/) of one train step."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from neat_b200 import synth
    from oracle import neat_oracle as O
    torch.set_num_threads(threads)
    conf = synth.dtu_conf()
    sd_np = synth.make_state_dict(conf, seed=1, perturb=0.15, beta=beta)
    sd = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in sd_np.items()}
    ci, c = conf["implicit_network"], conf["ray_sampler"]
    sconf = O.SamplerConf(near=c["near"], N_samples=c["N_samples"], N_samples_eval=c["N_samples_eval"],
                          N_samples_extra=c["N_samples_extra"], eps=c["eps"], beta_iters=c["beta_iters"],
                          max_total_iters=c["max_total_iters"])
    b = synth.make_batch(R_cpu, seed=1)
    T = lambda a: torch.from_numpy(np.asarray(a))
    times, k = [], 0
    for it in range(steps + 1):
        t0 = time.perf_counter()
        P = O.params_from_state_dict(sd, skip_in=tuple(ci["skip_in"]), multires=ci["multires"],
                                     multires_view=conf["rendering_network"]["multires_view"],
                                     sphere_radius=conf["scene_bounding_sphere"], sphere_scale=ci["sphere_scale"],
                                     beta_min=conf["density"]["beta_min"], track=True)
        g = torch.Generator().manual_seed(it)
        rnd = O.TrainRandoms(O.SamplerRandoms(torch.rand(R_cpu, sconf.N_samples_eval, generator=g),
                                              torch.rand(R_cpu, sconf.N_samples, generator=g),
                                              torch.randperm(sconf.N_samples_eval, generator=g)[:sconf.N_samples_extra],
                                              torch.randint(0, S, (R_cpu,), generator=g)),
                             torch.empty(R_cpu, 3).uniform_(-3, 3, generator=g))
        out = O.neat_forward(P, sconf, T(b["intrinsics"][0]), T(b["pose"][0]), T(b["uv"][0]), T(b["uv_proj"][0]),
                             gt_vertices=T(b["wf_vertices"]), training=True, rnd=rnd)
        lo = O.neat_loss(out, T(b["rgb"][0]), T(b["lines2d"][0]), out["K"])
        for v in sd.values():
            v.grad = None
        lo["loss"].backward()
        k = out["n_sampler_iters"]
        if it > 0:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"rays_per_s": R_cpu / sec, "ms_per_step": sec * 1e3, "steps": steps, "warmup": 1, "rays": R_cpu,
            "threads": threads, "device": "cpu", "k": k}


def reference_cpu(rays, beta, steps, warmup, budget_s):
    """The reference's CPU path on this box's host cores: (result dict, kind)."""
    threads = os.cpu_count() or 1
    try:
        from oracle import ref_bench
        if ref_bench.available():
            return ref_bench.train_steps(rays, steps, warmup, beta=beta, device="cpu", threads=threads,
                                         budget_s=budget_s), "reference"
    except Exception as e:  # never lose the whole line to the baseline leg
        sys.stderr.write("reference CPU leg failed (%s: %s); timing the oracle port instead\n" % (type(e).__name__, e))
    return cpu_port(min(rays, 256), beta, max(1, min(steps, 2)), threads), "port"


def reference_config(r, kind, beta):
    what = ("UNMODIFIED reference classes (model.networks.neat_wfr_rend_a.VolSDFNetwork + loss_wfr.VolSDFLoss, "
            "oracle/_ref/neat_ref_code.zip)" if kind == "reference" else "oracle/neat_oracle.py (CPU port of the reference)")
    return {"workload": "DTU-shaped synthetic batch: %d rays x 98 samples, 8x256 SDF + 4x256 rendering/attraction MLPs, "
                        "ErrorBoundSampler; train step = fwd+loss+zero_grad+backward+torch.optim.Adam, as "
                        "code/training/volsdf_train.py:361-374" % r["rays"],
            "implementation": what, "device": "host CPU, torch fp32 eager + autograd (double backward via create_graph)",
            "threads": r["threads"], "rays_per_step": r["rays"], "samples_per_ray": S, "beta": beta, "parallelism": "none"}


# ------------------------------------------------------------------------------------------------ our arm
def train_config(rays, beta, steps, warmup, rank, world, dev, lib, e2e=True, timers=True, clocks=None):
    """Times `steps` training steps at `rays` rays per GPU.  Returns a dict of raw measurements (max over ranks).
    value / e2e: neat_b200.trainer.FusedTrainStep (the step as two CUDA-graph replays + the junction hand-over);
    kernel timers and `plugin_ms`: the same step through the plugin classes (model -> loss -> backward -> Adam, eager
    launches), where CUDA events can bracket the individual kernels."""
    import torch
    import torch.distributed as dist
    from neat_b200 import synth
    from neat_b200 import trainer as TR

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(warmup, 4)   # 2 eager steps, the capture, one replay
    hb = TR.host_batch(rays, seed=1 + rank)
    inp, gt = TR.to_device(hb, dev)
    ts = TR.FusedTrainStep(synth.dtu_conf(), device=dev, seed=42, beta=beta)
    for _ in range(warmup):
        ts.step(inp, gt)
    barrier()
    if clocks is not None:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k_acc = torch.zeros(1, dtype=torch.int32, device=dev)  # the sampler's k is data dependent and drifts as beta trains
    wait_ms = 0.0
    ts.profile = []
    for _ in range(steps):
        ts.step(inp, gt)
        k_acc += ts.st.n_iters
        wait_ms += ts.last_host_ms["wait_for_gpu"]
    e1.record()
    barrier()
    prof = ts.profile_ms() or {}
    ts.profile = None
    r = {"rays": rays, "graph_ms": prof, "beta": beta, "steps": steps, "warmup": warmup, "ms_total": e0.elapsed_time(e1),
         "launches": int(ts.launches_per_step * steps), "launches_per_step": int(ts.launches_per_step),
         "k_last": int(ts.st.n_iters.item()), "k_mean": float(k_acc.item()) / steps,
         "host_junction_block": dict(ts.last_host_ms, mean_wait_for_gpu=wait_ms / steps), "h2d": TR.h2d_bytes(hb)}
    if e2e:
        # end to end from HOST buffers (same initial state and trajectory as the loop above): the step copies the pinned
        # host batch to the device and the loss is read back, every step
        del ts
        _release()
        ts = TR.FusedTrainStep(synth.dtu_conf(), device=dev, seed=42, beta=beta)
        for _ in range(warmup):
            ts.step(hb, hb)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        loss_host = 0.0
        for _ in range(steps):
            loss_host = float(ts.step(hb, hb)["loss"].item())
        e3.record()
        barrier()
        r["ms_e2e"] = e2.elapsed_time(e3)
        r["last_loss"] = loss_host
    if clocks is not None:
        r["clocks"] = clocks.stop()
    del ts
    _release()
    tm = {}
    if timers:
        # per-kernel CUDA-event timers: the plugin path, eager launches
        ps = TR.TrainStep(synth.dtu_conf(), device=dev, seed=42, beta=beta)
        rn = ps.model._get_renderer()
        for _ in range(3):
            ps.step(inp, gt)
        barrier()
        n_t = min(steps, 20)
        rn.timers = {}
        l0 = lib.neat_launch_count()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(n_t):
            ps.step(inp, gt)
        p1.record()
        barrier()
        tm = {k: (v[0] / n_t * steps, v[1] / n_t * steps) for k, v in rn.timer_ms().items()}  # scaled to `steps`
        rn.timers = None
        r["plugin_ms_per_step"] = p0.elapsed_time(p1) / n_t
        r["plugin_launches_per_step"] = (lib.neat_launch_count() - l0) / n_t
        del ps, rn
        _release()
    r["timers"] = tm
    t = torch.tensor([r["ms_total"], r.get("ms_e2e", 0.0), r.get("plugin_ms_per_step", 0.0)], device=dev, dtype=torch.float64)
    if world > 1:
        mine = torch.tensor([r["ms_total"] / steps, float(r["k_last"]), sum(v[1] for v in tm.values()) / steps,
                             r.get("last_loss", 0.0), r["host_junction_block"]["mean_wait_for_gpu"], r["k_mean"],
                             prof.get("graph_A_ms", 0.0), prof.get("graph_B_ms", 0.0), prof.get("gap_ms", 0.0)], device=dev,
                            dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        r["per_rank"] = [{"ms_per_step": round(float(a[0]), 3), "sampler_k": int(a[1]), "mlp_kernel_ms": round(float(a[2]), 3),
                          "last_loss": round(float(a[3]), 5), "host_wait_for_handover_ms": round(float(a[4]), 3),
                          "sampler_k_mean": round(float(a[5]), 3), "graph_A_ms": round(float(a[6]), 3),
                          "graph_B_ms_incl_allreduce_wait": round(float(a[7]), 3), "gap_ms": round(float(a[8]), 3)}
                         for a in allr]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    r["ms_total"], r["ms_e2e"], r["plugin_ms_per_step"] = float(t[0]), float(t[1]), float(t[2])
    return r


def _release():
    """Drop the previous configuration's workspaces (GBs of save records; the renderer / sampler / graph objects form
    reference cycles, so collect before asking the allocator to return the memory)."""
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()


def kernel_fractions(r, pk):
    """Per-kernel algorithmic TFLOP/s and GB/s (and their fractions of the measured peaks) from the live CUDA-event timers."""
    rays, M, tm = r["rays"], r["rays"] * S, r["timers"]
    # Tensor work per point: forward F + normal pass F (sdf_render); tangent F + reverse F (sdf_bwd); the two
    # outer-product accumulations 2 F plus the heads' (wgrad) -- SURVEY.md Appendix A, 6 F_sdf per render point.
    flops = {"sdf_bwd_M%d" % M: 2 * F_SDF * M, "sdf_render_M%d" % M: 2 * F_SDF * M,
             "sampler": 128.0 * r["k_mean"] * rays * F_SDF,
             "head_fwd": 0.5 * (F_REND + F_ATT) * M, "head_bwd": 0.5 * (F_REND + F_ATT) * M,  # per launch (one head)
             "wgrad": (2 * F_SDF + F_REND + F_ATT) * M + 2 * F_SDF * 2 * rays}
    # Algorithmic bytes = compulsory HBM traffic per 128-point tile of the save-record design
    MAIN, AUX = 131072.0, 24576.0
    WG_SDF_B, WG_HEAD_B = (33 * MAIN + 5 * AUX) / 128.0, 2 * (9 * MAIN + 2 * AUX) / 128.0
    BWD_B = (16 * MAIN + 8 * MAIN + MAIN + (8 * MAIN + AUX) + (9 * MAIN + AUX)) / 128.0 + 20.0
    REND_B = (8 * MAIN + 8 * MAIN + 8 * MAIN + AUX + MAIN) / 128.0 + 20.0
    hbm_bytes = {"wgrad": (WG_SDF_B + WG_HEAD_B) * M + WG_SDF_B * 2 * rays, "sdf_bwd_M%d" % M: BWD_B * M,
                 "sdf_render_M%d" % M: REND_B * M}
    out = {}
    for k, (n_l, ms) in tm.items():
        if k not in flops or n_l == 0:
            continue
        sec = ms / n_l * 1e-3
        d = {"ms_per_launch": round(ms / n_l, 4), "tensor_tflops": round(flops[k] / sec / 1e12, 1),
             "tensor_frac": round(flops[k] / sec / 1e12 / pk["bf16_tflops_sustained"], 4)}
        if k in hbm_bytes:
            d["hbm_gbs"] = round(hbm_bytes[k] / sec / 1e9, 1)
            d["hbm_frac"] = round(hbm_bytes[k] / sec / 1e9 / pk["hbm_gbs"], 4)
        out[k] = d
    return out


def eval_config(rays, beta, steps, warmup, rank, world, dev, lib):
    """The eval-mode forward (sampler + SDF with normals + both heads + compositing + geometry, no backward) over a chunk
    of `rays` rays per GPU -- BASELINE configs[4], chunked as code/utils/general.py:23-36 / neat-final-parsing.py:205-213
    do; rays shard over ranks with no collective."""
    import torch
    import torch.distributed as dist
    from neat_b200 import synth
    from neat_b200 import trainer as TR
    from neat_b200.model import VolSDFNetwork
    torch.manual_seed(42)
    model = VolSDFNetwork(synth.dtu_conf())
    with torch.no_grad():
        model.density.beta.fill_(beta)
    model = model.to(dev).eval()
    hb = TR.host_batch(rays, seed=1 + rank)
    inp, _ = TR.to_device(hb, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        out = model(inp)
    barrier()
    l0 = lib.neat_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = model(inp)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.neat_launch_count() - l0
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(steps):
        i2, _ = TR.to_device(hb, dev)
        o = model(i2)
        host = o["lines3d"].cpu()  # the result a caller keeps (neat-final-parsing.py:213)
    e3.record()
    barrier()
    t = torch.tensor([ms, e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    k = int(model._get_renderer().last_n_iters.item())
    r = {"metric": "eval_forward_rays_per_sec", "value": world * rays * steps / (ms * 1e-3), "unit": "rays/s",
         "rays_per_gpu": rays, "beta": beta, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "sampler_k": k,
         "e2e": {"value": world * rays * steps / (ms_e2e * 1e-3), "unit": "rays/s",
                 "h2d_bytes_per_step": sum(hb[kk].numel() * 4 for kk in ("intrinsics", "pose", "uv", "uv_proj")),
                 "d2h_bytes_per_step": int(host.numel() * 4)},
         "gpu_launches": int(launches),
         "full_image_1600x1200_seconds_at_this_rate": round(1600 * 1200 / (world * rays * steps / (ms * 1e-3)), 3)}
    del model
    _release()
    return r


def fast_mode_config(rays, rank, world, dev, steps=10):
    """BASELINE configs[2] says "bf16 tensor-core MLPs": the opt-in one-MMA-per-MAC mode (neat_set_precision(ctx, 1), hi
    planes only) at 8192 rays/GPU, with its MEASURED distance from the x3 parity mode on the first step (same weights,
    same draws).  Informational: it misses north_star's 1e-4 bound, so nothing else in this file uses it."""
    import torch
    import torch.distributed as dist
    from neat_b200 import _lib, synth
    from neat_b200 import trainer as TR
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    outs, ms = {}, {}
    for fast in (0, 1):
        ts = TR.FusedTrainStep(synth.dtu_conf(), device=dev, seed=42, beta=0.1)
        inp, gt = TR.to_device(TR.host_batch(rays, seed=1 + rank), dev)
        rn = ts.model._get_renderer()
        _lib.check(rn.ctx.lib.neat_set_precision(rn.ctx._h, fast))
        ts.model.seed_draws(7)
        ts.step(inp, gt)
        torch.cuda.synchronize()
        outs[fast] = {k: ts.out[k].detach().float().cpu() for k in ("rgb_values", "lines3d", "grad_theta")}
        if fast:
            for _ in range(4):
                ts.step(inp, gt)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                ts.step(inp, gt)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        del ts, rn
        _release()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
    return {"name": "configs[2] in the one-MMA mode (\"bf16 tensor-core MLPs\": hi planes only, neat_set_precision(ctx, 1)): "
                    "8192 rays/GPU, beta=0.1; NOT the parity mode",
            "rays_per_gpu": rays, "beta": 0.1, "steps": steps, "warmup": 5, "value": world * rays * steps / (ms * 1e-3),
            "unit": "rays/s", "ms_per_step": ms / steps,
            "measured_distance_from_x3_first_step": {k: float("%.3g" % rel(outs[1][k], outs[0][k])) for k in outs[0]}}


def dp_check(rank, world, dev):
    """N > 1: one step through the real plugin; the all-reduced flat bucket times 1/world must equal the mean of the
    per-rank gradients (gathered before the reduction)."""
    import torch
    import torch.distributed as dist
    from neat_b200 import synth
    from neat_b200 import trainer as TR
    ts = TR.TrainStep(synth.dtu_conf(), device=dev, seed=42, beta=0.1)
    inp, gt = TR.to_device(TR.host_batch(256, seed=1 + rank), dev)
    lo = ts.loss_fn(ts.model(inp), gt)
    ts.bucket.zero()
    lo["loss"].backward()
    mine = ts.bucket.flat.clone()
    allg = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allg, mine)
    mean = torch.stack(allg).double().mean(0)
    scale = ts.bucket.all_reduce_sum()
    red = ts.bucket.flat.double() * scale
    err = float((red - mean).abs().max() / mean.abs().max().clamp_min(1e-30))
    t = torch.tensor([err], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    losses = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(losses, lo["loss"].detach().reshape(1).float())
    del ts
    _release()
    return {"allreduce_vs_mean_of_rank_gradients_rel_err": float(t[0]), "ok": bool(float(t[0]) < 1e-5),
            "rank_losses": [round(float(x), 5) for x in losses], "rays_per_rank": 256}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rays", type=int, default=1024, help="rays per GPU per step (weak scaling)")
    ap.add_argument("--beta", type=float, default=0.1, help="density.beta (0.1 = init, k~2; 0.01 = trained-like, k=5)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=0, help="reference legs: rays per step (0 = the full --rays)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    ap.add_argument("--mode", default="train", choices=["train", "eval"],
                    help="train: the headline train step; eval: only the eval-mode forward over chunks of `--rays` rays")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warm = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        rays = args.cpu_rays or args.rays
        r, kind = reference_cpu(rays, args.beta, args.steps, max(args.warmup, 1), budget_s=200.0)
        sample = ("the full batch: %d rays per step, %d timed step(s) after %d untimed (requested %d + %d; bounded to ~200 s "
                  "of host time), %d host threads" % (r["rays"], r["steps"], r["warmup"], args.steps, args.warmup, r["threads"]))
        val = r["rays_per_s"]
        print(json.dumps({"impl": "reference", "metric": "train_step_rays_per_sec", "value": val, "unit": "rays/s",
                          "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": reference_config(r, kind, args.beta),
                          "cpu_baseline": {"value": val, "unit": "rays/s", "cores": r["threads"], "kind": kind, "sample": sample},
                          "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "last_loss": r.get("loss")}))
        return

    import torch
    import torch.distributed as dist
    from neat_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    if args.mode == "eval":
        r = eval_config(args.rays, args.beta, args.steps, warm, rank, world, dev, lib)
        if rank == 0:
            r.update(n_gpus=world, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16x3->f32", data="synthetic",
                     config={"workload": "eval-mode forward, %d rays/GPU per call x 98 samples, DTU nets" % args.rays,
                             "beta": args.beta, "parallelism": "dp%d" % world})
            print(json.dumps(r))
        if world > 1:
            dist.destroy_process_group()
        return

    config = {"workload": "DTU-shaped synthetic batch: %d rays/GPU x 98 samples, 8x256 SDF + 4x256 rendering/attraction "
                          "MLPs, ErrorBoundSampler (<=5 x 128 SDF queries/ray), train step = fwd+loss+bwd+Adam "
                          "(BASELINE configs[1])" % args.rays,
              "optimizer": "Adam(lr=5e-4), all parameter tensors in one launch, step count / lr on the device",
              "execution": "neat_b200.trainer.FusedTrainStep: two CUDA-graph replays per step + the host junction hand-over",
              "rays_per_gpu": args.rays, "samples_per_ray": S, "beta": args.beta, "parallelism": "dp%d" % world,
              "rng": "training draws (stratified jitter, inverse-CDF u, extra columns, eikonal points) made on the device",
              "precision_mode": "bf16x3 (hi/lo split operands, fp32 accumulate) on tcgen05",
              "l2": "no explicit flush: the per-step working set (GBs of saved activations at 1024 rays) is >> the 126 MB L2"}
    check = dp_check(rank, world, dev) if world > 1 else None
    clocks = ClockSampler(local) if rank == 0 else None
    r = train_config(args.rays, args.beta, args.steps, warm, rank, world, dev, lib, clocks=clocks)
    warm = r["warmup"]

    extras = []
    if not args.no_extra_configs:
        # the other BASELINE configs, same run, each with its own ms/step, sampler k and kernel fractions
        plan = [("configs[1] at the steady-state density: 1024 rays/GPU, beta=0.01 (sampler k=5)", 1024, 0.01, 30),
                ("configs[1], 100 timed steps (beta trains away from 0.1: k drifts)", 1024, 0.1, 100),
                ("configs[2]/[3]: 8192 rays/GPU, beta=0.1", 8192, 0.1, 10),
                ("configs[2]/[3]: 8192 rays/GPU, beta=0.01 (k=5)", 8192, 0.01, 10)]
        pk0, _ = peaks()
        for name, rays, beta, steps in plan:
            if rays == args.rays and beta == args.beta and steps == args.steps:
                continue
            x = train_config(rays, beta, steps, 3, rank, world, dev, lib, e2e=False)
            extras.append({"name": name, "rays_per_gpu": rays, "beta": beta, "steps": steps, "warmup": x["warmup"],
                           "value": world * rays * steps / (x["ms_total"] * 1e-3), "unit": "rays/s",
                           "ms_per_step": x["ms_total"] / steps, "plugin_path_ms_per_step": round(x["plugin_ms_per_step"], 4),
                           "sampler_k_last": x["k_last"],
                           "sampler_k_mean": round(x["k_mean"], 3), "gpu_launches_per_step": x["launches"] / steps,
                           "kernel_ms_per_step": {k: round(v[1] / steps, 4) for k, v in x["timers"].items()},
                           "kernels": kernel_fractions(x, pk0)})
        e = eval_config(65536, 0.01, 5, 3, rank, world, dev, lib)
        e["name"] = "configs[4]: eval-mode forward in 65536-ray chunks per GPU (full 1600x1200 image = 30 chunks), beta=0.01"
        extras.append(e)
        try:
            extras.append(fast_mode_config(8192, rank, world, dev))
        except Exception as ex:  # informational entry: never lose the line to it
            extras.append({"name": "configs[2] in the one-MMA mode", "failed": "%s: %s" % (type(ex).__name__, ex)})

    if rank == 0:
        pk, pk_kind = peaks()
        steps = args.steps
        ms_total, ms_e2e, timers = r["ms_total"], r["ms_e2e"], r["timers"]
        value = world * args.rays * steps / (ms_total * 1e-3)
        e2e = world * args.rays * steps / (ms_e2e * 1e-3)
        M = args.rays * S
        kf = kernel_fractions(r, pk)
        # dominant KERNEL = longest single launch ("sampler" is a group of up to 5 query launches + the per-ray kernels)
        dom = max((k for k in kf if k != "sampler"), key=lambda k: kf[k]["ms_per_launch"])
        d = kf[dom]
        ncu_traffic = NCU_DRAM_BYTES_1024.get(dom.split("_M")[0]) if args.rays == 1024 else None
        big = sum(v[1] for k, v in timers.items()) / steps
        line = {"metric": "train_step_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16x3->f32", "data": "synthetic", "config": config,
                "sampler_iters_k": r["k_last"], "sampler_iters_k_mean": round(r["k_mean"], 3),
                "mlp_samples_per_sec": world * (M + 128 * r["k_mean"] * args.rays + 3 * args.rays) * steps / (ms_total * 1e-3),
                "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / steps, "last_loss": r["last_loss"]},
                "gpu_launches": r["launches"], "clocks": r.get("clocks"),
                "roofline": {"bound": "tensor", "kernel": dom, "achieved": d["tensor_tflops"], "peak": pk["bf16_tflops_sustained"],
                             "unit": "TFLOP/s", "frac": d["tensor_tflops"] / pk["bf16_tflops_sustained"],
                             "hbm_frac": d.get("hbm_frac"), "hbm_gbs_algorithmic": d.get("hbm_gbs"),
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at 1024 rays (ncu --set full)
                             "traffic": ncu_traffic, "traffic_source": NCU_DRAM_BYTES_1024.get("source"),
                             "peak_kind": pk_kind + " sustained bf16 (cuBLAS) / " + pk_kind + " copy bandwidth for hbm_frac",
                             "note": "achieved = ALGORITHMIC fp32-equivalent FLOPs (2*MAC, SURVEY 8d) / CUDA-event time of the "
                                     "longest single launch; the kernel issues 3 bf16 MMAs per algorithmic MAC (hi*hi + hi*lo + "
                                     "lo*hi) to meet the 1e-4 parity bound, so the tensor pipe does 3x this.  hbm_frac = the "
                                     "kernel's compulsory save-record bytes / time / copy bandwidth.",
                             "step_tensor_frac": round((947.5e6 + 134.3e6 * r["k_mean"]) * value / world / 1e12
                                                       / pk["bf16_tflops_sustained"], 4),
                             "kernels": kf},
                "kernel_ms_per_step": {k: round(v[1] / steps, 4) for k, v in timers.items()},
                "step_minus_big_kernels_ms": round(ms_total / steps - big, 4),
                "plugin_path": {"ms_per_step": round(r["plugin_ms_per_step"], 4), "value": world * args.rays / (r["plugin_ms_per_step"] * 1e-3),
                                "gpu_launches_per_step": r["plugin_launches_per_step"],
                                "what": "the same step through the drop-in plugin classes (VolSDFNetwork.forward -> VolSDFLoss -> "
                                        "backward -> neat_b200.optim.Adam, eager launches): where the per-kernel CUDA-event "
                                        "timers of `kernel_ms_per_step` / `roofline` are taken"},
                "host_junction_block": r["host_junction_block"], "graph_ms": r.get("graph_ms"),
                "per_rank": r.get("per_rank"), "dp_check": check,
                "configs": extras}
        if world == 1 and not args.no_eager_baseline:
            try:
                from oracle import ref_bench
                if ref_bench.available():
                    g = ref_bench.train_steps(args.cpu_rays or args.rays, 10, 3, beta=args.beta, device="cuda:%d" % local)
                    line["gpu_eager_baseline"] = {
                        "value": g["rays_per_s"], "unit": "rays/s", "ms_per_step": g["ms_per_step"], "steps": g["steps"],
                        "warmup": g["warmup"], "rays": g["rays"], "kind": "reference",
                        "what": "the UNMODIFIED reference classes on this B200: fp32 eager PyTorch + cuBLAS, torch.optim.Adam, "
                                "the loop of volsdf_train.py:361-374 (host sync per sampler iteration, sklearn DBSCAN and scipy "
                                "assignment on the host, as shipped)"}
                    line["vs_gpu_eager"] = value / g["rays_per_s"]
                else:
                    line["gpu_eager_baseline"] = {"unavailable": "oracle/_ref/neat_ref_code.zip not built (oracle/build_ref.py)"}
            except Exception as ex:
                line["gpu_eager_baseline"] = {"unavailable": "%s: %s" % (type(ex).__name__, ex)}
        if world == 1 and not args.no_cpu_baseline:
            c, kind = reference_cpu(args.cpu_rays or args.rays, args.beta, 1, 1, budget_s=30.0)
            line["cpu_baseline"] = {"value": c["rays_per_s"], "unit": "rays/s", "cores": c["threads"], "kind": kind,
                                    "sample": "%d rays (of %d) per step, %d timed step after %d untimed, torch CPU fp32, all "
                                              "host threads; %s" % (c["rays"], args.rays, c["steps"], c["warmup"],
                                                                    "the unmodified reference classes" if kind == "reference"
                                                                    else "oracle port")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
