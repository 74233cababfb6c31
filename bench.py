#!/usr/bin/env python
"""Benchmark of the NEAT attraction-field training step (BASELINE.json: train-step rays/s, DTU-shaped batch,
98 samples/ray, 8x256 SDF + 4x256 rendering / attraction nets).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--beta B] [--impl ours|reference]

A "step" is what code/training/volsdf_train.py:361-374 does for one image: model(input) -> loss -> zero_grad ->
backward -> (all-reduce of the flat gradient bucket for N>1) -> Adam step.  Rank 0 prints ONE JSON line.
  value : whole-job rays/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e   : the same step driven from HOST buffers (pinned H2D of the batch and a D2H read of the loss every step)
  roofline     : the dominant kernel, timed live with CUDA events on the launching stream
  cpu_baseline : the CPU port of the reference algorithm (oracle/) on this box's host cores, bounded sample
`--impl reference` times that CPU path alone (the reference itself is Python under /root/reference, which does not
exist on the GPU box; the oracle is its line-by-line restatement pinned against the reference's outputs)."""
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
# ncu DRAM bytes (read + write) of one launch at 1024 rays (profiles/r01_v5_ncu_full_summary.csv)
NCU_DRAM_BYTES_1024 = {"wgrad": 6.165e9, "sdf_bwd": 5.809e9, "sdf_render": 3.073e9}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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


def cpu_baseline(R_cpu, beta, steps, threads):
    """The CPU port (oracle/) of one train step: forward + loss + autograd backward on `threads` host threads."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from neat_b200 import synth
    from oracle import neat_oracle as O
    torch.set_num_threads(threads)
    conf = synth.dtu_conf()
    sd_np = synth.make_state_dict(conf, seed=1, perturb=0.15, beta=beta)
    sd = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in sd_np.items()}
    ci = conf["implicit_network"]
    c = conf["ray_sampler"]
    sconf = O.SamplerConf(near=c["near"], N_samples=c["N_samples"], N_samples_eval=c["N_samples_eval"],
                          N_samples_extra=c["N_samples_extra"], eps=c["eps"], beta_iters=c["beta_iters"],
                          max_total_iters=c["max_total_iters"])
    b = synth.make_batch(R_cpu, seed=1)
    T = lambda a: torch.from_numpy(np.asarray(a))
    times = []
    k = 0
    for it in range(steps + 1):
        t0 = time.perf_counter()
        P = O.params_from_state_dict(sd, skip_in=tuple(ci["skip_in"]), multires=ci["multires"],
                                     multires_view=conf["rendering_network"]["multires_view"],
                                     sphere_radius=conf["scene_bounding_sphere"], sphere_scale=ci["sphere_scale"],
                                     beta_min=conf["density"]["beta_min"], track=True)
        g = torch.Generator().manual_seed(it)
        L_guess = sconf.N_samples_eval * sconf.max_total_iters
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
    return R_cpu / sec, sec, k


def eval_mode(args, rank, world, dev, lib):
    """Extra, informational: the eval-mode forward (sampler + SDF with normals + both heads + compositing + geometry,
    no backward) over a chunk of rays per GPU; rays shard over ranks with no collective."""
    import torch
    import torch.distributed as dist
    from neat_b200 import synth
    from neat_b200 import trainer as TR
    from neat_b200.model import VolSDFNetwork
    torch.manual_seed(42)
    model = VolSDFNetwork(synth.dtu_conf())
    with torch.no_grad():
        model.density.beta.fill_(args.beta)
    model = model.to(dev).eval()
    hb = TR.host_batch(args.rays, seed=1 + rank)
    inp, _ = TR.to_device(hb, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        model(inp)
    barrier()
    l0 = lib.neat_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = model(inp)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        i2, _ = TR.to_device(hb, dev)
        o = model(i2)
        host = o["lines3d"].cpu()  # the result a caller keeps (neat-final-parsing.py:213)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms, ms_e2e = float(t[0]), float(t[1])
        print(json.dumps({"metric": "eval_forward_rays_per_sec", "value": world * args.rays * args.steps / (ms * 1e-3),
                          "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "bf16x3->f32", "data": "synthetic",
                          "config": {"workload": "eval-mode forward, %d rays/GPU per call x 98 samples, DTU nets" % args.rays,
                                     "beta": args.beta, "parallelism": "dp%d" % world},
                          "e2e": {"value": world * args.rays * args.steps / (ms_e2e * 1e-3), "unit": "rays/s",
                                  "h2d_bytes_per_step": sum(hb[k].numel() * 4 for k in ("intrinsics", "pose", "uv", "uv_proj")),
                                  "d2h_bytes_per_step": int(host.numel() * 4)},
                          "gpu_launches": int(lib.neat_launch_count() - l0)}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rays", type=int, default=1024, help="rays per GPU per step (weak scaling)")
    ap.add_argument("--beta", type=float, default=0.1, help="density.beta (0.1 = init, k~2; 0.01 = trained-like, k=5)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="train", choices=["train", "eval"],
                    help="train: the headline train step; eval: the eval-mode forward over a chunk of `--rays` rays "
                         "(BASELINE configs[4]: full-image inference, chunked as neat-final-parsing.py / eval.py do)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "DTU-shaped synthetic batch: %d rays/GPU x 98 samples, 8x256 SDF + 4x256 rendering/attraction "
                          "MLPs, ErrorBoundSampler (<=5 x 128 SDF queries/ray), train step = fwd+loss+bwd+Adam" % args.rays,
              "optimizer": "Adam(lr=5e-4), all parameter tensors in one launch (neat_b200.optim.Adam)",
              "rays_per_gpu": args.rays, "samples_per_ray": S, "beta": args.beta, "parallelism": "dp%d" % world,
              "rng": "training draws (stratified jitter, inverse-CDF u, extra columns, eikonal points) made on the device",
              "precision_mode": "bf16x3 (hi/lo split operands, fp32 accumulate) on tcgen05",
              "l2": "no explicit flush: the per-step working set (~5 GB of saved activations at 1024 rays) is >> the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        steps = max(1, min(args.steps, 2))
        val, sec, k = cpu_baseline(args.cpu_rays, args.beta, steps, threads)
        sample = "%d rays (of %d) per step, %d timed step(s) after 1 untimed, sampler k=%d" % (args.cpu_rays, args.rays, steps, k)
        print(json.dumps({"impl": "reference", "metric": "train_step_rays_per_sec", "value": val, "unit": "rays/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from neat_b200 import _lib, synth
    from neat_b200 import trainer as TR
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    if args.mode == "eval":
        return eval_mode(args, rank, world, dev, lib)
    ts = TR.TrainStep(synth.dtu_conf(), device=dev, seed=42, beta=args.beta)
    hb = TR.host_batch(args.rays, seed=1 + rank)
    inp, gt = TR.to_device(hb, dev)
    rn = ts.model._get_renderer()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        ts.step(inp, gt)
    barrier()

    # ---- device-resident inputs ------------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    rn.timers = {}
    l0 = lib.neat_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k_acc = torch.zeros(1, dtype=torch.int32, device=dev)  # the sampler's k is data dependent and drifts as beta trains
    for _ in range(args.steps):
        ts.step(inp, gt)
        k_acc += ts.model.last_step.n_iters
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = lib.neat_launch_count() - l0
    timers = rn.timer_ms()
    rn.timers = None
    k_iters = int(ts.model.last_step.n_iters.item())
    k_mean = float(k_acc.item()) / args.steps

    # ---- end to end from host buffers (same initial state and trajectory as the loop above) ----------------
    del ts
    ts = TR.TrainStep(synth.dtu_conf(), device=dev, seed=42, beta=args.beta)
    for _ in range(max(args.warmup, 3)):
        ts.step(inp, gt)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    loss_host = 0.0
    for _ in range(args.steps):
        i2, g2 = TR.to_device(hb, dev)
        loss_host = float(ts.step(i2, g2).item())
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clk = clocks.stop() if rank == 0 else None

    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms_total / args.steps, float(k_iters), sum(v[1] for v in timers.values()) / args.steps],
                            device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"ms_per_step": round(float(a[0]), 3), "sampler_k": int(a[1]), "mlp_kernel_ms": round(float(a[2]), 3)}
                    for a in allr]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        pk, pk_kind = peaks()
        ms_step = ms_total / args.steps
        value = world * args.rays * args.steps / (ms_total * 1e-3)
        e2e = world * args.rays * args.steps / (ms_e2e * 1e-3)
        M = args.rays * S
        # Tensor work per point: forward F, normal pass F (sdf_render); tangent F + reverse F (sdf_bwd); the two
        # outer-product accumulations 2 F plus the heads' (wgrad) -- SURVEY.md Appendix A, 6 F_sdf per render point.
        flops = {"sdf_bwd_M%d" % M: 2 * F_SDF * M, "sdf_render_M%d" % M: 2 * F_SDF * M,
                 "sampler": 128.0 * k_mean * args.rays * F_SDF,
                 "head_fwd": 0.5 * (F_REND + F_ATT) * M, "head_bwd": 0.5 * (F_REND + F_ATT) * M,  # per launch (one head)
                 "wgrad": (2 * F_SDF + F_REND + F_ATT) * M + 2 * F_SDF * 2 * args.rays}
        # The training kernels are bound by HBM at least as much as by the tensor pipe (ncu: wgrad DRAM 62 % / tensor
        # 27 %, sdf_bwd 58 % / 20 %, sdf_render 35 % / 24 % of peak), so each of them gets BOTH fractions and the larger one
        # names the bound.  Algorithmic bytes = compulsory HBM traffic per 128-point tile:
        #   wgrad      reads every saved operand tile once: 33 main (128 KB: 256 columns x 128 points x bf16 hi + lo) + 5
        #              aux (24 KB) for the SDF net, 2 x (9 main + 2 aux) for the heads = 52.7 KB per render point
        #   sdf_bwd    reads sigma' twice (tangent + reverse sweep, 8 x 128 KB each) + a_l (8 main) + feat_bar (128 KB),
        #              writes p (8 main + aux) and z_bar (9 main + aux) = 42.4 KB per point (the zhat scratch is L2 traffic)
        #   sdf_render writes sigma' (8 x 128 KB), u (8 main), a (8 main), PE (aux), the feature tile = 25.2 KB per point
        MAIN, AUX = 131072.0, 24576.0
        WG_SDF_B, WG_HEAD_B = (33 * MAIN + 5 * AUX) / 128.0, 2 * (9 * MAIN + 2 * AUX) / 128.0
        BWD_B = (16 * MAIN + 8 * MAIN + MAIN + (8 * MAIN + AUX) + (9 * MAIN + AUX)) / 128.0 + 20.0
        REND_B = (8 * MAIN + 8 * MAIN + 8 * MAIN + AUX + MAIN) / 128.0 + 20.0
        hbm_bytes = {"wgrad": (WG_SDF_B + WG_HEAD_B) * M + WG_SDF_B * 2 * args.rays,
                     "sdf_bwd_M%d" % M: BWD_B * M, "sdf_render_M%d" % M: REND_B * M}
        shares = {k: v[1] / ms_total for k, v in timers.items()}
        # dominant KERNEL = longest single launch ("sampler" is a group of up to 5 query launches + the per-ray kernels)
        dom = max((k for k in timers if k in flops and k != "sampler"), key=lambda k: timers[k][1] / timers[k][0])
        n_l, ms_dom = timers[dom]
        sec_dom = ms_dom / n_l * 1e-3
        tensor_tflops = flops[dom] / sec_dom / 1e12
        tensor_frac = tensor_tflops / pk["bf16_tflops_sustained"]
        hbm_gbs = hbm_bytes[dom] / sec_dom / 1e9 if dom in hbm_bytes else 0.0
        hbm_frac = hbm_gbs / pk["hbm_gbs"]
        if hbm_frac >= tensor_frac:
            achieved, peak, unit, bound = hbm_gbs, pk["hbm_gbs"], "GB/s", "hbm"
            peak_kind = pk_kind + " copy bandwidth (read + write)"
            note = ("achieved = ALGORITHMIC bytes (compulsory HBM traffic of the kernel, each saved tile moved once) / "
                    "CUDA-event time; `traffic` = DRAM bytes of one launch from ncu.  The same launch does %.1f algorithmic "
                    "TFLOP/s = %.3f of the sustained bf16 tensor rate (x3 issued: bf16x3)" % (tensor_tflops, tensor_frac))
        else:
            achieved, peak, unit, bound = tensor_tflops, pk["bf16_tflops_sustained"], "TFLOP/s", "tensor"
            peak_kind = pk_kind + " sustained bf16 (cuBLAS)"
            note = ("achieved = ALGORITHMIC fp32-equivalent FLOPs (2*MAC) / CUDA-event time; the kernel issues 3 bf16 MMAs "
                    "per algorithmic MAC (hi*hi + hi*lo + lo*hi) to meet the 1e-4 parity bound, so the tensor pipe does 3x this")
        ncu_traffic = NCU_DRAM_BYTES_1024.get(dom.split("_M")[0]) if args.rays == 1024 else None
        line = {"metric": "train_step_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16x3->f32", "data": "synthetic", "config": config,
                "sampler_iters_k": k_iters, "sampler_iters_k_mean": round(k_mean, 3),
                "mlp_samples_per_sec": world * (M + 128 * k_mean * args.rays + 3 * args.rays) * args.steps / (ms_total * 1e-3),
                "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": TR.h2d_bytes(hb), "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps, "last_loss": loss_host},
                "gpu_launches": int(launches), "clocks": clk,
                "roofline": {"bound": bound, "kernel": dom, "achieved": achieved, "peak": peak, "unit": unit,
                             "frac": achieved / peak,
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at 1024 rays,
                             # from the ncu --set full capture summarised in profiles/r01_v5_ncu_full_summary.csv
                             "traffic": ncu_traffic,
                             "other_bound": {"tensor_frac": round(tensor_frac, 4), "hbm_frac": round(hbm_frac, 4)},
                             "peak_kind": peak_kind, "note": note,
                             "tensor_tflops_algorithmic": {k: round(flops[k] / (timers[k][1] / timers[k][0] * 1e-3) / 1e12, 1)
                                                           for k in timers if k in flops}},
                "kernel_time_share": {k: round(v, 4) for k, v in sorted(shares.items(), key=lambda kv: -kv[1])},
                "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in timers.items()},
                "host_junction_block": getattr(ts.model, "last_host_ms", None), "per_rank": per_rank}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            val, sec, kc = cpu_baseline(args.cpu_rays, args.beta, 1, threads)
            line["cpu_baseline"] = {"value": val, "unit": "rays/s", "cores": threads, "kind": "port",
                                    "sample": "%d rays (of %d) per step, 1 timed step after 1 untimed, sampler k=%d, "
                                              "torch CPU fp32, all host threads" % (args.cpu_rays, args.rays, kc)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
