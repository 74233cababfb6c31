"""One training step as the reference trainer runs it (code/training/volsdf_train.py:361-374):
model(input) -> loss(...) -> zero_grad -> backward -> [all-reduce] -> optimizer.step."""
import torch

from . import synth
from .loss import VolSDFLoss
from .model import VolSDFNetwork
from .optim import Adam
from .parallel import GradBucket


class Wireframe:
    """The members of utils.hawp_util.WireframeGraph the path touches (code/utils/hawp_util.py:7-95): `vertices` [J,2]
    (the model's junction block, the loss) and `line_segments()` [E,5] (dataset lines, finalisation)."""

    def __init__(self, vertices, edges=None, weights=None):
        self.vertices = torch.as_tensor(vertices, dtype=torch.float32)
        self.edges = None if edges is None else torch.as_tensor(edges, dtype=torch.long)
        self.weights = None if weights is None else torch.as_tensor(weights, dtype=torch.float32)

    def line_segments(self, threshold=0.05, device=None):
        """(x1, y1, x2, y2, weight) of the edges with weight > threshold (hawp_util.py:56-69)."""
        ok = self.weights > threshold
        lines = torch.cat((self.vertices[self.edges[ok, 0]], self.vertices[self.edges[ok, 1]], self.weights[ok][:, None]), dim=-1)
        return lines if device is None else lines.to(device)


def host_batch(R, seed, pinned=True, camera_seed=1):
    """A synthetic DTU-shaped batch as HOST tensors, shaped like the reference's
    dataloader output (scene_hawp_dataset.py:148-194).  `seed` draws the pixels / colours / wireframe; the camera
    is the one of `camera_seed` so that every data-parallel rank sees statistically identical work."""
    import math
    a = 0.3 + 0.7 * camera_seed
    pose = synth.look_at_pose((2.5 * math.cos(a) * 0.9, 2.5 * math.sin(a) * 0.9, 2.5 * 0.436))
    b = synth.make_batch(R, seed=seed, pose=pose)
    t = {k: torch.from_numpy(b[k]) for k in ("intrinsics", "pose", "uv", "uv_proj", "rgb", "lines2d")}
    if pinned and torch.cuda.is_available():
        t = {k: v.pin_memory() for k, v in t.items()}
    t["wireframe"] = [Wireframe(b["wf_vertices"], b["wf_edges"], b["wf_weights"])]
    return t


def to_device(hb, dev):
    inp = {k: hb[k].to(dev, non_blocking=True) for k in ("intrinsics", "pose", "uv", "uv_proj")}
    inp["wireframe"] = hb["wireframe"]
    gt = {"rgb": hb["rgb"].to(dev, non_blocking=True), "lines2d": hb["lines2d"].to(dev, non_blocking=True)}
    return inp, gt


def h2d_bytes(hb):
    return sum(hb[k].numel() * hb[k].element_size() for k in ("intrinsics", "pose", "uv", "uv_proj", "rgb", "lines2d"))


class FusedTrainStep:
    """The loop body of code/training/volsdf_train.py:361-374 -- model(input), loss, zero_grad, backward, [all-reduce],
    Adam -- as TWO CUDA-graph replays and one host hand-over per step, with no autograd, no eager PyTorch kernel and no
    host synchronisation other than the junction matching the algorithm itself needs:

        graph A : zero the gradient bucket, weight_norm + packing, every random draw (one Philox launch), rays, the
                  error-bound sampler, junction ffn, SDF / heads / compositing, DBSCAN, the pinned-host copy of the
                  junction candidates, the rest of the forward, the fused loss (+ its gradient), and the WHOLE backward of
                  the three MLPs.  (The junction terms only reach ffn / latents, neat_wfr_rend_a.py:463-466: the backward
                  of everything else does not wait for the host.)
        host    : polls the pinned hand-over flag, solves the two assignment problems (csrc/junction.cpp, ~0.2 ms) while
                  the GPU runs the backward, writes one pinned packet
        graph B : packet H2D, matched-junction terms + their adjoint through the projections and the ffn, [NCCL
                  all-reduce of the flat bucket], Adam for every tensor (step count and lr read from device memory)

    Same arithmetic, kernels and save records as the plugin path (VolSDFNetwork.forward + VolSDFLoss + optim.Adam):
    tests/test_gpu_fused.py checks the two against each other.  graphs=False runs the same launch sequences eagerly."""

    def __init__(self, conf=None, device="cuda:0", seed=42, beta=None, lr=5.0e-4, graphs=True, capture_allreduce=True):
        import torch.distributed as dist
        base = TrainStep(conf, device=device, seed=seed, beta=beta, lr=lr, rng="device")
        self.model, self.loss_fn, self.bucket, self.opt = base.model, base.loss_fn, base.bucket, base.opt
        self.device = torch.device(device)
        self.graphs, self.capture_allreduce = graphs, capture_allreduce
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.n_steps = 0
        self._R = None
        self.gA = self.gB = None
        self.launches_per_step = None
        self.profile = None   # bench.py: a list -> (start, after graph A, after graph B) CUDA events of every step

    # ------------------------------------------------------------------ static state
    def _setup(self, R, n_gt):
        import ctypes
        from . import _lib
        from . import ffn as F
        dev, m = self.device, self.model
        self.lib = _lib.load()
        self.rn = rn = m._get_renderer()
        self._R = R
        f = lambda *s: torch.zeros(*s, device=dev)
        self.inp = dict(intrinsics=f(1, 4, 4), pose=f(1, 4, 4), uv=f(1, R, 2), uv_proj=f(1, R, 2))
        self.gt = dict(rgb=f(1, R, 3), lines2d=f(1, R, 5))
        # junction ffn: parameters, activations, adjoint scratch
        self.lin = F.linears(m.ffn)
        G, H = m.latents.shape
        self.G = G
        self.acts = [f(G, l.weight.shape[0]) for l in self.lin]
        width = max([H] + [l.weight.shape[1] for l in self.lin])
        self.ffn_scratch = [f(G * width), f(G * width)]
        self.j2g, self.j2gc = f(G, 2), f(G, 2)
        self.g_j3g, self.g_j2gc, self.g_glob = f(G, 3), f(G, 2), f(G, 3)
        # the host -> device packet of the junction block: [n | rows | cols | local [cap,7]]
        self.cap = max(1, min(n_gt, 3 * R))   # local junctions <= min(ground-truth junctions, candidates)
        n_words = 1 + 2 * self.cap + 7 * self.cap
        self.packet_host = torch.zeros(n_words, dtype=torch.int32, pin_memory=True)
        self.packet_dev = torch.zeros(n_words, dtype=torch.int32, device=dev)
        self.jout = f(3)
        # fused loss
        self.loss_scratch, self.loss_out = f(8 + R), f(8)
        self.n_eik = 2 * R + (G if m.junction_eikonal else 0)   # eikonal points (neat_wfr_rend_a.py:515-525)
        self.g_rgb, self.g_calib, self.g_theta, self.l3b = f(R, 3), f(R, 4), f(self.n_eik, 3), f(R, 6)
        self.total = f(1)
        # optimizer: torch.optim.Adam's state layout, step count / hyper-parameters on the device
        self.params = [p for p in m.parameters() if p.requires_grad]
        for p in self.params:
            stt = self.opt.state[p]
            if not stt:
                stt["step"] = torch.zeros((), dtype=torch.float32)
                stt["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                stt["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if len(self.params) > 128:
            raise _lib.NeatError("FusedTrainStep: at most 128 parameter tensors")
        self.adam_table = (_lib.AdamTensor * len(self.params))()
        P = ctypes.c_void_p
        for i, p in enumerate(self.params):
            stt = self.opt.state[p]
            self.adam_table[i] = _lib.AdamTensor(P(p.data_ptr()), P(p.grad.data_ptr()), P(stt["exp_avg"].data_ptr()),
                                                 P(stt["exp_avg_sq"].data_ptr()), p.numel())
        # hyper-parameters of the step (lr, betas, eps, weight decay, 1/world): a ring of pinned slots, copied to the device
        # BEFORE graph A of their step.  (One slot read by a copy inside graph B raced with the host, which is already
        # writing the next step's values while graph B of the previous step is still queued behind graph A.)
        self.hyper_host = torch.zeros(4, 6, pin_memory=True)
        self.hyper_dev, self.adam_state = f(6), f(3)
        self.adam_state[0] = float(self.n_steps)

    def _stream(self):
        import ctypes
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ launch sequence A
    def _run_A(self):
        import ctypes
        from . import _lib
        from . import ffn as F
        from .autograd import StepState, step_forward, step_backward
        m, rn, lib, P = self.model, self.rn, self.lib, ctypes.c_void_p
        ptr = lambda t: P(t.data_ptr())
        R = self._R
        self.bucket.zero()
        st = StepState()
        st.uv, st.uv_proj = self.inp["uv"][0], self.inp["uv_proj"][0]
        st.pose, st.K = self.inp["pose"][0], self.inp["intrinsics"][0]
        st.param_layers = m._wn_layers()
        rn.sampler.rng = "device"
        # the junction ffn only feeds the hand-over (side stream, after the attraction head), the projections of the
        # global junctions below and, with junction_eikonal, the eikonal points: it runs on the side stream, off the
        # critical path of the sampler
        main, side = torch.cuda.current_stream(self.device), rn.side_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        with torch.cuda.stream(side):
            side.wait_event(fork)
            glob = F.forward(m.latents.detach(), [l.weight.detach() for l in self.lin], [l.bias.detach() for l in self.lin],
                             self.acts)
            glob_ready = torch.cuda.Event()
            glob_ready.record(side)
        if m.junction_eikonal:
            main.wait_event(glob_ready)
        st.junction_inputs = (glob, st.pose, st.K)
        st.dbscan_enabled = m.dbscan_enabled
        st.junction_eikonal = m.junction_eikonal
        st.use_l3d = m.use_l3d
        st.handover_counter = rn.draw_counter[:1]
        beta = m.density.beta.detach().reshape(1)
        rgb_values, lines3d, grad_theta = step_forward(rn, st, beta)
        self.st = m.last_step = st
        pose_inv = st.pose_inv.reshape(-1)
        main.wait_event(glob_ready)
        _lib.check(lib.neat_project_points(self.G, ptr(pose_inv), ptr(st.K), 4, ptr(glob), ptr(self.j2g), ptr(self.j2gc),
                                           self._stream()))
        # VolSDFLoss core terms fused with their own gradient (loss_wfr.py:47-79), then the adjoint of project2D(I, ...)
        lf = self.loss_fn
        a = _lib.LossArgs(R, self.n_eik, ptr(rgb_values), ptr(self.gt["rgb"]), ptr(st.lines2d), ptr(st.lines2d_calib),
                          ptr(self.gt["lines2d"]), None, ptr(st.K), 4, ptr(grad_theta), float(lf.eikonal_weight),
                          float(lf.line_weight), ptr(self.loss_scratch), ptr(self.loss_out), ptr(self.g_rgb), ptr(self.g_calib),
                          ptr(self.g_theta))
        _lib.check(lib.neat_loss_forward_backward(ctypes.byref(a), self._stream()))
        _lib.check(lib.neat_project_calib_backward(R, ptr(pose_inv), ptr(lines3d), ptr(self.g_calib), ptr(self.l3b),
                                                   self._stream()))
        targets = [tuple(None if p is None else p.grad for p in lay) for lay in st.param_layers]
        step_backward(rn, st, self.g_rgb, self.l3b, self.g_theta, m.density.beta.grad.reshape(1), targets, True)
        self.out = dict(rgb_values=rgb_values, lines3d=lines3d.view(R, 2, 3), grad_theta=grad_theta, j3d_global=glob)

    # ------------------------------------------------------------------ host hand-over
    def _host_junctions(self, wireframe, expect):
        import time
        import numpy as np
        from . import _lib, junction
        flag = self.rn.handover_flag.numpy()
        t0 = time.perf_counter()
        while int(flag[0]) != expect:       # written by the last device->host copy of the hand-over (graph A)
            if time.perf_counter() - t0 > 30.0:
                raise _lib.NeatError("junction hand-over: the device never delivered step %d (flag %d)" % (expect, int(flag[0])))
        t1 = time.perf_counter()
        n_h, cent_h, glob_h, pose_h, K_h = self.st.junction_host
        C = int(n_h[0])
        gt = wireframe.vertices.detach().cpu().numpy()
        local, b0, b1, n_close, med = junction.junction_match(cent_h[:C], gt, pose_h, K_h, glob_h, self.model.use_median)
        n = min(local.shape[0], self.cap)          # matched local junctions
        npair = min(len(b0), n)                    # pairs of the second assignment (<= number of global junctions)
        pk = self.packet_host.numpy()
        pk[0] = npair
        pk[1:1 + npair] = b0[:npair]
        pk[1 + self.cap:1 + self.cap + npair] = b1[:npair]
        pk[1 + 2 * self.cap:1 + 2 * self.cap + 7 * n].view(np.float32)[:] = local[:n].reshape(-1)
        self.last_host_ms = {"wait_for_gpu": (t1 - t0) * 1e3, "junction_host": (time.perf_counter() - t1) * 1e3,
                             "clusters": C, "gt_junctions": int(gt.shape[0]), "matched": int(n), "jcount": int(n_close)}
        self.model.last_host_ms = self.last_host_ms

    # ------------------------------------------------------------------ launch sequence B
    def _run_B(self):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        from . import ffn as F
        m, lib, P = self.model, self.lib, ctypes.c_void_p
        ptr = lambda t: P(t.data_ptr())
        st, lf, glob = self.st, self.loss_fn, self.out["j3d_global"]
        self.packet_dev.copy_(self.packet_host, non_blocking=True)
        _lib.check(lib.neat_junction_step(ptr(self.packet_dev), self.cap, self.G, ptr(glob), ptr(self.j2gc), ptr(self.j2g),
                                          float(lf.junction_3d_weight), float(lf.junction_2d_weight), ptr(self.jout),
                                          ptr(self.g_j3g), ptr(self.g_j2gc), self._stream()))
        _lib.check(lib.neat_project_points_backward(self.G, ptr(st.pose_inv.reshape(-1)), ptr(st.K), 4, ptr(glob), None,
                                                    ptr(self.g_j2gc), ptr(self.g_glob), self._stream()))
        self.g_glob.add_(self.g_j3g)
        F.backward(m.latents.detach(), [l.weight.detach() for l in self.lin], self.acts, self.g_glob, m.latents.grad,
                   [l.weight.grad for l in self.lin], [l.bias.grad for l in self.lin], self.ffn_scratch, True,
                   side=self.rn.side_stream())
        if self.world > 1:
            dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM)
        _lib.check(lib.neat_adam_step_device(self.adam_table, len(self.params), ptr(self.hyper_dev), ptr(self.adam_state),
                                             self._stream()))
        # loss = core + w3 j3d + w2 j2d (loss_wfr.py:82-125)
        torch.add(self.loss_out[0:1], self.jout[0:1], alpha=float(lf.junction_3d_weight), out=self.total)
        self.total.add_(self.jout[1:2], alpha=float(lf.junction_2d_weight))

    # ------------------------------------------------------------------ one step
    def step(self, inp, gt):
        """inp / gt: the trainer's dicts (host or device tensors; volsdf_train.py:361-366).  Returns the loss dict of
        VolSDFLoss (device scalars; `loss` is a [1] tensor valid until the next step)."""
        R = inp["uv"].reshape(-1, 2).shape[0]
        wf = inp["wireframe"][0]
        if self._R != R:
            self._setup(R, int(wf.vertices.shape[0]))
            self.gA = self.gB = None
            self._eager_done = 0
        for k in ("intrinsics", "pose", "uv", "uv_proj"):
            self.inp[k].copy_(inp[k].reshape(self.inp[k].shape), non_blocking=True)
        for k in ("rgb", "lines2d"):
            self.gt[k].copy_(gt[k].reshape(self.gt[k].shape), non_blocking=True)
        g = self.opt.param_groups[0]
        slot = self.hyper_host[self.n_steps % self.hyper_host.shape[0]]
        h = slot.numpy()
        h[0], h[1], h[2], h[3], h[4], h[5] = g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], 1.0 / self.world
        self.hyper_dev.copy_(slot, non_blocking=True)   # stream-ordered: after graph B of the previous step, before this step
        use_graphs = self.graphs and self._eager_done >= 2
        if use_graphs and self.gA is None:
            self._capture()
        expect = int(self._counter_base + self._replays + 1) if use_graphs else None
        ev = None
        if self.profile is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        if use_graphs:
            self.gA.replay()
            self._replays += 1
        else:
            self._run_A()
            torch.cuda.current_stream(self.device).synchronize()
            expect = int(self.rn.handover_flag[0])
            self._eager_done += 1
        if ev is not None:
            ev[1].record()
        self._host_junctions(wf, expect)
        if use_graphs:
            self.gB.replay()
        else:
            self._run_B()
        if ev is not None:
            ev[2].record()
            self.profile.append(ev)
        self.n_steps += 1
        self.model._packed_version = None   # the parameters changed under raw pointers: eval entry points must re-pack
        lo = self.loss_out
        return {"loss": self.total, "rgb_loss": lo[1], "eikonal_loss": lo[2], "line_loss": lo[3], "l2d_loss": lo[4],
                "count": lo[5], "j3d_loss": self.jout[0], "j2d_loss": self.jout[1], "j2d_stat": self.jout[2],
                "jcount": self.last_host_ms["jcount"]}

    def _capture(self):
        from . import _lib
        dev = self.device
        torch.cuda.synchronize(dev)
        l0 = self.lib.neat_launch_count()
        self.gA, self.gB = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.gA):
            self._run_A()
        lA = self.lib.neat_launch_count() - l0
        try:
            with torch.cuda.graph(self.gB, pool=self.gA.pool()):
                self._run_B()
        except Exception as e:
            raise _lib.NeatError("FusedTrainStep: capturing the optimizer graph failed (%s: %s); pass graphs=False" %
                                 (type(e).__name__, e))
        self.launches_per_step = int(self.lib.neat_launch_count() - l0)
        self.launches_A = int(lA)
        torch.cuda.synchronize(dev)
        self._counter_base = int(self.rn.draw_counter[0].item())
        self._replays = 0

    def profile_ms(self):
        """Mean per-step (graph A, graph B, idle gap before the next step) device times of the profiled steps.  Graph B
        holds the gradient all-reduce, so at N > 1 its time includes waiting for the slowest rank."""
        ev = self.profile or []
        if len(ev) < 2:
            return None
        n = len(ev)
        a = sum(e[0].elapsed_time(e[1]) for e in ev) / n
        b = sum(e[1].elapsed_time(e[2]) for e in ev) / n
        gap = sum(ev[i][2].elapsed_time(ev[i + 1][0]) for i in range(n - 1)) / (n - 1)
        return {"graph_A_ms": round(a, 4), "graph_B_ms": round(b, 4), "gap_ms": round(gap, 4)}

    def sync_optimizer_state(self):
        """Bring the host-side `step` entries of the torch.optim state (state_dict / checkpoints) up to date."""
        for p in self.params:
            self.opt.state[p]["step"].fill_(float(self.n_steps))


class TrainStep:
    def __init__(self, conf=None, device="cuda:0", seed=42, beta=None, lr=5.0e-4, rng="device"):
        conf = conf or synth.dtu_conf()
        torch.manual_seed(seed)
        self.model = VolSDFNetwork(conf)
        if beta is not None:
            with torch.no_grad():
                self.model.density.beta.fill_(beta)
        self.model = self.model.to(device).train()
        self.model.rng = rng
        self.loss_fn = VolSDFLoss(**synth.loss_conf())
        self.bucket = GradBucket(self.model.parameters())
        # same update rule as the reference's torch.optim.Adam(lr) (volsdf_train.py:178), all tensors in one launch
        self.opt = Adam(self.model.parameters(), lr=lr)

    def step(self, inp, gt):
        out = self.model(inp)
        lo = self.loss_fn(out, gt)
        self.bucket.zero()
        lo["loss"].backward()
        self.opt.grad_scale = self.bucket.all_reduce_sum()  # the 1/world scale rides in the Adam kernel
        self.opt.step()
        return lo["loss"]
