"""Forward orchestration of VolSDFNetwork.forward (code/model/networks/neat_wfr_rend_a.py:376-538) on the
C-ABI kernels.  Only device-memory allocation and launch ordering happen here."""
import ctypes

import torch

from . import _lib
from .context import Context, ErrorBoundSampler

_P = ctypes.c_void_p


def _ptr(t):
    return _P(t.data_ptr()) if t is not None else None


class _Timed:
    def __init__(self, rn, name):
        self.rn, self.name = rn, name

    def __enter__(self):
        if self.rn.timers is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.rn.ctx.device))

    def __exit__(self, *exc):
        if self.rn.timers is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.rn.ctx.device))
            self.rn.timers.setdefault(self.name, []).append((self.e0, e1))


class WorkspacePool:
    """Named, persistent device buffers.  The per-step save records are GB-sized; asking the caching allocator for
    them every step ends in cudaMalloc/cudaFree.  One step is in flight at a time, so each named buffer is simply
    reused (it is only re-allocated when a larger size is requested)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, numel, dtype=torch.float32):
        key = (name, dtype)
        t = self.bufs.get(key)
        if t is None or t.numel() < numel:
            self.bufs[key] = t = torch.empty(max(int(numel), 1), dtype=dtype, device=self.device)
        return t[:numel]

    def zeros(self, name, numel, dtype=torch.float32):
        t = self.get(name, numel, dtype)
        t.zero_()
        return t

    def total_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class Renderer:
    def __init__(self, ctx: Context, conf):
        self.ctx = ctx
        self.conf = conf
        self.sampler = ErrorBoundSampler(ctx, conf)
        self.sampler.renderer = self
        self.beta_min = float(conf["density"].get("beta_min", 1e-4))
        self.scene_bounding_sphere = float(conf.get("scene_bounding_sphere", 1.0))
        self.pool = WorkspacePool(ctx.device)
        self._pinned = {}
        # white_bkgd (neat_wfr_rend_a.py:262-264, 411-413): rgb_values += (1 - sum w) * bg_color; None = off
        self.bg_color = None
        if bool(conf.get("white_bkgd", False)):
            self.bg_color = torch.tensor([float(v) for v in conf.get("bg_color", [1.0, 1.0, 1.0])], dtype=torch.float32,
                                         device=ctx.device)
        self.timers = None  # bench.py: dict name -> [(start, end) CUDA events] on the launching stream
        self.generation = 0  # bumped by every forward that reuses the named workspaces (autograd.py checks it in backward)
        # device-side random draws of the training forward (csrc/train_aux.cuh): Philox keyed by (seed, a draw counter that
        # lives on the device and is bumped by the kernel: replayable inside a CUDA graph)
        self.draw_seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self.draw_counter = torch.zeros(2, dtype=torch.int64, device=ctx.device)

    def seed_draws(self, seed):
        """Restart the device-side draw sequence (same seed => the same samples, eikonal points, extra columns)."""
        self.draw_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.draw_counter.zero_()

    def draws(self, R):
        """All random draws of one training forward in one launch (neat_train_draws): dict of device tensors."""
        ctx, c, pool = self.ctx, self.sampler.cfg, self.pool
        d = dict(t_rand=pool.get("draw.t_rand", R * c.n_eval).view(R, c.n_eval),
                 u_final=pool.get("draw.u_final", R * c.n_final).view(R, c.n_final),
                 extra_table=pool.get("draw.extra", c.max_iters * max(c.n_extra, 1), torch.int64).view(c.max_iters, -1),
                 eik_idx=pool.get("draw.eik_idx", R, torch.int64),
                 eik_uniform=pool.get("draw.eik_uniform", R * 3).view(R, 3))
        _lib.check(ctx.lib.neat_train_draws(ctypes.byref(c), R, ctypes.c_float(self.scene_bounding_sphere),
                                            ctypes.c_ulonglong(self.draw_seed), _ptr(self.draw_counter), _ptr(d["t_rand"]),
                                            _ptr(d["u_final"]), _ptr(d["extra_table"]), _ptr(d["eik_idx"]),
                                            _ptr(d["eik_uniform"]), ctx._stream()))
        return d

    def timed(self, name):
        return _Timed(self, name)

    def timer_ms(self):
        """name -> (launch groups, total ms); call after a synchronize."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (self.timers or {}).items()}

    # ---- weight_norm + packing of all layers (two launches) --------------------------------------------
    def _wn_table(self, layers, grads=None):
        """layers: [(weight_g | None, weight_v, bias)] in flat-buffer order; grads: same structure (backward)."""
        arr = (_lib.WnLayer * len(layers))()
        for i, (g, v, b) in enumerate(layers):
            gg = gv = gb = None
            if grads is not None:
                gg, gv, gb = grads[i]
            arr[i] = _lib.WnLayer(_ptr(g), _ptr(v), _ptr(b), _ptr(gg), _ptr(gv), _ptr(gb), v.shape[0], v.shape[1], 0, 0)
        return arr

    def effective_weights(self, layers):
        """nn.utils.weight_norm of every layer into the flat parameter buffer, then the tcgen05 operand slabs."""
        ctx = self.ctx
        for g, v, b in layers:
            for t in (g, v, b):
                if t is not None and not (t.is_contiguous() and t.dtype == torch.float32 and t.device == ctx.device):
                    raise _lib.NeatError("parameters must be contiguous fp32 tensors on %s" % ctx.device)
        flat = self.pool.get("params.flat", ctx.n_params)
        arr = self._wn_table(layers)
        _lib.check(ctx.lib.neat_weight_norm_forward(ctx._h, arr, len(layers), _ptr(flat), ctx._stream()))
        ctx.pack_weights(flat)
        return flat

    def weight_norm_backward(self, layers, flat_grad, targets, accumulate):
        """Adjoint of effective_weights: flat_grad (layout of the flat parameter buffer) -> the (weight_g, weight_v,
        bias) gradients, written (accumulate=False) or added (True) straight into `targets` (same structure)."""
        ctx = self.ctx
        arr = self._wn_table(layers, targets)
        _lib.check(ctx.lib.neat_weight_norm_backward(ctx._h, arr, len(layers), _ptr(flat_grad), int(bool(accumulate)),
                                                     ctx._stream()))

    # ---- small helpers over the C entry points ------------------------------------------------
    def camera_rays(self, uv, pose, K):
        ctx = self.ctx
        R = uv.shape[0]
        dirs = torch.empty(R, 3, device=ctx.device)
        cam = torch.empty(3, device=ctx.device)
        _lib.check(ctx.lib.neat_camera_rays(ctx._chk(uv, (R, 2)), ctx._chk(pose, (4, 4)), ctx._chk(K, (4, 4)), R,
                                            _ptr(dirs), _ptr(cam), ctx._stream()))
        return dirs, cam

    def ray_points(self, cam, dirs, z):
        R, S = z.shape
        return _lib.Points(None, None, _ptr(cam), _ptr(dirs), _ptr(z), 0, R, S, R * S)

    def explicit_points(self, x, dirs=None):
        return _lib.Points(_ptr(x), _ptr(dirs), None, None, None, 0, 0, 0, x.shape[0])

    def sdf_outputs(self, pts, M, clamp=True, training=False, want_feat=True, want_sdf=True, tag="sdf"):
        """`tag` names the persistent workspace buffers of this call site (they stay valid until the same tag is
        used again, i.e. through the backward pass of the step)."""
        ctx = self.ctx
        dev = ctx.device
        pool = self.pool
        sdf = torch.empty(M, device=dev) if want_sdf else None
        grad = torch.empty(M, 3, device=dev)
        act = pool.get(tag + ".act", M) if training else None
        feat = pool.get(tag + ".feat", int(ctx.lib.neat_feat_tiles_bytes(M)), torch.uint8) if want_feat else None
        save = pool.get(tag + ".save", int(ctx.lib.neat_sdf_save_bytes(ctx._h, M, int(training))), torch.uint8)
        with self.timed("sdf_render_M%d" % M):
            _lib.check(ctx.lib.neat_sdf_outputs(ctx._h, ctypes.byref(pts), int(clamp), int(training), _ptr(sdf), _ptr(grad),
                                                _ptr(act), _ptr(feat), _ptr(save), ctx._stream()))
        return sdf, grad, act, feat, save

    def head_forward(self, head, pts, M, normals, feat, training=False):
        ctx = self.ctx
        if not training:
            self.generation += 1  # "head%d.out" is also what a pending training step's backward reads (st.rgb / st.lines)
        out = self.pool.get("head%d.out" % head, M * (3 if head == 0 else 6)).view(M, 3 if head == 0 else 6)
        save = (self.pool.get("head%d.save" % head, int(ctx.lib.neat_head_save_bytes(ctx._h, M)), torch.uint8)
                if training else None)
        with self.timed("head_fwd"):
            _lib.check(ctx.lib.neat_head_forward(ctx._h, head, ctypes.byref(pts), _ptr(normals), _ptr(feat), int(training),
                                                 _ptr(save), _ptr(out), ctx._stream()))
        return out, save

    def composite(self, z, sdf, rgb, lines, normals, cam, dirs, beta_param, want_normal_map):
        ctx = self.ctx
        dev = ctx.device
        R, S = z.shape
        w = self.pool.get("composite.w", R * S).view(R, S)
        rgb_values = torch.empty(R, 3, device=dev)
        lines3d = torch.empty(R, 6, device=dev)
        depth = torch.empty(R, device=dev)
        points3d = torch.empty(R, 3, device=dev)
        nmap = torch.empty(R, 3, device=dev) if want_normal_map else None
        a = _lib.CompositeArgs(R, S, _ptr(z), _ptr(sdf), _ptr(rgb), _ptr(lines), _ptr(normals) if want_normal_map else None,
                               _ptr(cam), _ptr(dirs), _ptr(beta_param), self.beta_min, _ptr(w), _ptr(rgb_values),
                               _ptr(lines3d), _ptr(depth), _ptr(points3d), _ptr(nmap), _ptr(self.bg_color))
        _lib.check(ctx.lib.neat_composite_forward(ctypes.byref(a), ctx._stream()))
        return w, rgb_values, lines3d, depth, points3d, nmap

    def composite_lines(self, z, sdf, lines, cam, dirs, beta_param):
        """First half of the training step's compositing: weights, lines3d, depth, points3d (everything the junction
        clustering and the geometry need) as soon as the attraction head is done."""
        ctx = self.ctx
        dev = ctx.device
        R, S = z.shape
        w = self.pool.get("composite.w", R * S).view(R, S)
        lines3d = torch.empty(R, 6, device=dev)
        depth = torch.empty(R, device=dev)
        points3d = torch.empty(R, 3, device=dev)
        a = _lib.CompositeArgs(R, S, _ptr(z), _ptr(sdf), None, _ptr(lines), None, _ptr(cam), _ptr(dirs), _ptr(beta_param),
                               self.beta_min, _ptr(w), None, _ptr(lines3d), _ptr(depth), _ptr(points3d), None)
        _lib.check(ctx.lib.neat_composite_forward(ctypes.byref(a), ctx._stream()))
        return w, lines3d, depth, points3d

    def composite_weights(self, z, sdf, beta_param):
        """volume_rendering alone: the alpha-compositing weights [R,S]."""
        ctx = self.ctx
        R, S = z.shape
        w = torch.empty(R, S, device=ctx.device)
        zero = torch.zeros(R, 3, device=ctx.device)
        a = _lib.CompositeArgs(R, S, _ptr(z), _ptr(sdf), None, None, None, _ptr(zero), _ptr(zero), _ptr(beta_param),
                               self.beta_min, _ptr(w), None, None, None, None, None)
        _lib.check(ctx.lib.neat_composite_forward(ctypes.byref(a), ctx._stream()))
        return w

    def composite_rgb(self, z, sdf, rgb, cam, dirs, beta_param):
        """Second half: rgb_values = sum_i w_i rgb_i once the rendering head is done."""
        ctx = self.ctx
        R, S = z.shape
        rgb_values = torch.empty(R, 3, device=ctx.device)
        a = _lib.CompositeArgs(R, S, _ptr(z), _ptr(sdf), _ptr(rgb), None, None, _ptr(cam), _ptr(dirs), _ptr(beta_param),
                               self.beta_min, None, _ptr(rgb_values), None, None, None, None, _ptr(self.bg_color))
        _lib.check(ctx.lib.neat_composite_forward(ctypes.byref(a), ctx._stream()))
        return rgb_values

    def line_geometry(self, pose, K, uv_proj, points3d, grad3d, lines3d):
        ctx = self.ctx
        dev = ctx.device
        R = uv_proj.shape[0]
        pose_inv = torch.empty(16, device=dev)
        l2d = torch.empty(R, 2, 2, device=dev)
        l2dc = torch.empty(R, 2, 2, device=dev)
        l3d = torch.empty(R, 3, device=dev)
        _lib.check(ctx.lib.neat_line_geometry(R, _ptr(pose), _ptr(K), _ptr(uv_proj), _ptr(points3d), _ptr(grad3d),
                                              _ptr(lines3d), _ptr(pose_inv), _ptr(l2d), _ptr(l2dc), _ptr(l3d),
                                              ctx._stream()))
        return l2d, l2dc, l3d, pose_inv.view(4, 4)

    def dbscan_async(self, points, eps=0.01):
        """Like dbscan() but without the device->host read: returns (centroid buffer [N/2+1,3], device count [1])."""
        ctx = self.ctx
        N = points.shape[0]
        ws = self.pool.get("dbscan.ws", int(ctx.lib.neat_dbscan_workspace_bytes(N)), torch.uint8)
        cent = self.pool.get("dbscan.cent", (N // 2 + 1) * 3).view(-1, 3)
        n = self.pool.get("dbscan.n", 1, torch.int32)
        _lib.check(ctx.lib.neat_dbscan(ctx._chk(points, (N, 3)), N, ctypes.c_float(eps), _ptr(ws), _ptr(cent), _ptr(n),
                                       ctx._stream()))
        return cent, n

    def l3d_candidates_async(self, lines3d, l3d):
        """use_l3d junction candidates (neat_wfr_rend_a.py:461-465) without a device->host read: (buffer [3R,3],
        device count [1])."""
        ctx = self.ctx
        R = l3d.shape[0]
        out = self.pool.get("l3d.cand", 3 * R * 3).view(-1, 3)
        n = self.pool.get("l3d.n", 1, torch.int32)
        _lib.check(ctx.lib.neat_l3d_candidates(R, ctx._chk(lines3d.reshape(R, 6), (R, 6)), ctx._chk(l3d, (R, 3)),
                                               _ptr(out), _ptr(n), None, ctx._stream()))
        return out, n

    def side_stream(self):
        """Second stream for the small launches that fit into the tail of a big persistent one."""
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.ctx.device)
        return self._side

    def pinned(self, name, numel, dtype=torch.float32):
        """A named, persistent pinned host buffer (staging for small host->device hand-overs)."""
        key = (name, dtype)
        h = self._pinned.get(key)
        if h is None or h.numel() < numel:
            self._pinned[key] = h = torch.empty(max(int(numel), 1), dtype=dtype, pin_memory=True)
        return h

    def to_host_async(self, tensors, counter=None):
        """Enqueue copies of several small device tensors into pinned host buffers; returns (event, numpy views).
        The views are valid after event.synchronize() and until the next call.  counter: a device int64 tensor copied
        LAST into its own pinned slot (self.handover_flag): a host that cannot wait on an event (the copies were captured
        into a CUDA graph) polls that slot for the value it expects instead."""
        outs = []
        for i, t in enumerate(tensors):
            key = ("host%d" % i, t.dtype)
            h = self._pinned.get(key)
            if h is None or h.numel() < t.numel():
                self._pinned[key] = h = torch.empty(max(t.numel(), 1), dtype=t.dtype, pin_memory=True)
            hv = h[:t.numel()].view(t.shape)
            hv.copy_(t, non_blocking=True)
            outs.append(hv)
        if counter is not None:
            flag = self.pinned("handover.flag", counter.numel(), counter.dtype)
            flag[:counter.numel()].copy_(counter, non_blocking=True)
            self.handover_flag = flag
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.ctx.device))
        return ev, [o.numpy() for o in outs]

    def to_host(self, tensors):
        """One synchronisation for several small device tensors: async copies into pinned buffers, then one wait."""
        outs = []
        for i, t in enumerate(tensors):
            key = ("host%d" % i, t.dtype)
            h = self._pinned.get(key)
            if h is None or h.numel() < t.numel():
                self._pinned[key] = h = torch.empty(max(t.numel(), 1), dtype=t.dtype, pin_memory=True)
            hv = h[:t.numel()].view(t.shape)
            hv.copy_(t, non_blocking=True)
            outs.append(hv)
        torch.cuda.current_stream(self.ctx.device).synchronize()
        return [o.numpy() for o in outs]

    def dbscan(self, points, eps=0.01):
        """cluster_dbscan (neat_wfr_rend_a.py:333-342) on device: centroids [C,3] of the eps-connected components
        with >= 2 points, ordered like sklearn's labels.  One 4-byte device->host read (C)."""
        ctx = self.ctx
        N = points.shape[0]
        ws = self.pool.get("dbscan.ws", int(ctx.lib.neat_dbscan_workspace_bytes(N)), torch.uint8)
        cent = torch.empty(N // 2 + 1, 3, device=ctx.device)
        n = torch.zeros(1, dtype=torch.int32, device=ctx.device)
        _lib.check(ctx.lib.neat_dbscan(ctx._chk(points, (N, 3)), N, ctypes.c_float(eps), _ptr(ws), _ptr(cent), _ptr(n),
                                       ctx._stream()))
        return cent[:int(n.item())]

    # ---- feature tiles <-> [M, F] (standalone ImplicitNetwork / head entry points only) ------------
    def unpack_features(self, tiles, M):
        F = self.ctx.cfg.feat
        t = tiles.view(torch.float16).view(-1, 2, 32, 128, 8).float()   # forward tiles hold fp16 hi / lo pairs
        full = (t[:, 0] + t[:, 1]).permute(0, 2, 1, 3).reshape(-1, 256)
        return full[:M, :F].contiguous()

    def pack_features(self, feat):
        M, F = feat.shape
        nt = (M + 127) // 128
        full = torch.zeros(nt * 128, 256, device=feat.device)
        full[:M, :F] = feat
        full = full.clamp(-65504.0, 65504.0)
        hi = full.to(torch.float16)
        lo = (full - hi.float()).to(torch.float16)
        t = torch.stack([hi, lo], 0).view(2, nt, 128, 32, 8).permute(1, 0, 3, 2, 4).contiguous()
        return t.view(torch.uint8).reshape(-1)

    # ---- eval-mode forward (no autograd) --------------------------------------------------------
    @torch.no_grad()
    def forward_eval(self, uv, pose, K, uv_proj, beta_param):
        """uv [R,2], pose [4,4], K [4,4], uv_proj [R,2] (device fp32).  Mirrors the eval branch of
        VolSDFNetwork.forward; returns the reference's output dict entries computed on device."""
        self.generation += 1  # shares the "render" / head workspaces with a pending training step, if any
        dirs, cam = self.camera_rays(uv, pose, K)
        z, z_eik, n_it = self.sampler.get_z_vals(cam, dirs, beta_param, training=False)
        self.last_n_iters = n_it
        R, S = z.shape
        M = R * S
        pts = self.ray_points(cam, dirs, z)
        sdf, grad, _, feat, _ = self.sdf_outputs(pts, M, clamp=True, tag="render")
        rgb, _ = self.head_forward(0, pts, M, grad, feat)
        lines, _ = self.head_forward(1, pts, M, grad, feat)
        w, rgb_values, lines3d, depth, points3d, nmap = self.composite(z, sdf, rgb, lines, grad, cam, dirs, beta_param, True)
        p3 = self.explicit_points(points3d)
        sdf3, grad3, _, _, _ = self.sdf_outputs(p3, R, clamp=True, want_feat=False, tag="surface")
        l2d, l2dc, l3d, _ = self.line_geometry(pose, K, uv_proj, points3d, grad3, lines3d)
        # rgb / lines / weights live in the reusable pool: hand out copies of what the caller may keep
        rgb, lines, w = rgb.clone(), lines.clone(), w.clone()
        return dict(points=cam[None, None, :] + z[:, :, None] * dirs[:, None, :], rgb_values=rgb_values, depth=depth,
                    xyz=points3d, points3d=points3d, lines3d=lines3d.view(R, 2, 3), lines2d=l2d, lines2d_calib=l2dc,
                    l3d=l3d, sdf=sdf3, normal_map=nmap, K=K[:3, :3], z_vals=z, weights=w, n_sampler_iters=n_it,
                    sdf_pts=sdf.view(R, S), grad_pts=grad.view(R, S, 3), rgb_pts=rgb.view(R, S, 3),
                    lines_pts=lines.view(R, S, 2, 3))
