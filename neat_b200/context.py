"""Thin Python owner of a `neat_ctx` (include/neat_b200.h): shapes from the reference conf, the flat
effective-parameter buffer, and typed wrappers over the C entry points.  torch is used only for
device memory and the current stream."""
import ctypes
import os

import torch

from . import _lib


def net_config_from_conf(conf):
    """conf: the `model` sub-tree of a NEAT conf (code/confs/dtu.conf:28-87) as nested dict / ConfigTree."""
    ci, cr, ca = conf["implicit_network"], conf["rendering_network"], conf["attraction_network"]
    dims = list(ci["dims"])
    if len(set(dims)) != 1:
        raise _lib.NeatError("ImplicitNetwork dims must be uniform")
    if list(cr["dims"]) != list(ca["dims"]) or len(set(cr["dims"])) != 1:
        raise _lib.NeatError("rendering / attraction dims must be uniform and equal")
    skip = list(ci.get("skip_in", []))
    if len(skip) > 1:
        raise _lib.NeatError("at most one skip connection is supported")
    if cr.get("mode", "idr") != "idr" or ca.get("mode", "idr") != "idr":
        raise _lib.NeatError("only mode='idr' heads are supported")
    return _lib.NetConfig(
        sdf_layers=len(dims) + 1, sdf_hidden=dims[0], sdf_skip=skip[0] if skip else -1,
        multires=int(ci.get("multires", 0)), feat=int(conf["feature_vector_size"]),
        head_layers=len(cr["dims"]) + 1, head_hidden=cr["dims"][0],
        multires_view=int(cr.get("multires_view", 0)),
        # white_bkgd builds the ImplicitNetwork with sdf_bounding_sphere = 0: no sphere clamp (neat_wfr_rend_a.py:266)
        sphere_radius=0.0 if bool(conf.get("white_bkgd", False)) else float(conf.get("scene_bounding_sphere", 1.0)),
        sphere_scale=float(ci.get("sphere_scale", 1.0)))


class Context:
    NETS = ("implicit_network", "rendering_network", "attraction_network")

    def __init__(self, conf, device="cuda:0"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NeatError("neat_b200 runs on CUDA devices only (no CPU fallback)")
        self.cfg = net_config_from_conf(conf)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.neat_create(ctypes.byref(self.cfg), ctypes.byref(self._h)))
            if os.environ.get("NEAT_L2_PREFETCH") is not None:  # A/B knob for profiling only
                _lib.check(self.lib.neat_debug_set_l2_prefetch(int(os.environ["NEAT_L2_PREFETCH"])))
        self.n_params = int(self.lib.neat_param_count(self._h))
        self.layers = []  # (net index, layer, in, out, w_off, b_off)
        for net, n_l in ((0, self.cfg.sdf_layers), (1, self.cfg.head_layers), (2, self.cfg.head_layers)):
            for l in range(n_l):
                i, o = ctypes.c_int(), ctypes.c_int()
                _lib.check(self.lib.neat_layer_dims(self._h, net, l, ctypes.byref(i), ctypes.byref(o)))
                self.layers.append((net, l, i.value, o.value,
                                    int(self.lib.neat_param_offset(self._h, net, l, 0)),
                                    int(self.lib.neat_param_offset(self._h, net, l, 1))))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.neat_destroy(h)

    # ------------------------------------------------------------------ test knobs
    def debug_grid_cap(self, max_ctas):
        """Tests: at most `max_ctas` persistent CTAs per tile-MLP launch (0 = one per SM), so that a small point count
        exercises the several-tiles-per-CTA loops the full-size step runs."""
        _lib.check(self.lib.neat_debug_set_grid_cap(self._h, int(max_ctas)))

    def debug_wgrad_split(self, max_split, tiles_per_split):
        """Tests: split every weight-gradient GEMM's tile range into up to `max_split` pieces of >= `tiles_per_split`
        tiles (0, 0 = the defaults 32 / 48)."""
        _lib.check(self.lib.neat_debug_set_wgrad_split(self._h, int(max_split), int(tiles_per_split)))

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t, shape=None):
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.NeatError("expected a contiguous fp32 tensor on %s" % self.device)
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise _lib.NeatError("expected shape %s, got %s" % (tuple(shape), tuple(t.shape)))
        return ctypes.c_void_p(t.data_ptr())

    def flatten_state_dict(self, sd):
        """Effective weights (weight_norm applied, neat_wfr_rend_a.py:71-72) in the flat layout of
        neat_param_count(); differentiable w.r.t. the tensors in `sd`."""
        parts = []
        for net, l, i, o, _, _ in self.layers:
            pre = "%s.lin%d" % (self.NETS[net], l)
            if pre + ".weight_g" in sd:
                v = sd[pre + ".weight_v"]
                w = sd[pre + ".weight_g"] * v / v.norm(2, dim=1, keepdim=True)
            else:
                w = sd[pre + ".weight"]
            assert tuple(w.shape) == (o, i), (pre, tuple(w.shape), (o, i))
            parts += [w.reshape(-1), sd[pre + ".bias"].reshape(-1)]
        flat = torch.cat(parts).to(self.device, torch.float32).contiguous()
        assert flat.numel() == self.n_params
        return flat

    # ------------------------------------------------------------------ entry points
    def pack_weights(self, flat):
        self._flat = flat  # keep alive until the kernels that read the packed copy ran
        _lib.check(self.lib.neat_pack_weights(self._h, self._chk(flat, (self.n_params,)), self._stream()))

    def sdf_points(self, x):
        M = x.shape[0]
        out = torch.empty(M, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.neat_sdf_points(self._h, self._chk(x, (M, 3)), M, self._chk(out), self._stream()))
        return out

    def sdf_rays(self, rays_o, rays_d, z):
        R, n = z.shape
        o_stride = 3 if rays_o.dim() == 2 and rays_o.shape[0] == R and R > 1 else 0
        if o_stride == 0:
            rays_o = rays_o.reshape(-1)[:3].contiguous()
        out = torch.empty(R, n, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.neat_sdf_rays(self._h, self._chk(rays_o), o_stride, self._chk(rays_d, (R, 3)),
                                          self._chk(z), R, n, self._chk(out), self._stream()))
        return out


def sampler_config_from_conf(conf):
    c = conf["ray_sampler"]
    # options of ErrorBoundSampler (code/model/ray_sampler.py:101-128) the kernels do not implement: refuse, never ignore
    if float(c.get("add_tiny", 0.0)) != 0.0:
        raise _lib.NeatError("ray_sampler.add_tiny != 0 is not supported (the kernels hard-code add_tiny = 0, the value "
                             "of every shipped conf)")
    if bool(c.get("inverse_sphere_bg", False)) or int(c.get("N_samples_inverse_sphere", 0)) > 0:
        raise _lib.NeatError("ray_sampler.inverse_sphere_bg / N_samples_inverse_sphere are not supported (no shipped conf "
                             "enables the inverse-sphere background, ray_sampler.py:261-263)")
    return _lib.SamplerConfig(
        n_eval=int(c["N_samples_eval"]), n_final=int(c["N_samples"]), n_extra=int(c.get("N_samples_extra", 0)),
        beta_iters=int(c["beta_iters"]), max_iters=int(c["max_total_iters"]), near_=float(c["near"]),
        far_=2.0 * float(conf.get("scene_bounding_sphere", 1.0)), eps=float(c["eps"]),
        beta_min=float(conf["density"].get("beta_min", 1e-4)))


class ErrorBoundSampler:
    """ErrorBoundSampler.get_z_vals (code/model/ray_sampler.py:130-283) on the CUDA kernels.

    eval mode: fully asynchronous.  training mode replays the reference's CPU-generator draws in
    order (rand[R,n_eval], randint[R] (unused), rand[R,n_final], randperm(L), randint[R]); randperm
    needs L = n_eval * k, so it costs the one device->host read of k that the reference pays k times."""

    def __init__(self, ctx, conf):
        self.ctx = ctx
        self.cfg = sampler_config_from_conf(conf)
        c = self.cfg
        self.n_out = c.n_final + 2 + c.n_extra
        tab = [torch.linspace(0, c.n_eval * k - 1, c.n_extra).long() for k in range(1, c.max_iters + 1)]
        self.eval_table = torch.stack(tab).to(ctx.device) if c.n_extra > 0 else None
        self._ws = None
        self._ws_R = -1

    def workspace(self, R):
        if self._ws_R < R:  # persistent: re-allocated only when the ray count grows
            n = int(self.ctx.lib.neat_sampler_workspace_bytes(R))
            self._ws = torch.empty(n, dtype=torch.uint8, device=self.ctx.device)
            self._ws_R = R
        return self._ws

    rng = "reference"  # "reference": replay the reference's CPU-generator stream; "device": draw on the GPU (no sync)

    def get_z_vals(self, rays_o, rays_d, beta_param, training=False, randoms=None, device_tables=False):
        """rays_o [3] or [R,3], rays_d [R,3], beta_param: 0-dim/1-elem device tensor (density.beta).
        randoms (training): dict(t_rand, u_final, extra_idx, eik_idx) or None to draw them like the reference.
        device_tables: `randoms` comes from Renderer.draws() (neat_train_draws): everything is already on the device,
        including the extra-column table for EVERY candidate iteration count -- no device->host read of k.
        Returns z_vals [R, n_out], z_eik [R,1], n_iters (device int32 tensor)."""
        ctx, c, lib = self.ctx, self.cfg, self.ctx.lib
        R = rays_d.shape[0]
        dev = ctx.device
        o_stride = 3 if rays_o.dim() == 2 and rays_o.shape[0] == R and R > 1 else 0
        if o_stride == 0:
            rays_o = rays_o.reshape(-1)[:3].contiguous()
        ws = self.workspace(R)
        rn = getattr(self, "renderer", None)
        n_it = rn.pool.get("sampler.n_it", 1, torch.int32) if rn is not None else torch.zeros(1, dtype=torch.int32, device=dev)
        z_vals = torch.empty(R, self.n_out, device=dev)
        z_eik = torch.empty(R, device=dev)
        P = ctypes.c_void_p
        t_rand = u_final = eik = None
        if training:
            if device_tables:
                t_rand, u_final = randoms["t_rand"], randoms["u_final"]
            elif randoms is None:
                if self.rng == "device":
                    raise _lib.NeatError("rng='device' draws come from Renderer.draws(); pass them as `randoms`")
                t_rand = torch.rand(R, c.n_eval).to(dev)       # ray_sampler.py:87
                torch.randint(0, c.n_eval, (R,))               # :91 (drawn and unused by the reference)
                u_final = torch.rand(R, c.n_final).to(dev)     # :234
            else:
                t_rand = randoms["t_rand"].to(dev, torch.float32).contiguous()
                u_final = randoms["u_final"].to(dev, torch.float32).contiguous()
        import contextlib
        with (rn.timed("sampler") if rn is not None else contextlib.nullcontext()):
            _lib.check(lib.neat_sampler_run(
                ctx._h, ctypes.byref(c), ctx._chk(rays_o), o_stride, ctx._chk(rays_d, (R, 3)), R,
                P(beta_param.data_ptr()), P(t_rand.data_ptr()) if training else None,
                P(u_final.data_ptr()) if training else None, P(ws.data_ptr()), P(n_it.data_ptr()), ctx._stream()))
        if training and device_tables:
            table, eik = randoms["extra_table"], randoms["eik_idx"]
        elif training:
            k = int(n_it.item())
            table = torch.zeros(c.max_iters, max(c.n_extra, 1), dtype=torch.int64)
            if randoms is None:
                if c.n_extra > 0:
                    table[k - 1] = torch.randperm(c.n_eval * k)[:c.n_extra]      # :265
                eik = torch.randint(0, self.n_out, (R,))                          # :275
            else:
                if c.n_extra > 0:
                    table[k - 1] = randoms["extra_idx"].long().cpu()
                eik = randoms["eik_idx"].long()
            table = table.to(dev)
            eik = eik.to(dev).contiguous()
        else:
            table = self.eval_table
        _lib.check(lib.neat_sampler_finish(
            ctx._h, ctypes.byref(c), R, P(table.data_ptr()) if table is not None else None,
            P(eik.data_ptr()) if training else None, P(ws.data_ptr()), P(z_vals.data_ptr()),
            P(z_eik.data_ptr()), ctx._stream()))
        return z_vals, z_eik[:, None], n_it
