"""Drop-in mirror of the reference's model plugin (`train.model_class`):

    model.networks.neat_wfr_rend_a.VolSDFNetwork   ->   neat_b200.model.VolSDFNetwork

Same constructor (`conf=` sub-tree), same parameter / state_dict names (implicit_network.lin{l}.{bias,
weight_g,weight_v}, rendering_network.*, attraction_network.*, density.beta, latents, ffn.{0,2,4}.*), same
input / output dict keys (code/model/networks/neat_wfr_rend_a.py:257-538).  All heavy work runs in the
sm_100a kernels behind include/neat_b200.h; there is no PyTorch / CPU fallback."""
import math
import time as _time

import numpy as np
import torch
from torch import nn

from . import _lib, junction
from . import ffn as ffn_kernels
from .autograd import NeatStepFunction, StepState
from .context import Context
from .render import Renderer


def _plain(conf):
    """pyhocon ConfigTree / nested dict -> nested plain dict."""
    if hasattr(conf, "items"):
        return {k: _plain(v) for k, v in conf.items()}
    return conf


def _embed_dim(multires, d=3):
    return d + 2 * d * multires if multires > 0 else d


class _ProjectCalib(torch.autograd.Function):
    """lines2d_calib = project2D(I, R, T, lines3d) (neat_wfr_rend_a.py:442): the forward value comes from the geometry
    kernel of the step; the backward is project_calib_bwd_kernel."""

    @staticmethod
    def forward(ctx, lines3d, lines2d_calib, pose_inv):
        ctx.save_for_backward(lines3d.detach(), pose_inv)
        return lines2d_calib.view_as(lines2d_calib).clone()

    @staticmethod
    def backward(ctx, g):
        import ctypes
        lines3d, pose_inv = ctx.saved_tensors
        lib = _lib.load()
        R = lines3d.shape[0]
        out = torch.empty(R, 2, 3, device=lines3d.device)
        P = ctypes.c_void_p
        _lib.check(lib.neat_project_calib_backward(R, P(pose_inv.data_ptr()), P(lines3d.contiguous().data_ptr()),
                                                   P(g.contiguous().float().data_ptr()), P(out.data_ptr()),
                                                   P(torch.cuda.current_stream(lines3d.device).cuda_stream)))
        return out, None, None


class _ProjectPoints(torch.autograd.Function):
    """(project2D(K, R, T, X), project2D(I, R, T, X)) of the global junctions (neat_wfr_rend_a.py:484-486) as ONE kernel
    each way (csrc/junction.cuh) instead of two chains of ~17 eager ops.  pose_inv: world-to-camera [16] from the
    geometry kernel of the step; K4: the [4,4] intrinsics."""

    @staticmethod
    def forward(ctx, X, pose_inv, K4):
        import ctypes
        lib = _lib.load()
        Xc = X.detach().float().contiguous()
        N = Xc.shape[0]
        pix = torch.empty(N, 2, device=Xc.device)
        cal = torch.empty(N, 2, device=Xc.device)
        P = ctypes.c_void_p
        _lib.check(lib.neat_project_points(N, P(pose_inv.data_ptr()), P(K4.data_ptr()), 4, P(Xc.data_ptr()),
                                           P(pix.data_ptr()), P(cal.data_ptr()),
                                           P(torch.cuda.current_stream(Xc.device).cuda_stream)))
        ctx.save_for_backward(Xc, pose_inv, K4)
        return pix, cal

    @staticmethod
    def backward(ctx, g_pix, g_cal):
        import ctypes
        Xc, pose_inv, K4 = ctx.saved_tensors
        lib = _lib.load()
        P = ctypes.c_void_p
        N = Xc.shape[0]
        gX = torch.empty(N, 3, device=Xc.device)
        gp = None if g_pix is None else g_pix.float().contiguous()
        gc = None if g_cal is None else g_cal.float().contiguous()
        _lib.check(lib.neat_project_points_backward(N, P(pose_inv.data_ptr()), P(K4.data_ptr()), 4, P(Xc.data_ptr()),
                                                    None if gp is None else P(gp.data_ptr()),
                                                    None if gc is None else P(gc.data_ptr()), P(gX.data_ptr()),
                                                    P(torch.cuda.current_stream(Xc.device).cuda_stream)))
        return gX, None, None


class _WeightNormMLP(nn.Module):
    """lin0..lin{n-1} with nn.utils.weight_norm, parameters only: the math runs in the kernels.  Layers are created,
    initialised (`init_fn(l, lin)`) and weight-normed one at a time, in the reference's order, so that the same
    torch.manual_seed gives the same initial weights as the reference constructors (neat_wfr_rend_a.py:46-72)."""

    def __init__(self, dims_in_out, weight_norm=True, init_fn=None):
        super().__init__()
        self.num_layers = len(dims_in_out) + 1
        self._weight_norm = weight_norm
        for l, (i, o) in enumerate(dims_in_out):
            lin = nn.Linear(i, o)
            if init_fn is not None:
                init_fn(l, lin)
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin%d" % l, lin)


class ImplicitNetwork(_WeightNormMLP):
    """Parameters + standalone entry points of ImplicitNetwork (neat_wfr_rend_a.py:14-137)."""

    def __init__(self, feature_vector_size, sdf_bounding_sphere, d_in, d_out, dims, geometric_init=True, bias=1.0,
                 skip_in=(), weight_norm=True, multires=0, sphere_scale=1.0, inside_out=False):
        if d_in != 3 or d_out != 1:
            raise _lib.NeatError("ImplicitNetwork: d_in=3, d_out=1 expected")
        d0 = _embed_dim(multires)
        full = [d0] + list(dims) + [d_out + feature_vector_size]
        io = []
        for l in range(len(full) - 1):
            io.append((full[l], full[l + 1] - d0 if (l + 1) in skip_in else full[l + 1]))
        n = len(io)

        def geometric(l, lin):                                  # neat_wfr_rend_a.py:55-69, same calls in the same order
            if not geometric_init:
                return
            i, o = io[l]
            if l == n - 1:
                torch.nn.init.normal_(lin.weight, mean=math.sqrt(math.pi) / math.sqrt(i), std=0.0001)
                torch.nn.init.constant_(lin.bias, -bias)
            elif multires > 0 and l == 0:
                torch.nn.init.constant_(lin.bias, 0.0)
                torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
                torch.nn.init.normal_(lin.weight[:, :3], 0.0, math.sqrt(2) / math.sqrt(o))
            elif multires > 0 and l in skip_in:
                torch.nn.init.constant_(lin.bias, 0.0)
                torch.nn.init.normal_(lin.weight, 0.0, math.sqrt(2) / math.sqrt(o))
                torch.nn.init.constant_(lin.weight[:, -(d0 - 3):], 0.0)
            else:
                torch.nn.init.constant_(lin.bias, 0.0)
                torch.nn.init.normal_(lin.weight, 0.0, math.sqrt(2) / math.sqrt(o))

        super().__init__(io, weight_norm, init_fn=geometric)
        self.sdf_bounding_sphere, self.sphere_scale, self.skip_in = sdf_bounding_sphere, sphere_scale, tuple(skip_in)
        self.multires = multires
        self._owner = None  # set by VolSDFNetwork (gives access to the kernel context)

    # --- standalone (inference) entry points used by evaluation / mesh extraction callers ---
    def _renderer(self):
        return self._owner()._sync_weights()

    @torch.no_grad()
    def get_sdf_vals(self, x):
        rn = self._renderer()
        return rn.ctx.sdf_points(x.detach().reshape(-1, 3).float().contiguous())[:, None]

    def get_outputs(self, x):
        rn = self._renderer()
        x = x.detach().reshape(-1, 3).float().contiguous()
        sdf, grad, _, feat, _ = rn.sdf_outputs(rn.explicit_points(x), x.shape[0], clamp=True)
        return sdf[:, None], rn.unpack_features(feat, x.shape[0]), grad

    def gradient(self, x):
        rn = self._renderer()
        x = x.detach().reshape(-1, 3).float().contiguous()
        _, grad, _, _, _ = rn.sdf_outputs(rn.explicit_points(x), x.shape[0], clamp=False, want_feat=False, want_sdf=False)
        return grad

    @torch.no_grad()
    def forward(self, x):
        """[M, 1 + feature_vector_size] raw network output (no sphere clamp), as mesh extraction expects."""
        rn = self._renderer()
        x = x.detach().reshape(-1, 3).float().contiguous()
        sdf, _, _, feat, _ = rn.sdf_outputs(rn.explicit_points(x), x.shape[0], clamp=False)
        return torch.cat([sdf[:, None], rn.unpack_features(feat, x.shape[0])], dim=1)


class _Head(_WeightNormMLP):
    def __init__(self, feature_vector_size, mode, d_in, d_out, dims, weight_norm=True, multires_view=0):
        if mode != "idr":
            raise _lib.NeatError("only mode='idr' is supported")
        d0 = d_in + feature_vector_size + (_embed_dim(multires_view) - 3 if multires_view > 0 else 0)
        full = [d0] + list(dims) + [d_out]
        super().__init__([(full[l], full[l + 1]) for l in range(len(full) - 1)], weight_norm)
        self.mode, self.multires_view = mode, multires_view
        self._owner = None
        self._head = 0

    @torch.no_grad()
    def forward(self, points, normals, view_dirs, feature_vectors):
        rn = self._owner()._sync_weights()
        x = points.detach().reshape(-1, 3).float().contiguous()
        M = x.shape[0]
        feat = rn.pack_features(feature_vectors.detach().float())
        out, _ = rn.head_forward(self._head, rn.explicit_points(x, view_dirs.detach().float().contiguous()), M,
                                 normals.detach().float().contiguous(), feat)
        out = out.clone()  # the kernel output lives in the renderer's reusable workspace
        return out if self._head == 0 else out.view(M, 2, 3)


class RenderingNetwork(_Head):
    pass


class AttractionFieldNetwork(_Head):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._head = 1


class LaplaceDensity(nn.Module):
    """code/model/density.py:16-30 (parameter `beta`; the density itself is evaluated inside the kernels)."""

    def __init__(self, params_init=None, beta_min=0.0001):
        super().__init__()
        for k, v in (params_init or {}).items():
            setattr(self, k, nn.Parameter(torch.tensor(float(v))))
        self.beta_min = float(beta_min)

    def get_beta(self):
        return self.beta.abs() + self.beta_min

    def forward(self, sdf, beta=None):
        beta = self.get_beta() if beta is None else beta
        return (1.0 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


class VolSDFNetwork(nn.Module):
    def __init__(self, conf):
        super().__init__()
        c = _plain(conf)
        self.conf = c
        self.feature_vector_size = int(c["feature_vector_size"])
        self.scene_bounding_sphere = float(c.get("scene_bounding_sphere", 1.0))
        self.white_bkgd = bool(c.get("white_bkgd", False))    # handled by the compositing kernels (render.Renderer.bg_color)
        if self.white_bkgd:
            self.register_buffer("bg_color", torch.tensor([float(v) for v in c.get("bg_color", [1.0, 1.0, 1.0])]),
                                 persistent=False)
        self.implicit_network = ImplicitNetwork(self.feature_vector_size,
                                                0.0 if self.white_bkgd else self.scene_bounding_sphere, **c["implicit_network"])
        self.rendering_network = RenderingNetwork(self.feature_vector_size, **c["rendering_network"])
        self.attraction_network = AttractionFieldNetwork(self.feature_vector_size, **c["attraction_network"])
        self.density = LaplaceDensity(**c["density"])
        cj = c.get("global_junctions", {})
        nl, hid = int(cj.get("num_layers", 2)), int(cj.get("dim_hidden", 256))
        self.latents = nn.Parameter(torch.empty(int(cj.get("num_junctions", 1024)), hid))
        nn.init.normal_(self.latents, mean=0.0, std=1)
        ffn = []
        for i in range(nl + 1):
            ffn.append(nn.Linear(hid, hid if i != nl else 3))
            if i != nl:
                ffn.append(nn.ReLU())
        self.ffn = nn.Sequential(*ffn)
        self.dbscan_enabled = bool(c.get("dbscan_enabled", True))
        self.use_median = bool(c.get("use_median", False))
        self.junction_eikonal = bool(c.get("junction_eikonal", False))
        self.use_l3d = bool(c.get("use_l3d", False))
        import weakref
        ref = weakref.ref(self)
        for m in (self.implicit_network, self.rendering_network, self.attraction_network):
            m._owner = ref
        self._renderer = None
        self._packed_version = None
        self.replay = None  # tests: dict(sampler=..., eik_uniform=..., samples=(z, z_eik)) recorded draws to replay
        # "reference": the training-mode random draws replay the reference's CPU-generator calls in order (same seed =>
        # same samples; costs pageable H2D copies and one device->host read of the sampler's iteration count);
        # "device": the same distributions drawn on the GPU, fully asynchronous.
        self.rng = "reference"

    # ------------------------------------------------------------------ kernel context plumbing
    def _get_renderer(self):
        dev = self.latents.device
        if dev.type != "cuda":
            raise _lib.NeatError("neat_b200.VolSDFNetwork must be moved to a CUDA device first (.cuda()); "
                                 "there is no CPU path")
        if self._renderer is None or self._renderer.ctx.device != dev:
            self._renderer = Renderer(Context(self.conf, device=dev), self.conf)
            self._renderer.scene_bounding_sphere = self.scene_bounding_sphere
            self._packed_version = None
        return self._renderer

    def _param_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(id(p) for p in self.parameters())

    def _wn_layers(self):
        """[(weight_g | None, weight_v | weight, bias)] of every MLP layer, in flat-buffer order."""
        out = []
        for mod in (self.implicit_network, self.rendering_network, self.attraction_network):
            for l in range(mod.num_layers - 1):
                lin = getattr(mod, "lin%d" % l)
                if hasattr(lin, "weight_g"):
                    out.append((lin.weight_g, lin.weight_v, lin.bias))
                else:
                    out.append((None, lin.weight, lin.bias))
        return out

    def _sync_weights(self):
        """(inference entry points) re-pack the weight slabs if any parameter changed."""
        rn = self._get_renderer()
        v = self._param_version()
        if v != self._packed_version:
            with torch.no_grad():
                rn.effective_weights([(None if g is None else g.detach(), w.detach(), b.detach())
                                      for g, w, b in self._wn_layers()])
            self._packed_version = v
        return rn

    def seed_draws(self, seed):
        """rng='device': restart the device-side sequence of training draws (same seed => same samples)."""
        self._get_renderer().seed_draws(seed)

    # ------------------------------------------------------------------ reference helpers kept for callers
    def project2D(self, K, R, T, points3d):
        """VolSDFNetwork.project2D (neat_wfr_rend_a.py:317-331) for the evaluation callers, on the projection kernel
        (csrc/junction.cuh).  K [3,3], R [3,3], T [3,1]: world-to-camera; points3d [...,3] -> [...,2].  No gradient."""
        import ctypes
        rn = self._get_renderer()
        dev = rn.ctx.device
        shape = points3d.shape
        X = points3d.detach().to(dev, torch.float32).reshape(-1, 3).contiguous()
        rt = torch.zeros(4, 4, device=dev)
        rt[:3, :3] = R.detach().to(dev, torch.float32)
        rt[:3, 3] = T.detach().to(dev, torch.float32).reshape(3)
        K3 = K.detach().to(dev, torch.float32).contiguous()
        out = torch.empty(X.shape[0], 2, device=dev)
        P = ctypes.c_void_p
        with torch.cuda.device(dev):
            _lib.check(rn.ctx.lib.neat_project_points(X.shape[0], P(rt.data_ptr()), P(K3.data_ptr()), K3.shape[-1],
                                                      P(X.data_ptr()), P(out.data_ptr()), None, rn.ctx._stream()))
        return out.reshape(*shape[:-1], 2)

    def cluster_dbscan(self, points, eps=0.01, min_samples=2):
        """points: [N,3] device tensor (or numpy array, as the reference passes) -> cluster centroids [C,3].
        min_samples must be 2 (the only value the reference uses): DBSCAN is then the connected components of the
        eps-graph, computed on the GPU (dbscan.cuh)."""
        if min_samples != 2:
            raise _lib.NeatError("cluster_dbscan: only min_samples=2 is supported")
        rn = self._get_renderer()
        if not torch.is_tensor(points):
            points = torch.as_tensor(np.asarray(points), dtype=torch.float32)
        return rn.dbscan(points.detach().to(rn.ctx.device, torch.float32).reshape(-1, 3).contiguous(), eps)

    def volume_rendering(self, z_vals, sdf):
        """VolSDFNetwork.volume_rendering (neat_wfr_rend_a.py:540-554): weights [R,S] from depths and sdf, on the
        compositing kernel (weights-only call).  No gradient."""
        rn = self._get_renderer()
        dev = rn.ctx.device
        z = z_vals.detach().to(dev, torch.float32).contiguous()
        R, S = z.shape
        s = sdf.detach().to(dev, torch.float32).reshape(R, S).contiguous()
        beta = self.density.beta.detach().reshape(1).float().contiguous()
        return rn.composite_weights(z, s, beta)

    # ------------------------------------------------------------------ forward
    def forward(self, input):
        rn = self._get_renderer()
        dev = rn.ctx.device
        K4 = input["intrinsics"][0].to(dev, torch.float32)
        if tuple(K4.shape) == (3, 3):     # BlenderDataset hands out the 3x3 calibration (blender_hawp_dataset.py:40-41)
            K4 = torch.block_diag(K4, torch.ones(1, 1, device=dev))
        K4 = K4.contiguous()
        pose = input["pose"][0].to(dev, torch.float32).contiguous()
        uv = input["uv"].reshape(-1, 2).to(dev, torch.float32).contiguous()
        uv_proj = input["uv_proj"].reshape(-1, 2).to(dev, torch.float32).contiguous()
        R = uv.shape[0]
        out = {}
        if not self.training:
            self._sync_weights()
            beta = self.density.beta.detach().reshape(1).float().contiguous()
            o = rn.forward_eval(uv, pose, K4, uv_proj, beta)
            for k in ("points", "rgb_values", "depth", "xyz", "points3d", "lines3d", "lines2d", "lines2d_calib", "l3d",
                      "sdf", "normal_map"):
                out[k] = o[k]
            out["wireframe-gt"] = input.get("wireframe")
            out["K"] = K4[:3, :3]
            return out

        st = StepState()
        st.uv, st.pose, st.K, st.uv_proj = uv, pose, K4, uv_proj
        rn.sampler.rng = self.rng
        st.sampler_randoms = self.replay.get("sampler") if self.replay else None
        st.eik_uniform = self.replay.get("eik_uniform") if self.replay else None
        st.samples_override = self.replay.get("samples") if self.replay else None
        # global junctions (independent of the render step): enqueue first so that their host copy rides in the step's
        # single device->host transfer
        glob = ffn_kernels.apply_module(self.ffn, self.latents)   # own fp32 GEMM kernels (ffn.py), no cuBLAS
        st.junction_inputs = (glob.detach(), pose, K4)
        st.dbscan_enabled = self.dbscan_enabled
        st.junction_eikonal = self.junction_eikonal
        st.use_l3d = self.use_l3d
        st.param_layers = self._wn_layers()
        self._packed_version = None  # the step packs its own copy
        # RNG order of the reference: the sampler's draws, then the eikonal uniform_ (neat_wfr_rend_a.py:518); the junction
        # block in between draws nothing, so the step makes the eikonal draw right after the sampler and the (host-side)
        # junction block follows on the detached outputs.
        anchor = self.density.beta
        if not anchor.requires_grad:  # the step hangs off one differentiable input; see NeatStepFunction.forward
            anchor = anchor.detach().requires_grad_(True)
        rgb_values, lines3d, grad_theta = NeatStepFunction.apply(anchor, rn, st)
        self.last_step = st
        K3 = K4[:3, :3]
        out.update(points=st.cam[None, None, :] + st.z[:, :, None] * st.dirs[:, None, :], rgb_values=rgb_values,
                   depth=st.depth, xyz=st.points3d, points3d=st.points3d, lines3d=lines3d, l3d=st.l3d,
                   lines2d=st.lines2d, lines2d_calib=_ProjectCalib.apply(lines3d, st.lines2d_calib, st.pose_inv.reshape(-1)),
                   sdf=st.sdf3,
                   K=K3, grad_theta=grad_theta)
        out["wireframe-gt"] = input.get("wireframe")
        # ---- junction block (neat_wfr_rend_a.py:457-496).  DBSCAN runs on the GPU; the two Hungarian assignments (here and
        # in the loss, loss_wfr.py:104-108) stay on the host as in the reference, but behind ONE device->host transfer:
        # everything they need (cluster centroids, their count, the global junctions) is fetched together, and the
        # loss' assignment is handed over in the output dict so that it need not synchronise again.
        j2d_global, j2d_global_calib = _ProjectPoints.apply(glob, st.pose_inv.reshape(-1).contiguous(), K4)
        _t0 = _time.perf_counter()
        st.junction_event.synchronize()
        _t1 = _time.perf_counter()
        n_h, cent_h, glob_h, pose_h, K_h = st.junction_host
        C = int(n_h[0])
        gt = input["wireframe"][0].vertices.detach().cpu().numpy()
        # projections, both assignments and the < 10 px filter in native host code (csrc/junction.cpp)
        local, b0, b1, n_close, med = junction.junction_match(cent_h[:C], gt, pose_h, K_h, glob_h, self.use_median)
        if self.use_median:
            out["median"] = torch.tensor(float(med), device=dev)
        n = local.shape[0]
        # one pinned staging buffer, one host->device copy: [n,7] floats, then the n + n assignment indices
        stage = rn.pinned("junction.stage", 9 * max(n, 1), torch.float32)
        stage[:7 * n].copy_(torch.from_numpy(local.reshape(-1)))
        idx = stage[7 * n:9 * n].view(torch.int32)
        idx[:n].copy_(torch.from_numpy(b0.astype(np.int32)))
        idx[n:].copy_(torch.from_numpy(b1.astype(np.int32)))
        on_dev = stage[:9 * n].to(dev, non_blocking=True)
        packed = on_dev[:7 * n].view(n, 7)
        if n:
            di = on_dev[7 * n:].view(torch.int32)
            out["_junction_assignment"] = (di[:n], di[n:], int(n_close))
        self.last_host_ms = {"wait_for_gpu": (_t1 - _t0) * 1e3, "junction_host": (_time.perf_counter() - _t1) * 1e3,
                             "clusters": C, "gt_junctions": int(gt.shape[0]), "matched": int(n)}
        out.update(j2d_local=packed[:, 3:5], j3d_local=packed[:, 0:3], j3d_global=glob, j2d_global=j2d_global,
                   j2d_local_calib=packed[:, 5:7], j2d_global_calib=j2d_global_calib)
        return out
