"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the NEAT attraction-field training step.

A plain-torch (CPU, fp32 or fp64) restatement of the reference algorithm on the hot path
(SURVEY.md section 8a).  It is NOT the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import it.  The product path
(``neat_b200``) never imports anything from ``oracle/`` and has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
golden vectors produced by the unmodified reference (``oracle/make_golden.py`` run in the
build container through ``oracle/ref_shim.py``; fixtures in ``tests/golden/``), and -- when
``/root/reference`` is present -- against the live reference classes.

Each function cites the reference lines it follows (paths relative to ``/root/reference``).
The functions work on explicit *effective* weights (weight_norm already applied), because
that is what the CUDA kernels consume.

Besides the forward path the oracle spells out the hand-derived backward recurrences the
CUDA kernels implement; ``tests/test_oracle_backward.py`` proves
them equal to ``torch.autograd`` on the forward restatement.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

SOFTPLUS_BETA = 100.0      # nn.Softplus(beta=100): code/model/networks/neat_wfr_rend_a.py:76
SOFTPLUS_THRESHOLD = 20.0  # torch default threshold


# ----------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------
@dataclass
class NeatParams:
    """Effective (weight-normed) weights of the three MLPs + scalars of the model."""
    sdf_W: List[torch.Tensor]
    sdf_b: List[torch.Tensor]
    rend_W: List[torch.Tensor]
    rend_b: List[torch.Tensor]
    att_W: List[torch.Tensor]
    att_b: List[torch.Tensor]
    beta_param: torch.Tensor                 # density.beta (raw parameter)
    beta_min: float = 1e-4
    skip_in: Sequence[int] = (4,)
    multires: int = 6
    multires_view: int = 4                   # rendering head only; attraction head: 0
    sphere_radius: float = 3.0               # scene_bounding_sphere
    sphere_scale: float = 20.0
    ffn_W: List[torch.Tensor] = field(default_factory=list)
    ffn_b: List[torch.Tensor] = field(default_factory=list)
    latents: Optional[torch.Tensor] = None
    dbscan_enabled: bool = True              # neat_wfr_rend_a.py:312-315 (model conf)
    use_median: bool = False
    bg_color: Optional[torch.Tensor] = None  # white_bkgd (neat_wfr_rend_a.py:262-263, 411-413); also means sphere_radius = 0 (:266)
    junction_eikonal: bool = False           # neat_wfr_rend_a.py:524-525
    use_l3d: bool = False                    # neat_wfr_rend_a.py:461-465 (only read when dbscan_enabled is False)

    def sdf_sphere(self):
        # ImplicitNetwork's sdf_bounding_sphere: 0 (no clamp) under white_bkgd (neat_wfr_rend_a.py:266)
        return 0.0 if self.bg_color is not None else self.sphere_radius

    def beta(self):
        # code/model/density.py:28-30
        return self.beta_param.abs() + self.beta_min

    def to(self, dtype):
        cv = lambda ts: [t.to(dtype) for t in ts]
        return NeatParams(cv(self.sdf_W), cv(self.sdf_b), cv(self.rend_W), cv(self.rend_b),
                          cv(self.att_W), cv(self.att_b), self.beta_param.to(dtype),
                          self.beta_min, tuple(self.skip_in), self.multires, self.multires_view,
                          self.sphere_radius, self.sphere_scale, cv(self.ffn_W), cv(self.ffn_b),
                          None if self.latents is None else self.latents.to(dtype), self.dbscan_enabled, self.use_median,
                          None if self.bg_color is None else self.bg_color.to(dtype), self.junction_eikonal, self.use_l3d)


def weight_norm_effective(g, v):
    """nn.utils.weight_norm (dim=0): w = g * v / ||v||_row
    (code/model/networks/neat_wfr_rend_a.py:71-72)."""
    return g * v / v.norm(2, dim=1, keepdim=True)


def params_from_state_dict(sd, skip_in=(4,), multires=6, multires_view=4, sphere_radius=3.0,
                           sphere_scale=20.0, beta_min=1e-4, track=False):
    """Build NeatParams from a VolSDFNetwork.state_dict() (keys: SURVEY.md section 5).
    track=True keeps the autograd graph to the tensors in ``sd`` (for gradient parity)."""
    if not track:
        sd = {k: v.detach().clone() for k, v in sd.items()}
    def mlp(prefix):
        Ws, bs, l = [], [], 0
        while f"{prefix}.lin{l}.bias" in sd:
            if f"{prefix}.lin{l}.weight_g" in sd:
                W = weight_norm_effective(sd[f"{prefix}.lin{l}.weight_g"], sd[f"{prefix}.lin{l}.weight_v"])
            else:
                W = sd[f"{prefix}.lin{l}.weight"]
            Ws.append(W)
            bs.append(sd[f"{prefix}.lin{l}.bias"])
            l += 1
        return Ws, bs
    sW, sb = mlp("implicit_network")
    rW, rb = mlp("rendering_network")
    aW, ab = mlp("attraction_network")
    fW, fb = [], []
    i = 0
    while f"ffn.{i}.weight" in sd:
        fW.append(sd[f"ffn.{i}.weight"])
        fb.append(sd[f"ffn.{i}.bias"])
        i += 2
    return NeatParams(sW, sb, rW, rb, aW, ab, sd["density.beta"], beta_min,
                      tuple(skip_in), multires, multires_view, sphere_radius, sphere_scale,
                      fW, fb, sd["latents"] if "latents" in sd else None)


# ----------------------------------------------------------------------------------------
# elementary pieces
# ----------------------------------------------------------------------------------------
def embed(x, n_freq):
    """NeRF positional encoding [x, sin(2^j x), cos(2^j x)]_j  (code/model/embedder.py:5-36)."""
    if n_freq <= 0:
        return x
    out = [x]
    for j in range(n_freq):
        f = float(2.0 ** j)
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, dim=-1)


def embed_jacobian_diag(x, n_freq):
    """d embed_c / d x_(c mod 3): the Jacobian of ``embed`` is block diagonal per coordinate."""
    out = [torch.ones_like(x)]
    for j in range(n_freq):
        f = float(2.0 ** j)
        out.append(f * torch.cos(x * f))
        out.append(-f * torch.sin(x * f))
    return torch.cat(out, dim=-1)


def softplus100(z):
    """nn.Softplus(beta=100, threshold=20)."""
    bz = z * SOFTPLUS_BETA
    # the unselected branch is evaluated on a clamped argument so that autograd never sees inf/inf
    soft = torch.log1p(torch.exp(bz.clamp(max=SOFTPLUS_THRESHOLD))) / SOFTPLUS_BETA
    return torch.where(bz > SOFTPLUS_THRESHOLD, z, soft)


def softplus100_d1(z):
    """sigma'(z) as autograd computes it for Softplus: sigmoid(beta z), 1 above the threshold."""
    bz = z * SOFTPLUS_BETA
    e = torch.exp(bz.clamp(max=SOFTPLUS_THRESHOLD))
    return torch.where(bz > SOFTPLUS_THRESHOLD, torch.ones_like(z), e / (e + 1.0))


def softplus100_d2(z):
    """sigma''(z) = beta * s * (1 - s); 0 above the threshold (autograd's double backward)."""
    bz = z * SOFTPLUS_BETA
    s = torch.sigmoid(bz)
    return torch.where(bz > SOFTPLUS_THRESHOLD, torch.zeros_like(z), SOFTPLUS_BETA * s * (1.0 - s))


def laplace_density(sdf, beta):
    """code/model/density.py:21-26."""
    return (1.0 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


# ----------------------------------------------------------------------------------------
# SDF network (ImplicitNetwork): forward, clamp, analytic normal
# ----------------------------------------------------------------------------------------
def sdf_forward(P: NeatParams, x, save=False):
    """ImplicitNetwork.forward (code/model/networks/neat_wfr_rend_a.py:78-96) -> [M, 1+F].
    With save=True also returns the pre-activations z_l and layer inputs u_l."""
    h0 = embed(x, P.multires)
    h = h0
    zs, us = [], []
    L = len(P.sdf_W)
    for l in range(L):
        if l in P.skip_in:
            h = torch.cat([h, h0], dim=1) / math.sqrt(2.0)
        us.append(h)
        z = h @ P.sdf_W[l].t() + P.sdf_b[l]
        zs.append(z)
        h = softplus100(z) if l < L - 1 else z
    return (h, zs, us) if save else h


def sphere_sdf(P: NeatParams, x):
    return P.sphere_scale * (P.sdf_sphere() - x.norm(2, 1, keepdim=True))


def sdf_vals(P: NeatParams, x):
    """ImplicitNetwork.get_sdf_vals (neat_wfr_rend_a.py:131-137)."""
    s = sdf_forward(P, x)[:, :1]
    if P.sdf_sphere() > 0.0:
        s = torch.minimum(s, sphere_sdf(P, x))
    return s


def sdf_outputs(P: NeatParams, x, clamp=True):
    """ImplicitNetwork.get_outputs (neat_wfr_rend_a.py:111-129), normal by the hand-derived
    reverse pass of SURVEY.md Appendix A step 2 (no autograd).  clamp=False gives
    ImplicitNetwork.gradient semantics (neat_wfr_rend_a.py:98-109: no sphere clamp).
    Returns sdf[M,1], feat[M,F], grad[M,3] and a dict of saved tensors."""
    out, zs, us = sdf_forward(P, x, save=True)
    L = len(P.sdf_W)
    s_raw = out[:, :1]
    feat = out[:, 1:]
    clamp = clamp and P.sdf_sphere() > 0.0
    if clamp:
        sph = sphere_sdf(P, x)
        act = (s_raw <= sph).to(x.dtype)      # torch.minimum sends the gradient to `self` on ties
        sdf = torch.minimum(s_raw, sph)
    else:
        act = torch.ones_like(s_raw)
        sdf = s_raw
    d_embed = 3 + 6 * P.multires if P.multires > 0 else 3
    # reverse pass: g_l = d s_raw / d h_l (gradient w.r.t. the OUTPUT of layer l-1)
    a = torch.zeros_like(out)
    a[:, 0] = 1.0
    gs = [None] * (L + 1)
    r_skip = torch.zeros(x.shape[0], d_embed, dtype=x.dtype)
    for l in range(L - 1, -1, -1):
        v = a @ P.sdf_W[l]
        if l in P.skip_in:
            v = v / math.sqrt(2.0)
            r_skip = r_skip + v[:, -d_embed:]
            v = v[:, :-d_embed]
        gs[l] = v
        if l > 0:
            a = softplus100_d1(zs[l - 1]) * v
    J = embed_jacobian_diag(x, P.multires)
    t = (gs[0] + r_skip) * J
    n_net = t.reshape(x.shape[0], -1, 3).sum(1)
    n_sph = -P.sphere_scale * x / x.norm(2, 1, keepdim=True)
    grad = act * n_net + (1.0 - act) * n_sph if clamp else n_net
    saved = dict(zs=zs, us=us, gs=gs, act=act, s_raw=s_raw, n_net=n_net)
    return sdf, feat, grad, saved


def sdf_outputs_autograd(P: NeatParams, x, create_graph=False):
    """Same as ``sdf_outputs`` but with torch.autograd, exactly as the reference does."""
    x = x.detach().clone().requires_grad_(True)
    out = sdf_forward(P, x)
    sdf = out[:, :1]
    if P.sdf_sphere() > 0.0:
        sdf = torch.minimum(sdf, sphere_sdf(P, x))
    g = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=create_graph,
                            retain_graph=True)[0]
    return sdf, out[:, 1:], g, x


def sdf_gradient_autograd(P: NeatParams, x, create_graph=False):
    """ImplicitNetwork.gradient (neat_wfr_rend_a.py:98-109): no sphere clamp."""
    x = x.detach().clone().requires_grad_(True)
    y = sdf_forward(P, x)[:, :1]
    return torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph,
                               retain_graph=True)[0]


# ----------------------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------------------
def head_forward(Ws, bs, inp, save=False):
    h = inp
    zs, us = [], []
    for l, (W, b) in enumerate(zip(Ws, bs)):
        us.append(h)
        z = h @ W.t() + b
        zs.append(z)
        h = torch.relu(z) if l < len(Ws) - 1 else z
    return (h, zs, us) if save else h


def rendering_input(P, pts, normals, dirs, feat):
    # mode 'idr': cat[points, PE(view), normals, feat]   (neat_wfr_rend_a.py:235-240)
    return torch.cat([pts, embed(dirs, P.multires_view), normals, feat], dim=-1)


def attraction_input(P, pts, normals, dirs, feat):
    # attraction head does not embed the view dirs (dtu.conf:54-61; neat_wfr_rend_a.py:175-180)
    return torch.cat([pts, dirs, normals, feat], dim=-1)


def rendering_forward(P, pts, normals, dirs, feat):
    """RenderingNetwork.forward (neat_wfr_rend_a.py:235-255) -> rgb[M,3] in (0,1)."""
    return torch.sigmoid(head_forward(P.rend_W, P.rend_b, rendering_input(P, pts, normals, dirs, feat)))


def attraction_forward(P, pts, normals, dirs, feat):
    """AttractionFieldNetwork.forward (neat_wfr_rend_a.py:175-197) -> lines3d[M,2,3]."""
    off = head_forward(P.att_W, P.att_b, attraction_input(P, pts, normals, dirs, feat))
    return pts[:, None] + off.reshape(-1, 2, 3)


# ----------------------------------------------------------------------------------------
# camera / geometry
# ----------------------------------------------------------------------------------------
def camera_rays(uv, pose, K):
    """rend_util.get_camera_params + lift (code/utils/rend_util.py:55-81, 95-108).
    uv [R,2], pose [4,4] camera-to-world, K [4,4] or [3,3] -> dirs [R,3] (unit), cam_loc [3]."""
    fx, fy, cx, cy, sk = K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1]
    x, y = uv[:, 0], uv[:, 1]
    z = torch.ones_like(x)
    xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    yl = (y - cy) / fy * z
    pc = torch.stack([xl, yl, z, torch.ones_like(z)], dim=-1)          # [R,4]
    world = (pose @ pc.t()).t()[:, :3]
    cam = pose[:3, 3]
    d = world - cam[None]
    d = d / d.norm(2, dim=1, keepdim=True).clamp_min(1e-12)           # F.normalize
    return d, cam


def project2d(K3, Rm, T, X):
    """VolSDFNetwork.project2D (neat_wfr_rend_a.py:317-331).  X [...,3] -> [...,2]."""
    shp = X.shape
    Xf = X.reshape(-1, 3)
    x = (K3 @ (Rm @ Xf.t() + T)).t()
    den = x[:, -1:]
    sign = torch.where(den >= 0, torch.ones_like(den), -torch.ones_like(den))
    eps = torch.where(den.abs() < 1e-8, torch.full_like(den, 1e-8), torch.zeros_like(den))
    x = x / (den + eps * sign)
    return x.reshape(*shp)[..., :2]


def pose_inverse_rt(pose):
    """Closed form of pose.inverse()[:3] for a rigid camera-to-world matrix would be
    [R^T | -R^T t]; the reference uses a general inverse (neat_wfr_rend_a.py:433), and so do we."""
    inv = torch.linalg.inv(pose)
    return inv[:3, :3], inv[:3, 3:]


# ----------------------------------------------------------------------------------------
# volume rendering
# ----------------------------------------------------------------------------------------
def volume_weights(z_vals, sdf, beta):
    """VolSDFNetwork.volume_rendering (neat_wfr_rend_a.py:540-554). sdf [R,S] -> w [R,S]."""
    sigma = laplace_density(sdf, beta)
    d = z_vals[:, 1:] - z_vals[:, :-1]
    d = torch.cat([d, torch.full_like(z_vals[:, :1], 1e10)], dim=-1)
    fe = d * sigma
    sh = torch.cat([torch.zeros_like(fe[:, :1]), fe[:, :-1]], dim=-1)
    alpha = 1.0 - torch.exp(-fe)
    T = torch.exp(-torch.cumsum(sh, dim=-1))
    return alpha * T


def volume_weights_backward(z_vals, sdf, beta, w_bar):
    """Hand-derived adjoint of ``volume_weights`` w.r.t. sdf and beta (what the CUDA compositing
    backward implements).  Returns (sdf_bar [R,S], beta_bar scalar)."""
    sigma = laplace_density(sdf, beta)
    d = z_vals[:, 1:] - z_vals[:, :-1]
    d = torch.cat([d, torch.full_like(z_vals[:, :1], 1e10)], dim=-1)
    fe = d * sigma
    sh = torch.cat([torch.zeros_like(fe[:, :1]), fe[:, :-1]], dim=-1)
    T = torch.exp(-torch.cumsum(sh, dim=-1))
    em = torch.exp(-fe)
    w = (1.0 - em) * T
    ww = w_bar * w
    # suffix sum over i>k of w_bar_i w_i
    suffix = torch.flip(torch.cumsum(torch.flip(ww, [1]), 1), [1]) - ww
    fe_bar = w_bar * em * T - suffix
    sigma_bar = fe_bar * d
    e = torch.exp(-sdf.abs() / beta)
    sgn2 = sdf.sign() ** 2
    dsig_ds = -e / (2.0 * beta * beta) * sgn2
    dsig_db = -sigma / beta + sdf * e / (2.0 * beta ** 3)
    return sigma_bar * dsig_ds, (sigma_bar * dsig_db).sum()


# ----------------------------------------------------------------------------------------
# error-bound sampler (VolSDF Algorithm 1)
# ----------------------------------------------------------------------------------------
@dataclass
class SamplerConf:
    near: float = 0.0
    N_samples: int = 64
    N_samples_eval: int = 128
    N_samples_extra: int = 32
    eps: float = 0.1
    beta_iters: int = 10
    max_total_iters: int = 5
    add_tiny: float = 0.0


@dataclass
class SamplerRandoms:
    """The CPU-generator draws of a training-mode call, in the reference's order
    (SURVEY.md section 3.3): rand[R,Ne], rand[R,Ns], randperm(L)[:extra], randint[R]."""
    t_rand: torch.Tensor          # ray_sampler.py:87
    u_final: torch.Tensor         # ray_sampler.py:234
    extra_idx: torch.Tensor       # ray_sampler.py:265 (already truncated to N_samples_extra)
    eik_idx: torch.Tensor         # ray_sampler.py:275


def uniform_z(R, conf: SamplerConf, far, t_rand=None, dtype=torch.float32):
    """UniformSampler.get_z_vals (code/model/ray_sampler.py:69-95)."""
    t = torch.linspace(0.0, 1.0, steps=conf.N_samples_eval, dtype=dtype)
    near = torch.full((R, 1), conf.near, dtype=dtype)
    farv = torch.full((R, 1), far, dtype=dtype)
    z = near * (1.0 - t) + farv * t
    if t_rand is not None:
        mids = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mids, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


def error_bound(sdf, dists, d_star, beta, beta_q):
    """ErrorBoundSampler.get_error_bound (ray_sampler.py:285-293); beta_q [R,1] or scalar."""
    sigma = laplace_density(sdf, beta_q)
    sh = torch.cat([torch.zeros_like(dists[:, :1]), dists * sigma[:, :-1]], dim=-1)
    integral = torch.cumsum(sh, dim=-1)
    eps_sec = torch.exp(-d_star / beta_q) * (dists ** 2.0) / (4.0 * beta_q ** 2)
    eint = torch.cumsum(eps_sec, dim=-1)
    bound = (torch.clamp(torch.exp(eint), max=1.0e6) - 1.0) * torch.exp(-integral[:, :-1])
    return bound.max(-1)[0]


def interval_dstar(z, d):
    """Theorem-1 bound d* per interval (ray_sampler.py:161-173)."""
    a = z[:, 1:] - z[:, :-1]
    b = d[:, :-1].abs()
    c = d[:, 1:].abs()
    first = a.pow(2) + b.pow(2) <= c.pow(2)
    second = a.pow(2) + c.pow(2) <= b.pow(2)
    s = (a + b + c) / 2.0
    area = s * (s - a) * (s - b) * (s - c)
    tri = ~first & ~second & (b + c - a > 0)
    ds = torch.zeros_like(a)
    ds = torch.where(first, b, ds)
    ds = torch.where(second, c, ds)
    ds = torch.where(tri, 2.0 * torch.sqrt(torch.where(tri, area, torch.zeros_like(area))) / a, ds)
    same = (d[:, 1:].sign() * d[:, :-1].sign() == 1).to(z.dtype)
    return same * ds, a


def inverse_cdf(bins, cdf, u):
    """ray_sampler.py:237-249 (searchsorted right=True, clamp, lerp, denom<1e-5 -> 1)."""
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp_min(0)
    above = inds.clamp_max(cdf.shape[-1] - 1)
    c0 = torch.gather(cdf, 1, below)
    c1 = torch.gather(cdf, 1, above)
    b0 = torch.gather(bins, 1, below)
    b1 = torch.gather(bins, 1, above)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0)


def error_bound_sampler(P: NeatParams, conf: SamplerConf, dirs, cam, training=False,
                        rnd: Optional[SamplerRandoms] = None, sdf_fn=None, trace=None):
    """ErrorBoundSampler.get_z_vals (code/model/ray_sampler.py:130-283).
    dirs [R,3], cam [R,3].  Returns z_vals [R, N_samples+2+extra], z_eik [R,1], n_iters.
    ``sdf_fn(points[M,3]) -> [M,1]`` defaults to the oracle SDF; ``trace`` (a list) collects
    the per-iteration state for stage-wise kernel tests."""
    if sdf_fn is None:
        sdf_fn = lambda p: sdf_vals(P, p)
    dt = dirs.dtype
    R = dirs.shape[0]
    far = 2.0 * P.sphere_radius
    beta0 = P.beta().detach().to(dt)
    z = uniform_z(R, conf, far, rnd.t_rand if training else None, dtype=dt)
    samples, order = z, None
    dists0 = z[:, 1:] - z[:, :-1]
    beta = torch.sqrt((1.0 / (4.0 * math.log(conf.eps + 1.0))) * (dists0 ** 2.0).sum(-1))
    it, not_conv = 0, True
    sdf = None
    while not_conv and it < conf.max_total_iters:
        pts = cam[:, None, :] + samples[:, :, None] * dirs[:, None, :]
        with torch.no_grad():
            new_sdf = sdf_fn(pts.reshape(-1, 3)).reshape(R, -1)
        if order is not None:
            sdf = torch.gather(torch.cat([sdf, new_sdf], -1), 1, order)
        else:
            sdf = new_sdf
        d_star, dists = interval_dstar(z, sdf)
        err = error_bound(sdf, dists, d_star, beta0, beta0)
        beta = torch.where(err <= conf.eps, beta0.expand_as(beta), beta)
        lo, hi = beta0.expand(R).clone(), beta.clone()
        for _ in range(conf.beta_iters):
            mid = (lo + hi) / 2.0
            err = error_bound(sdf, dists, d_star, beta0, mid[:, None])
            ok = err <= conf.eps
            hi = torch.where(ok, mid, hi)
            lo = torch.where(ok, lo, mid)
        beta = hi
        sigma = laplace_density(sdf, beta[:, None])
        dd = torch.cat([dists, torch.full_like(dists[:, :1], 1e10)], -1)
        fe = dd * sigma
        sh = torch.cat([torch.zeros_like(fe[:, :1]), fe[:, :-1]], -1)
        alpha = 1.0 - torch.exp(-fe)
        T = torch.exp(-torch.cumsum(sh, -1))
        w = alpha * T
        it += 1
        not_conv = bool(beta.max() > beta0)
        more = not_conv and it < conf.max_total_iters
        if more:
            N = conf.N_samples_eval
            eps_sec = torch.exp(-d_star / beta[:, None]) * (dists ** 2.0) / (4.0 * beta[:, None] ** 2)
            eint = torch.cumsum(eps_sec, -1)
            pdf = (torch.clamp(torch.exp(eint), max=1.0e6) - 1.0) * T[:, :-1] + conf.add_tiny
        else:
            N = conf.N_samples
            pdf = w[:, :-1] + 1e-5
        pdf = pdf / pdf.sum(-1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
        if more or not training:
            u = torch.linspace(0.0, 1.0, steps=N, dtype=dt)[None].repeat(R, 1)
        else:
            u = rnd.u_final.to(dt)
        samples = inverse_cdf(z, cdf, u.contiguous())
        if trace is not None:
            trace.append(dict(z=z.clone(), sdf=sdf.clone(), beta=beta.clone(), d_star=d_star.clone(),
                              weights=w.clone(), cdf=cdf.clone(), samples=samples.clone(), more=more))
        if more:
            z, order = torch.sort(torch.cat([z, samples], -1), -1)
    near = torch.full((R, 1), conf.near, dtype=dt)
    farv = torch.full((R, 1), far, dtype=dt)
    if conf.N_samples_extra > 0:
        if training:
            idx = rnd.extra_idx
        else:
            idx = torch.linspace(0, z.shape[1] - 1, conf.N_samples_extra).long()
        extra = torch.cat([near, farv, z[:, idx]], -1)
    else:
        extra = torch.cat([near, farv], -1)
    z_out, _ = torch.sort(torch.cat([samples, extra], -1), -1)
    if training:
        eidx = rnd.eik_idx
    else:
        eidx = torch.zeros(R, dtype=torch.long)
    z_eik = torch.gather(z_out, 1, eidx[:, None])
    return z_out, z_eik, it


# ----------------------------------------------------------------------------------------
# render-point pass + compositing  (VolSDFNetwork.forward, neat_wfr_rend_a.py:392-429)
# ----------------------------------------------------------------------------------------
def render_rays(P: NeatParams, dirs, cam, z_vals):
    """Returns dict with per-point and per-ray quantities of the forward (eval semantics)."""
    R, S = z_vals.shape
    rays_d = z_vals[:, :, None] * dirs[:, None, :]
    depth_ratio = rays_d.norm(dim=-1)
    pts = cam[:, None, :] + rays_d
    pf = pts.reshape(-1, 3)
    df = dirs[:, None, :].expand(R, S, 3).reshape(-1, 3)
    sdf, feat, grad, saved = sdf_outputs(P, pf)
    rgb = rendering_forward(P, pf, grad, df, feat).reshape(R, S, 3)
    w = volume_weights(z_vals, sdf.reshape(R, S), P.beta())
    rgb_values = (w[..., None] * rgb).sum(1)
    if P.bg_color is not None:                                # neat_wfr_rend_a.py:411-413
        rgb_values = rgb_values + (1.0 - w.sum(-1))[..., None] * P.bg_color.to(w.dtype)[None]
    l3 = attraction_forward(P, pf, grad, df, feat).reshape(R, S, 2, 3)
    lines3d = (w[:, :, None, None].detach() * l3).sum(1)      # weights detached: :410
    depth = (w * depth_ratio).sum(-1)
    points3d = (w[..., None] * pts).sum(1)
    nrm = grad / grad.norm(2, -1, keepdim=True)
    normal_map = (w[..., None] * nrm.reshape(R, S, 3)).sum(1)
    return dict(points=pts, sdf_pts=sdf.reshape(R, S), grad=grad.reshape(R, S, 3), feat=feat,
                rgb_pts=rgb, weights=w, rgb_values=rgb_values, lines3d_pts=l3, lines3d=lines3d,
                depth=depth, xyz=points3d, points3d=points3d, normal_map=normal_map)


def line_geometry(P: NeatParams, K, pose, uv_proj, points3d, lines3d):
    """neat_wfr_rend_a.py:429-456: second get_outputs at points3d, 2D projections, l3d."""
    sdf3, _, g3, _ = sdf_outputs(P, points3d)
    Rm, T = pose_inverse_rt(pose)
    K3 = K[:3, :3]
    lines2d = project2d(K3, Rm, T, lines3d.detach())           # :439
    lines2d_calib = project2d(torch.eye(3, dtype=K.dtype), Rm, T, lines3d)
    rd, ro = camera_rays(uv_proj, pose, K)
    den = (rd * g3).sum(-1)
    den_eps = torch.where(den >= 0, torch.full_like(den, 1e-6), torch.full_like(den, -1e-6))
    t = (((points3d - ro[None]) * g3).sum(-1) / (den + den_eps)).detach()   # :452-453
    l3d = ro[None] + rd * t[:, None]
    return dict(sdf=sdf3.flatten(), lines2d=lines2d, lines2d_calib=lines2d_calib, l3d=l3d)


# ----------------------------------------------------------------------------------------
# junctions (neat_wfr_rend_a.py:333-342, 457-496)
# ----------------------------------------------------------------------------------------
def dbscan_centroids(points, eps=0.01, min_samples=2):
    """With min_samples=2 every non-noise point is a core point, so DBSCAN == connected
    components of the eps-graph minus singletons (SURVEY.md section 7 hard part 5).
    Components are ordered by their smallest point index (sklearn labels clusters in scan
    order, which is the same).  points: numpy [N,3] float -> numpy [C,3] float64 means."""
    import numpy as np
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    assert min_samples == 2
    pts = np.asarray(points, dtype=np.float64)
    n = len(pts)
    pairs = cKDTree(pts).query_pairs(eps, output_type="ndarray")
    if len(pairs) == 0:
        return np.zeros((0, 3))
    # query_pairs uses <= eps like sklearn's radius_neighbors
    g = coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    counts = np.bincount(lab)
    first = {}
    for i, l in enumerate(lab):
        if counts[l] >= 2 and l not in first:
            first[l] = i
    order = sorted(first, key=lambda l: first[l])
    return np.stack([pts[lab == l].mean(axis=0) for l in order]) if order else np.zeros((0, 3))


def junction_ffn(P: NeatParams):
    h = P.latents
    for i, (W, b) in enumerate(zip(P.ffn_W, P.ffn_b)):
        h = h @ W.t() + b
        if i < len(P.ffn_W) - 1:
            h = torch.relu(h)
    return h


# ----------------------------------------------------------------------------------------
# loss (code/model/networks/loss_wfr.py)
# ----------------------------------------------------------------------------------------
def line_loss(lines2d, gt, weight, threshold=100):
    """VolSDFLoss.get_line_loss (loss_wfr.py:34-45)."""
    sw = gt[:, [2, 3, 0, 1]]
    d1 = ((lines2d - gt) ** 2).sum(-1, keepdim=True).detach()
    d2 = ((lines2d - sw) ** 2).sum(-1, keepdim=True).detach()
    tgt = torch.where(d1 < d2, gt, sw)
    per = (lines2d - tgt).abs().mean(-1)
    lab = (per.detach() < threshold).long()
    return (per * weight.flatten() * lab).sum() / lab.sum().clamp_min(1), per.detach()


def neat_loss(out, rgb_gt, lines2d_gt5, K3, eikonal_weight=0.1, line_weight=0.01,
              junction_3d_weight=0.1, junction_2d_weight=0.01):
    """VolSDFLoss.forward (loss_wfr.py:47-139), rgb_loss = L1 mean."""
    from scipy.optimize import linear_sum_assignment
    gt, wgt = lines2d_gt5[:, :4], lines2d_gt5[:, 4:]
    l2d_uncal, thr = line_loss(out["lines2d"].reshape(-1, 4), gt, wgt)
    count = (thr < 100).sum()
    g2 = gt.reshape(-1, 2)
    gh = torch.cat([g2, torch.ones_like(g2[:, :1])], -1)
    gh = (torch.linalg.inv(K3) @ gh.t()).t()
    gcal = (gh[:, :2] / gh[:, 2, None]).reshape(-1, 4)
    l_line, _ = line_loss(out["lines2d_calib"].reshape(-1, 4), gcal, wgt * (thr < 100).reshape(-1, 1))
    rgb_loss = (out["rgb_values"] - rgb_gt.reshape(-1, 3)).abs().mean()
    if "grad_theta" in out:
        eik = ((out["grad_theta"].norm(2, dim=1) - 1) ** 2).mean()
    else:
        eik = torch.zeros((), dtype=rgb_loss.dtype)
    loss = rgb_loss + eikonal_weight * eik + line_weight * l_line
    res = dict(rgb_loss=rgb_loss, eikonal_loss=eik, line_loss=l_line, l2d_loss=l2d_uncal, count=count,
               j3d_loss=torch.zeros(()), j2d_loss=torch.zeros(()), j2d_stat=torch.zeros(()),
               jcount=torch.zeros(()))
    if "j3d_local" in out and out["j3d_local"].shape[0] > 0:
        with torch.no_grad():
            c = torch.cdist(out["j3d_local"], out["j3d_global"], p=1) + \
                0.1 * torch.cdist(out["j2d_local_calib"], out["j2d_global_calib"], p=1)
        a0, a1 = linear_sum_assignment(c.detach().cpu().numpy())
        l3 = (out["j3d_local"][a0] - out["j3d_global"][a1]).abs().sum(-1).mean()
        l2 = (out["j2d_local_calib"][a0] - out["j2d_global_calib"][a1]).abs().sum(-1).mean()
        with torch.no_grad():
            l2u = (out["j2d_local"][a0] - out["j2d_global"][a1]).abs().sum(-1).mean()
        loss = loss + junction_3d_weight * l3 + junction_2d_weight * l2
        res.update(j3d_loss=l3, j2d_loss=l2, j2d_stat=l2u, jcount=(c[a0, a1] < 10).sum())
    res["loss"] = loss
    return res


# ----------------------------------------------------------------------------------------
# hand-derived backward recurrences -- what the CUDA kernels implement
# ----------------------------------------------------------------------------------------
def head_backward(Ws, zs, us, out_bar):
    """Reverse sweep of a ReLU head.  Returns (input_bar, [W_bar], [b_bar])."""
    L = len(Ws)
    gW, gb = [None] * L, [None] * L
    g = out_bar
    for l in range(L - 1, -1, -1):
        zb = g if l == L - 1 else g * (zs[l] > 0).to(g.dtype)
        gW[l] = zb.t() @ us[l]
        gb[l] = zb.sum(0)
        g = zb @ Ws[l]
    return g, gW, gb


def sdf_double_backward(P: NeatParams, x, saved, n_bar=None, o_bar=None):
    """Gradients of the SDF-network weights given
         n_bar [M,3]  = dL/d(normal)   (normal = d sdf / d x from ``sdf_outputs``; for eikonal
                        points pass saved['act']=1, i.e. no sphere clamp), and
         o_bar [M,1+F] = dL/d(raw network output) (the sdf column already masked by ``act``).
    Follows Appendix A steps 4-5.  Returns ([W_bar], [b_bar])."""
    L = len(P.sdf_W)
    zs, us, gs, act = saved["zs"], saved["us"], saved["gs"], saved["act"]
    M = x.shape[0]
    dt = x.dtype
    gW = [torch.zeros_like(W) for W in P.sdf_W]
    gb = [torch.zeros_like(b) for b in P.sdf_b]
    d_embed = us[0].shape[1]
    zhat = [None] * L
    if n_bar is not None:
        # a_l = d s_raw / d z_l  (input of the transposed layer l in the normal pass)
        a = [None] * L
        a[L - 1] = torch.zeros(M, P.sdf_W[L - 1].shape[0], dtype=dt)
        a[L - 1][:, 0] = 1.0
        for l in range(L - 2, -1, -1):
            a[l] = softplus100_d1(zs[l]) * gs[l + 1]
        # tangent sweep with seed p0 = J * (act * n_bar)
        J = embed_jacobian_diag(x, P.multires)
        nb = (act * n_bar)
        p0 = J * nb.repeat(1, d_embed // 3)
        p = p0
        for l in range(L):
            p_in = torch.cat([p, p0], 1) / math.sqrt(2.0) if l in P.skip_in else p
            q = p_in @ P.sdf_W[l].t()
            gW[l] += a[l].t() @ p_in
            if l < L - 1:
                zhat[l] = softplus100_d2(zs[l]) * gs[l + 1] * q
                p = softplus100_d1(zs[l]) * q
    g = o_bar if o_bar is not None else torch.zeros(M, P.sdf_W[L - 1].shape[0], dtype=dt)
    for l in range(L - 1, -1, -1):
        if l == L - 1:
            zb = g
        else:
            zb = softplus100_d1(zs[l]) * g
            if zhat[l] is not None:
                zb = zb + zhat[l]
        gW[l] += zb.t() @ us[l]
        gb[l] += zb.sum(0)
        g = zb @ P.sdf_W[l]
        if l in P.skip_in:
            g = g[:, :-d_embed] / math.sqrt(2.0)
    return gW, gb


# ----------------------------------------------------------------------------------------
# full forward (VolSDFNetwork.forward, neat_wfr_rend_a.py:376-538) composed from the pieces
# ----------------------------------------------------------------------------------------
@dataclass
class TrainRandoms:
    sampler: SamplerRandoms
    eik_uniform: torch.Tensor      # neat_wfr_rend_a.py:518  uniform_(-r, r) [R,3]


def l3d_candidates(lines3d, l3d):
    """neat_wfr_rend_a.py:454-455, 461-465 (use_l3d): the rays whose tangent-plane point l3d lies closer to their 3D
    line than the median ray (but at least 0.01) hand both end points, then their l3d, to the junction matching."""
    l3 = lines3d.detach().reshape(-1, 2, 3)
    l3d = l3d.detach()
    score = torch.norm(torch.cross(l3d - l3[:, 0], l3d - l3[:, 1], dim=-1), dim=-1) / torch.norm(l3[:, 0] - l3[:, 1], dim=-1)
    med = score.median()
    thr = med if float(med) > 0.01 or bool(torch.isnan(med)) else torch.tensor(0.01, dtype=score.dtype)   # python max(median, 0.01)
    sel = score < thr
    return torch.cat((l3[sel].reshape(-1, 3), l3d[sel]), dim=0), score


def junction_block(P: NeatParams, K, pose, lines3d, gt_vertices, l3d=None):
    """neat_wfr_rend_a.py:457-496: candidates = DBSCAN centroids (dbscan_enabled, :459-460), the l3d selection (use_l3d,
    :461-465) or every attraction end point (:466-467); match filter < 10 px or < median matched cost (use_median,
    :475-482)."""
    from scipy.optimize import linear_sum_assignment
    dt = lines3d.dtype
    Rm, T = pose_inverse_rt(pose)
    K3 = K[:3, :3]
    I3 = torch.eye(3, dtype=dt)
    if P.dbscan_enabled:
        cent = dbscan_centroids(lines3d.detach().cpu().numpy().reshape(-1, 3), eps=0.01, min_samples=2)
        j3d = torch.tensor(cent).float().to(dt).reshape(-1, 3)
    elif P.use_l3d:
        j3d, _ = l3d_candidates(lines3d, l3d)
    else:
        j3d = lines3d.detach().reshape(-1, 3)
    j2d = project2d(K3, Rm, T, j3d)
    j2d_cal = project2d(I3, Rm, T, j3d)
    gt = gt_vertices.to(dt)
    cost = ((j2d[None] - gt[:, None]) ** 2).sum(-1).sqrt()
    a0, a1 = linear_sum_assignment(cost.detach().cpu().numpy())
    extra = {}
    if P.use_median:
        median = cost[a0, a1].detach().median()
        if torch.isnan(median):
            median = torch.tensor(10, dtype=torch.float32)
        ok = cost[a0, a1] < median
        extra["median"] = median
    else:
        ok = cost[a0, a1] < 10
    glob = junction_ffn(P)
    return dict(j3d_local=j3d[a1][ok], j2d_local=j2d[a1][ok], j2d_local_calib=j2d_cal[a1][ok],
                j3d_global=glob, j2d_global=project2d(K3, Rm, T, glob),
                j2d_global_calib=project2d(I3, Rm, T, glob), **extra)


def neat_forward(P: NeatParams, sconf: SamplerConf, K, pose, uv, uv_proj, gt_vertices=None,
                 training=False, rnd: Optional[TrainRandoms] = None, samples=None):
    """K [4,4], pose [4,4], uv [R,2], uv_proj [R,2].  Output dict mirrors the reference's.
    samples=(z_vals, z_eik) bypasses the error-bound sampler (whose discrete decisions make the sample
    positions sensitive to 1e-6-level differences in the SDF) so that everything downstream can be
    compared at identical sample positions."""
    dirs, cam = camera_rays(uv, pose, K)
    R = dirs.shape[0]
    camr = cam[None].expand(R, 3)
    if samples is not None:
        z_vals, z_eik, k = samples[0], samples[1], -1
    else:
        z_vals, z_eik, k = error_bound_sampler(P, sconf, dirs, camr, training=training,
                                               rnd=rnd.sampler if training else None)
    rr = render_rays(P, dirs, camr, z_vals)
    geo = line_geometry(P, K, pose, uv_proj, rr["points3d"], rr["lines3d"])
    out = dict(points=rr["points"], rgb_values=rr["rgb_values"], depth=rr["depth"], xyz=rr["xyz"],
               points3d=rr["points3d"], lines3d=rr["lines3d"], l3d=geo["l3d"],
               lines2d=geo["lines2d"], lines2d_calib=geo["lines2d_calib"], sdf=geo["sdf"],
               K=K[:3, :3], z_vals=z_vals, weights=rr["weights"], n_sampler_iters=k)
    if training:
        out.update(junction_block(P, K, pose, rr["lines3d"], gt_vertices, l3d=geo["l3d"]))
        near = camr + z_eik * dirs
        eik_pts = torch.cat([rnd.eik_uniform.to(dirs.dtype), near], 0)
        if P.junction_eikonal:                                # neat_wfr_rend_a.py:524-525
            eik_pts = torch.cat([eik_pts, out["j3d_global"].detach()], 0)
        _, _, g, _ = sdf_outputs(P, eik_pts, clamp=False)
        out["grad_theta"] = g
        out["eik_points"] = eik_pts
    else:
        out["normal_map"] = rr["normal_map"]
    return out
