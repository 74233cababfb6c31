"""GPU parity at the sizes the benchmark runs (VERDICT r01 "Next round" #1): BASELINE configs[1] (1024 rays x 98 samples =
784 tiles = 5.3 tiles per persistent CTA, 16 wgrad splits) against the oracle at identical samples, configs[2] (8192 rays,
6272 tiles, 32 splits) through a size-independent property (the parameter gradient of a batch = the sum over its ray
chunks) plus the oracle on a ray subset, and the several-tiles-per-CTA loops forced at a small size with the grid cap.
Reference: code/training/volsdf_train.py:366-374 (model -> loss -> backward)."""
import numpy as np
import pytest
import torch

import golden_io as G
import parity_util as PU
from neat_b200 import synth

pytestmark = pytest.mark.gpu
T = PU.T


@pytest.mark.parametrize("beta", [0.1, 0.01])
def test_train_step_1024_rays_vs_oracle(beta):
    """The benchmarked configuration: outputs and loss terms to 1e-4, EVERY parameter gradient to rel-L2 1e-3 and
    max-abs 2e-3 of the tensor's largest entry (density.beta 1e-2), against autograd through the oracle."""
    conf = synth.dtu_conf()
    sd_np = synth.make_state_dict(conf, seed=5, perturb=0.15, beta=beta)
    model = PU.make_model(conf, sd_np)
    b = synth.make_batch(1024, seed=4)
    out, lo = PU.gpu_step(model, b)
    st = model.last_step
    assert st.R == 1024 and st.S == 98
    k = int(st.n_iters.item())
    assert (k >= 4) if beta == 0.01 else (k >= 1)      # beta 0.01: the steady-state 5-iteration sampler
    oo, ol, leaves = PU.oracle_step(conf, sd_np, b, st)
    for key, e in PU.output_errors(out, oo).items():
        assert e < 1e-4, (key, e)
    for key, e in PU.loss_errors(lo, ol).items():
        assert e < 1e-4, (key, e)
    table = PU.grad_errors({n: p.grad for n, p in model.named_parameters()}, {n: v.grad for n, v in leaves.items()})
    assert len(table) >= 50
    PU.assert_grads(table)


def test_eval_forward_1024_rays_vs_oracle():
    """Eval-mode forward INCLUDING the sampler at 1024 rays, beta = 0.01 (5 sampler iterations).  The sampler is chaotic at
    the 1e-4 level BY CONSTRUCTION of the algorithm: its per-ray beta bisection and inverse-CDF search are discrete, so
    rounding noise moves samples -- the oracle evaluated in float32 and in float64 (the same code, both exact restatements
    of code/model/ray_sampler.py:130-283) agree on the samples of < 1 % of the rays and disagree by > 1e-4 on the composited
    geometry of ~1 % of them, up to 2e-3 (scripts/measure_parity.py -> profiles/r02_parity_measured.json: 10 of 1024 rays
    for lines3d; the kernels: 49 of 1024, up to 1.3e-3 -- their SDF queries carry ~3e-6 of error against the fp32
    reference's ~1e-6, and the chaos amplifies the difference).  So the statement is statistical, with the float64 oracle as
    the truth: colours of EVERY ray within 1e-4; for the geometry outputs the median ray within 1e-5, 90 % of the rays within
    1e-4, at most 8 % beyond 1e-4 and none beyond 2e-2; the 1e-4 bound at identical samples is
    test_train_step_1024_rays_vs_oracle's."""
    from neat_b200.context import Context
    from neat_b200.render import Renderer
    from oracle import neat_oracle as O
    conf = synth.dtu_conf()
    sd_np = synth.make_state_dict(conf, seed=5, perturb=0.15, beta=0.01)
    ctx = Context(conf)
    sd = {k: torch.from_numpy(v).cuda() for k, v in sd_np.items()}
    ctx.pack_weights(ctx.flatten_state_dict(sd))
    rn = Renderer(ctx, conf)
    b = synth.make_batch(1024, seed=9)
    out = rn.forward_eval(T(b["uv"][0]).cuda(), T(b["pose"][0]).cuda(), T(b["intrinsics"][0]).cuda(),
                          T(b["uv_proj"][0]).cuda(), sd["density.beta"].reshape(1))
    P, _ = G.oracle_params(conf, sd_np, dtype=torch.float64)
    D = lambda a: T(a).double()
    ref = O.neat_forward(P, G.sampler_conf(conf), D(b["intrinsics"][0]), D(b["pose"][0]), D(b["uv"][0]),
                         D(b["uv_proj"][0]), training=False)
    assert int(out["n_sampler_iters"].item()) == ref["n_sampler_iters"] == 5
    z = out["z_vals"].cpu()
    assert bool((z[:, 1:] >= z[:, :-1]).all())
    for key in ("rgb_values", "depth", "points3d", "lines3d", "lines2d", "lines2d_calib", "normal_map"):
        got, want = out[key].cpu().numpy().astype(np.float64), ref[key].numpy()
        err = np.abs(got - want).reshape(1024, -1).max(axis=1) / np.abs(want).max()
        if key in ("rgb_values", "lines2d", "lines2d_calib"):
            assert err.max() < 1e-4 if key == "rgb_values" else err.max() < 1e-3, (key, err.max())
        assert np.median(err) < 1e-5, (key, np.median(err))
        assert np.quantile(err, 0.9) < 1e-4, (key, np.quantile(err, 0.9))
        assert (err > 1e-4).mean() <= 0.08, (key, (err > 1e-4).mean())
        assert err.max() < 2e-2, (key, err.max())


def test_multiple_tiles_per_cta_and_forced_wgrad_splits():
    """130 DTU rays = 100 tiles on 8 persistent CTAs (12-13 tiles each: mbarrier phases, TMEM halves and save-record
    indexing carried across tiles) with the weight-gradient GEMMs forced into 16 tile-range splits."""
    conf = synth.dtu_conf()
    sd_np = synth.make_state_dict(conf, seed=6, perturb=0.15, beta=0.1)
    model = PU.make_model(conf, sd_np)
    rn = model._get_renderer()
    rn.ctx.debug_grid_cap(8)
    rn.ctx.debug_wgrad_split(16, 4)
    try:
        b = synth.make_batch(130, seed=8)
        out, lo = PU.gpu_step(model, b)
        st = model.last_step
        oo, ol, leaves = PU.oracle_step(conf, sd_np, b, st)
        for key, e in PU.output_errors(out, oo).items():
            assert e < 1e-4, (key, e)
        for key, e in PU.loss_errors(lo, ol).items():
            assert e < 1e-4, (key, e)
        PU.assert_grads(PU.grad_errors({n: p.grad for n, p in model.named_parameters()},
                                       {n: v.grad for n, v in leaves.items()}))
        # the capped grid and the full grid run the same arithmetic per tile: identical outputs
        rn.ctx.debug_grid_cap(0)
        rn.ctx.debug_wgrad_split(0, 0)
        g_capped = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        out2, _ = PU.gpu_step(model, b)
        for key in ("rgb_values", "lines3d", "grad_theta"):
            assert torch.equal(out[key], out2[key]), key
        for n, p in model.named_parameters():
            if n in g_capped:
                # other split boundaries = another grouping of the (round-toward-zero) TMEM accumulation and of the
                # fp32 atomics: summation-order noise only (measured 2e-5)
                assert G.rel_err(p.grad.cpu(), g_capped[n].cpu()) < 1e-4, n
    finally:
        rn.ctx.debug_grid_cap(0)
        rn.ctx.debug_wgrad_split(0, 0)


def test_8192_rays_gradient_is_sum_of_chunks_and_subset_vs_oracle():
    """BASELINE configs[2] size (8192 rays = 6272 tiles = 42 tiles per CTA, 32 wgrad splits).  (a) linearity: with the
    upstream gradients d loss / d (rgb_values, lines3d, grad_theta) held fixed, the parameter gradient of the batch equals
    the sum over its eight 1024-ray chunks (each of which is the configuration checked against the oracle above), at
    identical samples; (b) the rendered outputs of a 256-ray subset against the oracle at the same samples."""
    from oracle import neat_oracle as O
    conf = synth.dtu_conf()
    sd_np = synth.make_state_dict(conf, seed=5, perturb=0.15, beta=0.1)
    model = PU.make_model(conf, sd_np)
    R, C = 8192, 1024
    b = synth.make_batch(R, seed=4)
    out, lo = PU.gpu_step(model, b, keep_bars=True)
    st = model.last_step
    assert st.R == R
    names = [n for n, _ in model.named_parameters() if n.split(".")[0] in
             ("implicit_network", "rendering_network", "attraction_network", "density")]
    params = dict(model.named_parameters())
    full = {n: params[n].grad.detach().clone() for n in names}
    bars = {k: out[k].grad.detach().clone() for k in ("rgb_values", "lines3d", "grad_theta")}
    z, z_eik = st.z.detach().clone(), ((st.eik_pts[R:] - st.cam[None]) * st.dirs).sum(1, keepdim=True)
    eik_u = st.eik_uniform.detach().clone()
    got_full = {k: out[k].detach().clone() for k in ("rgb_values", "lines3d", "grad_theta", "depth", "points3d")}
    for p in model.parameters():
        p.grad = None
    for c0 in range(0, R, C):
        sl = slice(c0, c0 + C)
        inp, _ = PU.device_inputs(b, sl)
        model.replay = dict(samples=(z[sl], z_eik[sl]), eik_uniform=eik_u[sl])
        o = model(inp)
        for k in ("rgb_values", "lines3d", "depth", "points3d"):   # per-ray outputs do not depend on the batch
            assert G.rel_err(o[k].detach().cpu(), got_full[k][sl].cpu()) < 1e-6, k
        gt_bar = torch.cat([bars["grad_theta"][sl], bars["grad_theta"][R + c0:R + c0 + C]], 0)
        torch.autograd.backward([o["rgb_values"], o["lines3d"], o["grad_theta"]],
                                [bars["rgb_values"][sl], bars["lines3d"][sl], gt_bar])
    torch.cuda.synchronize()
    model.replay = None
    table = PU.grad_errors({n: full[n] for n in names}, {n: params[n].grad for n in names})
    PU.assert_grads(table, tol_l2=1e-4, tol_max=2e-4)
    # (b) a ray subset against the oracle (forward only)
    P, _ = G.oracle_params(conf, sd_np)
    idx = torch.arange(0, R, R // 256)[:256]
    dirs, cam = O.camera_rays(T(b["uv"][0])[idx], T(b["pose"][0]), T(b["intrinsics"][0]))
    rr = O.render_rays(P, dirs, cam[None].expand(len(idx), 3), z.cpu()[idx])
    for k in ("rgb_values", "lines3d", "depth", "points3d"):
        assert G.rel_err(got_full[k].cpu()[idx], rr[k].detach().reshape(got_full[k][idx].shape)) < 1e-4, k
