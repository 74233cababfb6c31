"""The training step of VolSDFNetwork.forward over the C-ABI kernels: two plain functions (step_forward / step_backward:
launch sequences only, no autograd, graph-capturable) and the torch.autograd.Function that the plugin wraps around them.

forward : camera rays -> error-bound sampler -> ImplicitNetwork (sdf, analytic normals, features) ->
          rendering / attraction heads -> compositing -> second get_outputs at the surface points ->
          2D line geometry -> eikonal points (reference: neat_wfr_rend_a.py:376-538)
backward: compositing adjoint -> head reverse sweeps -> ImplicitNetwork double backward (tangent + reverse
          sweeps) -> weight-gradient GEMMs            (reference: loss.backward())

Differentiable inputs: density.beta and the (weight_g, weight_v, bias) parameters of every MLP layer (weight_norm and
its adjoint are one kernel launch each).  Differentiable outputs:
rgb_values [R,3], lines3d [R,2,3], grad_theta [2R,3]; everything else the reference returns is detached
there as well or does not reach a loss (depth, xyz, points3d, sdf, l3d, lines2d)."""
import ctypes

import torch

from . import _lib

_P = ctypes.c_void_p


def _ptr(t):
    return _P(t.data_ptr()) if t is not None else None


class StepState:
    """Non-differentiable inputs and every saved buffer of one step (kept alive until backward ran)."""
    samples_override = None
    sampler_randoms = None
    eik_uniform = None
    junction_inputs = None
    dbscan_enabled = True


def step_forward(renderer, st, beta):
    """st: uv, pose, K, uv_proj, param_layers (+ optional replay fields); beta: density.beta as a [1] device tensor.
    Enqueues the whole training forward; returns (rgb_values [R,3], lines3d [R,6], grad_theta [2R,3]).  The junction
    hand-over (device -> pinned host copies of the cluster centroids etc.) is enqueued right after the attraction head:
    st.junction_event / st.junction_host."""
    ctx = renderer.ctx
    dev = ctx.device
    # the save records live in named, reused workspaces (render.WorkspacePool): one step in flight at a time.
    # backward checks that no later forward has overwritten them instead of silently using the wrong records.
    renderer.generation += 1
    st.generation = renderer.generation
    for lay in st.param_layers:
        for t in lay:
            # the MLP parameters are not autograd inputs of the step (the backward writes their .grad itself), so
            # tensor hooks on them would silently never fire: refuse instead ("autograd contract")
            if t is not None and (t._backward_hooks or getattr(t, "_post_accumulate_grad_hooks", None)):
                raise _lib.NeatError("hooks on the MLP parameters are not supported: the step writes p.grad directly "
                                     "(use neat_b200.parallel.GradBucket for data parallelism, not DDP hooks)")
    layers = [tuple(None if t is None else t.detach() for t in lay) for lay in st.param_layers]
    st.wn_layers = layers
    renderer.effective_weights(layers)
    st.beta = beta
    uv, pose, K = st.uv, st.pose, st.K
    dirs, cam = renderer.camera_rays(uv, pose, K)
    R = uv.shape[0]
    device_draws = (st.samples_override is None and st.sampler_randoms is None and renderer.sampler.rng == "device")
    if device_draws:
        # every random draw of the step in ONE launch (csrc/train_aux.cuh), no eager PyTorch RNG kernels
        dr = renderer.draws(R)
        st.sampler_randoms = dr
        if st.eik_uniform is None:
            st.eik_uniform = dr["eik_uniform"]
    if st.samples_override is not None:
        # tests: (z_vals [R,S], z_eik [R,1]) handed in, e.g. to replay one batch in ray chunks at identical samples
        z, z_eik = (t.to(dev, torch.float32).contiguous() for t in st.samples_override)
        n_it = renderer.pool.zeros("step.n_it", 1, torch.int32)
    else:
        z, z_eik, n_it = renderer.sampler.get_z_vals(cam, dirs, beta, training=True, randoms=st.sampler_randoms,
                                                     device_tables=device_draws)
    R, S = z.shape
    M = R * S
    st.R, st.S, st.dirs, st.cam, st.z, st.n_iters = R, S, dirs, cam, z, n_it
    # eikonal points (neat_wfr_rend_a.py:515-527): R uniform in the bounding cube + R near-surface; they only need the
    # sampler's output
    if st.eik_uniform is None:
        r = renderer.scene_bounding_sphere
        st.eik_uniform = torch.empty(R, 3).uniform_(-r, r).to(dev)  # the reference's CPU-generator draw
    near = cam[None, :] + z_eik * dirs
    parts = [st.eik_uniform.to(dev, torch.float32), near]
    if getattr(st, "junction_eikonal", False):       # neat_wfr_rend_a.py:524-525: + the (detached) global junctions
        parts.append(st.junction_inputs[0].detach().to(dev, torch.float32))
    st.eik_pts = torch.cat(parts, 0).contiguous()
    st.n_eik = st.eik_pts.shape[0]
    pe = renderer.explicit_points(st.eik_pts)
    # Stream plan.  Every tile-MLP launch is 148 persistent CTAs (one per SM) whose last round of tiles leaves most
    # SMs idle, and the small launches (eikonal points: 16 tiles, surface points: 8 tiles) would each occupy a
    # handful of SMs for the latency of a whole tile.  They go to a side stream, where the block scheduler fits them
    # into the tails of the big launches:
    #   main : render -> attraction head -> line compositing -> rendering head -> colours
    #   side : eikonal points (forked before the render launch) ... DBSCAN -> host hand-over -> surface points -> geometry
    # The big launches stay serialised (running the two heads concurrently was measured: no gain at 1024 rays, 3 %
    # slower at 8192).  The junction hand-over only needs the attraction head, so the host-side matching overlaps the
    # rendering head (plugin path) or the whole backward (FusedTrainStep) instead of an idle GPU.
    main, side = torch.cuda.current_stream(dev), renderer.side_stream()
    fork = torch.cuda.Event()
    fork.record(main)
    pts = renderer.ray_points(cam, dirs, z)
    st.sdf, st.grad, st.act, st.feat, st.sdf_save = renderer.sdf_outputs(pts, M, clamp=True, training=True, tag="render")
    with torch.cuda.stream(side):
        side.wait_event(fork)
        _, grad_theta, _, _, st.eik_save = renderer.sdf_outputs(pe, st.n_eik, clamp=False, training=True,
                                                                want_feat=False, want_sdf=False, tag="eik")
    grad_theta.record_stream(main)
    st.lines, st.att_save = renderer.head_forward(1, pts, M, st.grad, st.feat, training=True)
    w, lines3d, depth, points3d = renderer.composite_lines(z, st.sdf, st.lines, cam, dirs, beta)
    st.weights, st.depth, st.points3d = w, depth, points3d
    lines_done = torch.cuda.Event()
    lines_done.record(main)
    st.rgb, st.rend_save = renderer.head_forward(0, pts, M, st.grad, st.feat, training=True)
    with torch.cuda.stream(side):
        side.wait_event(lines_done)
        # junction candidates first (the host is waiting for them): DBSCAN is ~100 us of tiny launches (8-block grids),
        # which used to sit on the main stream between the two heads (ncu launch list, profiles/r02_*)
        use_l3d = st.junction_inputs is not None and not st.dbscan_enabled and getattr(st, "use_l3d", False)

        def hand_over(cent_d, n_d):
            st.junction_event, st.junction_host = renderer.to_host_async([n_d, cent_d] + list(st.junction_inputs),
                                                                         counter=getattr(st, "handover_counter", None))

        if st.junction_inputs is not None and not use_l3d:
            if st.dbscan_enabled:
                cent_d, n_d = renderer.dbscan_async(lines3d.view(-1, 3), 0.01)
            else:  # abc-neat-a.conf: every attraction end point is a junction candidate (neat_wfr_rend_a.py:465-466)
                cent_d = lines3d.view(-1, 3)
                n_d = renderer.pool.get("step.n_all", 1, torch.int32)
                n_d.fill_(2 * R)
            hand_over(cent_d, n_d)
        p3 = renderer.explicit_points(points3d)
        st.sdf3, st.grad3, _, _, _ = renderer.sdf_outputs(p3, R, clamp=True, want_feat=False, tag="surface")
        st.lines2d, st.lines2d_calib, st.l3d, st.pose_inv = renderer.line_geometry(pose, K, st.uv_proj, points3d,
                                                                                   st.grad3, lines3d)
        if use_l3d:  # the candidates need the tangent-plane points l3d (neat_wfr_rend_a.py:461-465)
            hand_over(*renderer.l3d_candidates_async(lines3d, st.l3d))
        side_done = torch.cuda.Event()
        side_done.record(side)
    for t in (st.sdf3, st.grad3, st.lines2d, st.lines2d_calib, st.l3d, st.pose_inv):
        t.record_stream(main)
    rgb_values = renderer.composite_rgb(z, st.sdf, st.rgb, cam, dirs, beta)
    main.wait_event(side_done)
    return rgb_values, lines3d, grad_theta


def step_backward(renderer, st, rvb, l3b, gtb, beta_bar, targets, accumulate):
    """rvb [R,3] = dL/d rgb_values, l3b [R,6] = dL/d lines3d, gtb [2R,3] = dL/d grad_theta (contiguous fp32 device
    tensors); beta_bar [1]: dL/d density.beta is ADDED to it; targets: per layer (gg | None, gv, gb) gradient tensors
    the weight_norm adjoint writes (accumulate=False) or adds to (True)."""
    if st.generation != renderer.generation:
        raise _lib.NeatError("backward() of a training forward whose saved activations were overwritten by a later "
                             "forward of the same module: call loss.backward() before the next model(...) call "
                             "(one step in flight at a time, as in code/training/volsdf_train.py:366-374)")
    ctx = renderer.ctx
    lib, dev = ctx.lib, ctx.device
    R, S = st.R, st.S
    M = R * S
    stream = ctx._stream()
    pool = renderer.pool
    main, side = torch.cuda.current_stream(dev), renderer.side_stream()
    fork = torch.cuda.Event()  # the eikonal points' backward depends on grad_theta_bar only
    fork.record(main)
    rgb_pre_bar = pool.get("bwd.rgb_pre_bar", M * 3).view(M, 3)
    lines_bar = pool.get("bwd.lines_bar", M * 6).view(M, 6)
    sdf_bar = pool.get("bwd.sdf_bar", M)
    a = _lib.CompositeBwdArgs(R, S, _ptr(st.z), _ptr(st.sdf), _ptr(st.weights), _ptr(st.rgb), _ptr(st.act), _ptr(rvb),
                              _ptr(l3b), _ptr(st.beta), renderer.beta_min, _ptr(rgb_pre_bar), _ptr(lines_bar),
                              _ptr(sdf_bar), _ptr(beta_bar), _ptr(renderer.bg_color))
    _lib.check(lib.neat_composite_backward(ctypes.byref(a), stream))
    feat_bar = pool.get("bwd.feat_bar", int(lib.neat_feat_bar_bytes(M)) // 4)
    n_bar = pool.get("bwd.n_bar", M * 3).view(M, 3)
    hb = [pool.get("bwd.head%d" % h, int(lib.neat_head_bwd_save_bytes(ctx._h, M)), torch.uint8) for h in range(2)]
    with renderer.timed("head_bwd"):
        _lib.check(lib.neat_head_backward(ctx._h, 0, M, _ptr(rgb_pre_bar), _ptr(st.rend_save), _ptr(hb[0]),
                                          _ptr(feat_bar), _ptr(n_bar), 0, stream))
        _lib.check(lib.neat_head_backward(ctx._h, 1, M, _ptr(lines_bar), _ptr(st.att_save), _ptr(hb[1]),
                                          _ptr(feat_bar), _ptr(n_bar), 1, stream))
    pts = renderer.ray_points(st.cam, st.dirs, st.z)
    sb = pool.get("bwd.sdf", int(lib.neat_sdf_bwd_save_bytes(ctx._h, M)), torch.uint8)
    scratch = pool.get("bwd.scratch", int(lib.neat_sdf_bwd_scratch_bytes(ctx._h, M)), torch.uint8)
    with renderer.timed("sdf_bwd_M%d" % M):
        _lib.check(lib.neat_sdf_backward(ctx._h, ctypes.byref(pts), _ptr(n_bar), _ptr(sdf_bar), _ptr(feat_bar),
                                         _ptr(st.act), _ptr(st.sdf_save), _ptr(sb), _ptr(scratch), stream))
    # the eikonal points' double backward fills the tails of the launches above (side stream, own scratch; it was
    # forked at the top of backward)
    pe = renderer.explicit_points(st.eik_pts)
    sbe = pool.get("bwd.sdf_eik", int(lib.neat_sdf_bwd_save_bytes(ctx._h, st.n_eik)), torch.uint8)
    scratch_e = pool.get("bwd.scratch_eik", int(lib.neat_sdf_bwd_scratch_bytes(ctx._h, st.n_eik)), torch.uint8)
    with torch.cuda.stream(side):
        side.wait_event(fork)
        _lib.check(lib.neat_sdf_backward(ctx._h, ctypes.byref(pe), _ptr(gtb), None, None, None, _ptr(st.eik_save),
                                         _ptr(sbe), _ptr(scratch_e), ctx._stream()))
        eik_done = torch.cuda.Event()
        eik_done.record(side)
    gtb.record_stream(side)
    main.wait_event(eik_done)
    # persistent: the weight-gradient job table holds this address (a fresh allocation per step would re-upload it)
    flat_grad = pool.zeros("bwd.flat_grad", ctx.n_params)
    groups = (_lib.GradGroup * 2)()
    groups[0] = _lib.GradGroup(M, _ptr(st.sdf_save), _ptr(sb), _ptr(st.feat),
                               (_P * 2)(st.rend_save.data_ptr(), st.att_save.data_ptr()),
                               (_P * 2)(hb[0].data_ptr(), hb[1].data_ptr()))
    groups[1] = _lib.GradGroup(st.n_eik, _ptr(st.eik_save), _ptr(sbe), None, (_P * 2)(None, None), (_P * 2)(None, None))
    with renderer.timed("wgrad"):
        _lib.check(lib.neat_weight_gradients(ctx._h, groups, 2, _ptr(flat_grad), stream))
    st.debug = dict(rgb_pre_bar=rgb_pre_bar, lines_bar=lines_bar, sdf_bar=sdf_bar, n_bar=n_bar, feat_bar=feat_bar)
    renderer.weight_norm_backward(st.wn_layers, flat_grad, targets, accumulate)


class NeatStepFunction(torch.autograd.Function):
    @staticmethod
    def forward(fctx, beta_param, renderer, st):
        """The MLP parameters are NOT autograd inputs: st.param_layers = [(weight_g | None, weight_v | weight, bias)]
        holds the nn.Parameters, and backward() writes (or adds to) their .grad itself in ONE kernel launch, instead of
        handing 57 tensors to 57 AccumulateGrad nodes (one elementwise launch each).  density.beta is the graph anchor."""
        beta = beta_param.detach().reshape(1).contiguous()
        rgb_values, lines3d, grad_theta = step_forward(renderer, st, beta)
        fctx.renderer, fctx.st = renderer, st
        fctx.beta_shape = beta_param.shape
        return rgb_values, lines3d.view(st.R, 2, 3), grad_theta

    @staticmethod
    def backward(fctx, rgb_values_bar, lines3d_bar, grad_theta_bar):
        renderer, st = fctx.renderer, fctx.st
        dev = renderer.ctx.device
        R = st.R
        z = lambda *s: torch.zeros(*s, device=dev)
        rvb = rgb_values_bar.contiguous().float() if rgb_values_bar is not None else z(R, 3)
        l3b = lines3d_bar.reshape(R, 6).contiguous().float() if lines3d_bar is not None else z(R, 6)
        gtb = grad_theta_bar.contiguous().float() if grad_theta_bar is not None else z(st.n_eik, 3)
        beta_bar = torch.zeros(1, device=dev)
        # parameter gradients: straight into p.grad (allocated here if the caller cleared it, added to otherwise)
        with torch.no_grad():
            have = [p.grad is not None for lay in st.param_layers for p in lay if p is not None and p.requires_grad]
            targets = []
            for lay in st.param_layers:
                tg = []
                for p in lay:
                    if p is None:
                        tg.append(None)
                    elif not p.requires_grad:
                        tg.append(torch.empty_like(p))  # frozen parameter: computed and dropped
                    else:
                        if p.grad is None:
                            p.grad = torch.zeros_like(p) if (have and any(have)) else torch.empty_like(p)
                        elif not (p.grad.is_contiguous() and p.grad.dtype == torch.float32):
                            raise _lib.NeatError("parameter gradients must be contiguous fp32 tensors")
                        tg.append(p.grad)
                targets.append(tuple(tg))
            step_backward(renderer, st, rvb, l3b, gtb, beta_bar, targets, bool(have) and any(have))
        return beta_bar.reshape(fctx.beta_shape), None, None
