"""CPU: the hand-derived backward recurrences spelled out in the oracle (what the CUDA kernels
implement) equal torch.autograd on the oracle's forward restatement, in float64."""
import numpy as np
import pytest
import torch

import golden_io as G
from neat_b200 import synth
from oracle import neat_oracle as O


def _params(conf_fn, seed=2):
    conf = conf_fn()
    sd_np = synth.make_state_dict(conf, seed=seed, perturb=0.15, beta=0.1)
    P, leaves = G.oracle_params(conf, sd_np, dtype=torch.float64, track=True)
    return conf, P, leaves


def _close(a, b, tol=1e-8):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("clamp", [True, False])
def test_analytic_normal_equals_autograd(clamp):
    """Appendix A step 2 (sdf_outputs) vs autograd.grad, incl. the sphere-clamp branch of min()."""
    _, P, _ = _params(synth.toy_conf)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(300, 3, generator=g, dtype=torch.float64) - 0.5) * 7.0   # reaches beyond the bounding sphere (r = 3)
    sdf, feat, grad, saved = O.sdf_outputs(P, x, clamp=clamp)
    if clamp:
        s_ref, f_ref, g_ref, _ = O.sdf_outputs_autograd(P, x)
        assert float(saved["act"].mean()) < 1.0  # the clamp is exercised
    else:
        g_ref = O.sdf_gradient_autograd(P, x)
        s_ref, f_ref = O.sdf_forward(P, x)[:, :1], O.sdf_forward(P, x)[:, 1:]
    assert _close(sdf, s_ref) and _close(feat, f_ref) and _close(grad, g_ref)


def test_sdf_double_backward_equals_autograd():
    """Appendix A steps 4-5 (tangent + reverse sweep): weight / bias gradients of L = <n_bar, normal> + <o_bar, out>
    vs autograd through autograd.grad(create_graph=True), the reference's double backward."""
    _, P, leaves = _params(synth.toy_conf)
    g = torch.Generator().manual_seed(1)
    M = 200
    x = (torch.rand(M, 3, generator=g, dtype=torch.float64) - 0.5) * 7.0
    n_bar = torch.randn(M, 3, generator=g, dtype=torch.float64)
    F = P.sdf_W[-1].shape[0] - 1
    o_bar = torch.randn(M, 1 + F, generator=g, dtype=torch.float64)
    sdf, feat, grad, saved = O.sdf_outputs(P, x, clamp=True)
    act = saved["act"]
    ob = torch.cat([o_bar[:, :1] * act, o_bar[:, 1:]], 1)          # the sdf column reaches the network only where it is active
    gW, gb = O.sdf_double_backward(P, x, saved, n_bar=n_bar, o_bar=ob)
    # autograd reference
    s_a, f_a, g_a, _ = O.sdf_outputs_autograd(P, x, create_graph=True)
    loss = (g_a * n_bar).sum() + (s_a * o_bar[:, :1]).sum() + (f_a * o_bar[:, 1:]).sum()
    wn = [k for k in leaves if k.startswith("implicit_network")]
    grads = torch.autograd.grad(loss, [leaves[k] for k in wn], allow_unused=True, retain_graph=True)
    # compare through the weight_norm chain: effective-weight gradients -> (g, v, b) gradients by autograd of the chain only
    eff = []
    for l in range(len(P.sdf_W)):
        eff += [(P.sdf_W[l], gW[l]), (P.sdf_b[l], gb[l])]
    chain = torch.autograd.grad([t for t, _ in eff], [leaves[k] for k in wn], [gbar for _, gbar in eff], allow_unused=True)
    for k, a, b in zip(wn, chain, grads):
        assert (a is None) == (b is None), k
        if a is not None:
            assert _close(a, b, 1e-7), (k, float((a - b).abs().max()))


def test_head_backward_equals_autograd():
    g = torch.Generator().manual_seed(2)
    dims = [41, 32, 32, 6]
    Ws = [torch.randn(dims[i + 1], dims[i], generator=g, dtype=torch.float64, requires_grad=True) for i in range(3)]
    bs = [torch.randn(dims[i + 1], generator=g, dtype=torch.float64, requires_grad=True) for i in range(3)]
    inp = torch.randn(50, 41, generator=g, dtype=torch.float64, requires_grad=True)
    out, zs, us = O.head_forward(Ws, bs, inp, save=True)
    ob = torch.randn(50, 6, generator=g, dtype=torch.float64)
    gi, gW, gb = O.head_backward(Ws, zs, us, ob)
    ref = torch.autograd.grad((out * ob).sum(), [inp] + Ws + bs)
    assert _close(gi, ref[0])
    for a, b in zip(gW + gb, ref[1:]):
        assert _close(a, b)


def test_volume_weights_backward_equals_autograd():
    g = torch.Generator().manual_seed(3)
    R, S = 40, 98
    z = torch.sort(torch.rand(R, S, generator=g, dtype=torch.float64) * 6, dim=1).values
    z[:, 5] = z[:, 4]                                                # zero-length intervals occur in real batches
    sdf = (torch.randn(R, S, generator=g, dtype=torch.float64) * 0.3).requires_grad_(True)
    beta = torch.tensor(0.07, dtype=torch.float64, requires_grad=True)
    w_bar = torch.randn(R, S, generator=g, dtype=torch.float64)
    w = O.volume_weights(z, sdf, beta)
    ref_s, ref_b = torch.autograd.grad((w * w_bar).sum(), [sdf, beta])
    got_s, got_b = O.volume_weights_backward(z, sdf.detach(), beta.detach(), w_bar)
    assert _close(got_s, ref_s, 1e-7) and abs(float(got_b) - float(ref_b)) <= 1e-7 * max(1.0, abs(float(ref_b)))
