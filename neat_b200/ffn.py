"""The junction `ffn` (code/model/networks/neat_wfr_rend_a.py:274-303, applied at :488): nn.Sequential(Linear, ReLU, ...,
Linear) on the [num_junctions, dim_hidden] latents -> the global 3D junctions.  Forward and backward run on the library's
own fp32 GEMM (csrc/train_aux.cuh: gemm_f32_kernel / colsum_f32_kernel) -- no cuBLAS launch is left on the step.  The
nn.Sequential stays the owner of the parameters (state_dict keys `ffn.{0,2,4}.{weight,bias}` as in the reference)."""
import ctypes

import torch

from . import _lib

_P = ctypes.c_void_p


def _ptr(t):
    return _P(t.data_ptr()) if t is not None else None


def _stream(dev):
    return _P(torch.cuda.current_stream(dev).cuda_stream)


def linears(ffn):
    return [m for m in ffn if isinstance(m, torch.nn.Linear)]


def forward(X, Ws, bs, acts):
    """X [M,K0]; Ws[i] [out_i, in_i], bs[i] [out_i]; acts[i] [M, out_i] receives layer i's output (ReLU'd except the last)."""
    lib = _lib.load()
    h = X
    n = len(Ws)
    for i, (W, b) in enumerate(zip(Ws, bs)):
        M, K, N = h.shape[0], h.shape[1], W.shape[0]
        _lib.check(lib.neat_gemm_f32(_ptr(h), _ptr(W), _ptr(acts[i]), M, N, K, K, K, N, 0, 1, _ptr(b), int(i < n - 1), None, 0,
                                     0, _stream(X.device)))
        h = acts[i]
    return acts[-1]


def backward(X, Ws, acts, gY, gX, gWs, gbs, scratch, accumulate, side=None):
    """Adjoint of forward(): gY [M, out_last] -> gX [M,K0], gWs[i], gbs[i] (written, or added to when accumulate).
    scratch: two [M, max hidden] buffers for the hidden-layer adjoints.
    side: an optional second stream.  Per layer the three launches (input adjoint, weight gradient, bias gradient) only
    share their input g, and on 1024 x 256 matrices each is a few tens of microseconds of a fraction of the GPU: with
    `side`, the weight / bias gradients leave the critical path (graph B of FusedTrainStep: 9 serial launches -> 3)."""
    lib = _lib.load()
    dev = X.device
    main = torch.cuda.current_stream(dev)
    acc = int(bool(accumulate))
    n = len(Ws)
    g = gY
    freed = {}            # scratch buffer index -> event: the side stream has finished reading it
    g_buf = None
    for i in range(n - 1, -1, -1):
        W = Ws[i]
        inp = acts[i - 1] if i > 0 else X
        M, N, K = g.shape[0], W.shape[0], W.shape[1]
        # dW [N,K] = g^T inp ; db [N] = column sums of g
        if side is not None:
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                st = _stream(dev)
                _lib.check(lib.neat_gemm_f32(_ptr(g), _ptr(inp), _ptr(gWs[i]), N, K, M, N, K, K, 1, 0, None, 0, None, 0, acc, st))
                _lib.check(lib.neat_colsum_f32(_ptr(g), M, N, N, _ptr(gbs[i]), acc, st))
                if g_buf is not None:
                    freed[g_buf] = torch.cuda.Event()
                    freed[g_buf].record(side)
        else:
            st = _stream(dev)
            _lib.check(lib.neat_gemm_f32(_ptr(g), _ptr(inp), _ptr(gWs[i]), N, K, M, N, K, K, 1, 0, None, 0, None, 0, acc, st))
            _lib.check(lib.neat_colsum_f32(_ptr(g), M, N, N, _ptr(gbs[i]), acc, st))
        st = _stream(dev)
        # d inp [M,K] = g W, masked by the ReLU of the layer below (inp > 0) for hidden layers
        if i > 0:
            b = i % 2
            if b in freed:                       # the side stream may still be reading this buffer's previous content
                main.wait_event(freed.pop(b))
            dst = scratch[b][:M * K].view(M, K)
            _lib.check(lib.neat_gemm_f32(_ptr(g), _ptr(W), _ptr(dst), M, K, N, N, K, K, 0, 0, None, 0, _ptr(inp), K, 0, st))
            g, g_buf = dst, b
        else:
            _lib.check(lib.neat_gemm_f32(_ptr(g), _ptr(W), _ptr(gX), M, K, N, N, K, K, 0, 0, None, 0, None, 0, acc, st))
    if side is not None:
        done = torch.cuda.Event()
        done.record(side)
        main.wait_event(done)


class JunctionFFN(torch.autograd.Function):
    """glob = ffn(latents) for the plugin path (autograd hands the gradients to the parameters' .grad)."""

    @staticmethod
    def forward(ctx, latents, *wb):
        Ws, bs = list(wb[0::2]), list(wb[1::2])
        f = lambda t: t.detach().float().contiguous()
        X = f(latents)
        Ws, bs = [f(w) for w in Ws], [f(b) for b in bs]
        acts = [torch.empty(X.shape[0], w.shape[0], device=X.device) for w in Ws]
        forward(X, Ws, bs, acts)
        ctx.save_for_backward(X, *Ws, *acts)
        ctx.n = len(Ws)
        return acts[-1]

    @staticmethod
    def backward(ctx, gY):
        saved = ctx.saved_tensors
        n = ctx.n
        X, Ws, acts = saved[0], list(saved[1:1 + n]), list(saved[1 + n:1 + 2 * n])
        gY = gY.float().contiguous()
        dev = X.device
        gX = torch.empty_like(X)
        gWs = [torch.empty_like(w) for w in Ws]
        gbs = [torch.empty(w.shape[0], device=dev) for w in Ws]
        width = max([X.shape[1]] + [w.shape[1] for w in Ws])
        scratch = [torch.empty(X.shape[0] * width, device=dev) for _ in range(2)]
        backward(X, Ws, acts, gY, gX, gWs, gbs, scratch, False)
        out = [gX]
        for gw, gb in zip(gWs, gbs):
            out += [gw, gb]
        return tuple(out)


def apply_module(ffn, latents):
    """ffn: the model's nn.Sequential; returns ffn(latents) through the kernels above."""
    wb = []
    for m in ffn:
        if isinstance(m, torch.nn.Linear):
            wb += [m.weight, m.bias]
        elif not isinstance(m, torch.nn.ReLU):
            raise _lib.NeatError("junction ffn: only Linear / ReLU layers are supported")
    return JunctionFFN.apply(latents, *wb)
