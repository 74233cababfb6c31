"""Drop-in mirror of the reference's loss plugin (`train.loss_class`):

    model.networks.loss_wfr.VolSDFLoss   ->   neat_b200.loss.VolSDFLoss

(code/model/networks/loss_wfr.py:16-139).  The loss works on [R]-sized per-ray tensors that are already on
the device; the gradients it sends back (d rgb_values, d lines2d_calib -> d lines3d, d grad_theta) enter the
hand-written backward through neat_b200.autograd.NeatStepFunction."""
import ctypes
import importlib

import torch
from torch import nn

from . import _lib

_P = ctypes.c_void_p


def _ptr(t):
    return _P(t.data_ptr()) if t is not None else None


class _FusedCoreLoss(torch.autograd.Function):
    """rgb L1 + eikonal + the two 2D line losses in three small kernels that also produce the gradients
    (neat_b200/csrc/loss.cuh).  Returns (loss_core, stats[8]); stats carries no gradient."""

    @staticmethod
    def forward(ctx, rgb_values, lines2d_calib, grad_theta, lines2d, rgb_gt, lines_gt, labels, K3, eik_w, line_w):
        lib = _lib.load()
        dev = rgb_values.device
        f = lambda t: None if t is None else t.detach().to(dev, torch.float32).contiguous()
        rgb_values, lines2d_calib, grad_theta, lines2d = f(rgb_values), f(lines2d_calib), f(grad_theta), f(lines2d)
        rgb_gt, lines_gt, labels, K3 = f(rgb_gt), f(lines_gt), f(labels), f(K3)
        R = rgb_values.shape[0]
        n_eik = 0 if grad_theta is None else grad_theta.shape[0]
        scratch = torch.empty(8 + R, device=dev)
        out = torch.zeros(8, device=dev)
        g_rgb = torch.empty(R, 3, device=dev)
        g_calib = torch.empty(R, 4, device=dev)
        g_theta = torch.empty(max(n_eik, 1), 3, device=dev)
        a = _lib.LossArgs(R, n_eik, _ptr(rgb_values), _ptr(rgb_gt), _ptr(lines2d), _ptr(lines2d_calib), _ptr(lines_gt),
                          _ptr(labels), _ptr(K3), 3, _ptr(grad_theta), float(eik_w), float(line_w), _ptr(scratch), _ptr(out),
                          _ptr(g_rgb), _ptr(g_calib), _ptr(g_theta))
        _lib.check(lib.neat_loss_forward_backward(ctypes.byref(a), _P(torch.cuda.current_stream(dev).cuda_stream)))
        ctx.save_for_backward(g_rgb, g_calib, g_theta)
        ctx.has_theta = n_eik > 0
        ctx.shapes = (lines2d_calib.shape,)
        stats = out.clone()
        ctx.mark_non_differentiable(stats)
        return out[0], stats

    @staticmethod
    def backward(ctx, g, _g_stats):
        g_rgb, g_calib, g_theta = ctx.saved_tensors
        return (g * g_rgb, (g * g_calib).view(-1, 2, 2), g * g_theta if ctx.has_theta else None,
                None, None, None, None, None, None, None)


def _get_class(path):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


class _JunctionTerms(torch.autograd.Function):
    """The Hungarian-matched junction terms (loss_wfr.py:110-121) as one kernel each way (csrc/junction.cuh):
    returns out[3] = (j3d_loss, j2d_loss, j2d_stat); gradients flow to the GLOBAL junctions and their calibrated
    projections only (the local junctions are detached DBSCAN centroids, in the reference too)."""

    @staticmethod
    def forward(ctx, j3g, j2gc, j3l, j2lc, j2l, j2g, rows, cols):
        lib = _lib.load()
        dev = j3g.device
        f = lambda t: t.detach().to(dev, torch.float32).contiguous()
        j3g_, j2gc_, j3l, j2lc, j2l, j2g = f(j3g), f(j2gc), f(j3l), f(j2lc), f(j2l), f(j2g)
        rows = rows.to(dev, torch.int32).contiguous()
        cols = cols.to(dev, torch.int32).contiguous()
        n = rows.shape[0]
        out = torch.empty(3, device=dev)
        _lib.check(lib.neat_junction_terms(n, _ptr(j3l), _ptr(j3g_), _ptr(j2lc), _ptr(j2gc_), _ptr(j2l), _ptr(j2g),
                                           _ptr(rows), _ptr(cols), _ptr(out), _P(torch.cuda.current_stream(dev).cuda_stream)))
        ctx.save_for_backward(j3l, j3g_, j2lc, j2gc_, rows, cols)
        return out

    @staticmethod
    def backward(ctx, g_out):
        j3l, j3g, j2lc, j2gc, rows, cols = ctx.saved_tensors
        lib = _lib.load()
        dev = j3g.device
        G = j3g.shape[0]
        g3 = torch.empty(G, 3, device=dev)
        g2 = torch.empty(G, 2, device=dev)
        g_out = g_out.to(dev, torch.float32).contiguous()
        _lib.check(lib.neat_junction_terms_backward(rows.shape[0], G, _ptr(j3l), _ptr(j3g), _ptr(j2lc), _ptr(j2gc),
                                                    _ptr(rows), _ptr(cols), _ptr(g_out), _ptr(g3), _ptr(g2),
                                                    _P(torch.cuda.current_stream(dev).cuda_stream)))
        return g3, g2, None, None, None, None, None, None


class VolSDFLoss(nn.Module):
    def __init__(self, rgb_loss, eikonal_weight, line_weight, junction_3d_weight=0.1, junction_2d_weight=0.01):
        super().__init__()
        self.eikonal_weight, self.line_weight = eikonal_weight, line_weight
        self.rgb_loss = _get_class(rgb_loss)(reduction="mean")
        self.steps = 0
        self.junction_3d_weight, self.junction_2d_weight = junction_3d_weight, junction_2d_weight

    def forward(self, model_outputs, ground_truth):
        self.steps += 1
        dev = model_outputs["rgb_values"].device
        if dev.type != "cuda":
            raise _lib.NeatError("neat_b200.loss.VolSDFLoss runs on CUDA tensors only: there is no CPU or eager-PyTorch "
                                 "path (the CPU restatement of loss_wfr.py lives in oracle/, for tests)")
        return self._forward_fused(model_outputs, ground_truth, dev)

    def _forward_fused(self, mo, gt, dev):
        """Same contract as forward(); the per-ray terms and their gradients come from the CUDA kernels."""
        if not isinstance(self.rgb_loss, nn.L1Loss):
            raise _lib.NeatError("the fused loss implements rgb_loss = torch.nn.L1Loss (the shipped confs)")
        labels = gt["labels"][0] if "labels" in gt else None
        core, stats = _FusedCoreLoss.apply(mo["rgb_values"], mo["lines2d_calib"].reshape(-1, 2, 2), mo.get("grad_theta"),
                                           mo["lines2d"].reshape(-1, 4), gt["rgb"].reshape(-1, 3), gt["lines2d"][0], labels,
                                           mo["K"], self.eikonal_weight, self.line_weight)
        zero = torch.zeros((), device=dev)
        out = {"rgb_loss": stats[1], "eikonal_loss": stats[2], "line_loss": stats[3], "l2d_loss": stats[4],
               "count": stats[5].long(), "j3d_loss": zero, "j2d_loss": zero, "j2d_stat": zero, "jcount": zero}
        loss = core
        if "j3d_local" in mo and mo["j3d_local"].shape[0] > 0:
            j3l, j3g = mo["j3d_local"], mo["j3d_global"]
            j2l, j2g = mo["j2d_local"].detach(), mo["j2d_global"].detach()
            j2lc, j2gc = mo["j2d_local_calib"], mo["j2d_global_calib"]
            if "_junction_assignment" in mo:  # computed by neat_b200.model in its single host round trip
                a0, a1, jcount = mo["_junction_assignment"]
                jcount = torch.tensor(jcount, device=dev)
            else:  # outputs of another model (e.g. the reference's): the assignment is solved here, natively on the host
                from . import junction
                with torch.no_grad():
                    cost = torch.cdist(j3l, j3g, p=1) + 0.1 * torch.cdist(j2lc, j2gc, p=1)
                r, c = junction.linear_sum_assignment(cost.cpu().numpy())
                a0 = torch.as_tensor(r, device=dev)
                a1 = torch.as_tensor(c, device=dev)
                jcount = (cost[a0, a1] < 10).sum()
            terms = _JunctionTerms.apply(j3g, j2gc, j3l, j2lc, j2l, j2g, a0, a1)
            l3, l2, l2u = terms[0], terms[1], terms[2].detach()
            loss = loss + self.junction_3d_weight * l3 + self.junction_2d_weight * l2
            out.update(j3d_loss=l3, j2d_loss=l2, j2d_stat=l2u, jcount=jcount)
        out["loss"] = loss
        if "median" in mo:
            out["median"] = mo["median"]
        return out
