"""Drop-in mirror of the reference's loss plugin (`train.loss_class`):

    model.networks.loss_wfr.VolSDFLoss   ->   neat_b200.loss.VolSDFLoss

(code/model/networks/loss_wfr.py:16-139).  The loss works on [R]-sized per-ray tensors that are already on
the device; the gradients it sends back (d rgb_values, d lines2d_calib -> d lines3d, d grad_theta) enter the
hand-written backward through neat_b200.autograd.NeatStepFunction."""
import importlib

import torch
from torch import nn


def _get_class(path):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


class VolSDFLoss(nn.Module):
    def __init__(self, rgb_loss, eikonal_weight, line_weight, junction_3d_weight=0.1, junction_2d_weight=0.01):
        super().__init__()
        self.eikonal_weight, self.line_weight = eikonal_weight, line_weight
        self.rgb_loss = _get_class(rgb_loss)(reduction="mean")
        self.steps = 0
        self.junction_3d_weight, self.junction_2d_weight = junction_3d_weight, junction_2d_weight

    def get_rgb_loss(self, rgb_values, rgb_gt):
        return self.rgb_loss(rgb_values, rgb_gt.reshape(-1, 3))

    def get_eikonal_loss(self, grad_theta):
        return ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()

    def get_line_loss(self, lines2d, lines2d_gt, lines_weight, threshold=100):
        swapped = lines2d_gt[:, [2, 3, 0, 1]]
        d1 = ((lines2d - lines2d_gt) ** 2).sum(-1, keepdim=True).detach()
        d2 = ((lines2d - swapped) ** 2).sum(-1, keepdim=True).detach()
        per = (lines2d - torch.where(d1 < d2, lines2d_gt, swapped)).abs().mean(dim=-1)
        labels = (per.detach() < threshold).long()
        return (per * lines_weight.flatten() * labels).sum() / labels.sum().clamp_min(1), per.detach()

    def forward(self, model_outputs, ground_truth):
        from scipy.optimize import linear_sum_assignment
        self.steps += 1
        dev = model_outputs["rgb_values"].device
        lines2d_gt, lines_weight = ground_truth["lines2d"][0].to(dev).split(4, dim=-1)
        if "labels" in ground_truth:
            lines_weight = lines_weight * ground_truth["labels"][0, :, None].to(dev)
        l2d_uncal, thr = self.get_line_loss(model_outputs["lines2d"].reshape(-1, 4), lines2d_gt, lines_weight)
        count = (thr < 100).sum()
        g = lines2d_gt.reshape(-1, 2)
        gh = torch.cat([g, torch.ones_like(g[:, :1])], dim=-1)
        gh = (model_outputs["K"].inverse() @ gh.t()).t()
        gcal = (gh[:, :2] / gh[:, 2, None]).reshape(-1, 4)
        line_loss, _ = self.get_line_loss(model_outputs["lines2d_calib"].reshape(-1, 4), gcal,
                                          lines_weight * (thr < 100).reshape(-1, 1))
        if torch.isnan(line_loss):
            raise FloatingPointError("line loss is NaN")  # the reference drops into pdb here (loss_wfr.py:66-67)
        rgb_loss = self.get_rgb_loss(model_outputs["rgb_values"], ground_truth["rgb"].to(dev))
        if "grad_theta" in model_outputs:
            eik = self.get_eikonal_loss(model_outputs["grad_theta"])
        else:
            eik = torch.tensor(0.0, device=dev)
        loss = rgb_loss + self.eikonal_weight * eik + self.line_weight * line_loss
        zero = torch.tensor(0.0, device=dev)
        out = {"rgb_loss": rgb_loss, "eikonal_loss": eik, "line_loss": line_loss, "l2d_loss": l2d_uncal, "count": count,
               "j3d_loss": zero, "j2d_loss": zero, "j2d_stat": zero, "jcount": zero}
        if "j3d_local" in model_outputs and model_outputs["j3d_local"].shape[0] > 0:
            j3l, j3g = model_outputs["j3d_local"], model_outputs["j3d_global"]
            j2l, j2g = model_outputs["j2d_local"].detach(), model_outputs["j2d_global"].detach()
            j2lc, j2gc = model_outputs["j2d_local_calib"], model_outputs["j2d_global_calib"]
            with torch.no_grad():
                cost = torch.cdist(j3l, j3g, p=1) + 0.1 * torch.cdist(j2lc, j2gc, p=1)
            a0, a1 = linear_sum_assignment(cost.detach().cpu().numpy())
            l3 = (j3l[a0] - j3g[a1]).abs().sum(-1).mean()
            l2 = (j2lc[a0] - j2gc[a1]).abs().sum(-1).mean()
            with torch.no_grad():
                l2u = (j2l[a0] - j2g[a1]).abs().sum(-1).mean()
            loss = loss + self.junction_3d_weight * l3 + self.junction_2d_weight * l2
            out.update(j3d_loss=l3, j2d_loss=l2, j2d_stat=l2u, jcount=(cost[a0, a1] < 10).sum())
        out["loss"] = loss
        if "median" in model_outputs:
            out["median"] = model_outputs["median"]
        return out
