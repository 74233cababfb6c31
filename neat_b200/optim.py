"""torch.optim.Adam as the reference trainer uses it (code/training/volsdf_train.py:178: Adam(lr), stepped at :374, lr
driven by ExponentialLR :180-182) with the update of ALL parameter tensors in one kernel launch (csrc/adam.cuh).
It is a torch.optim.Optimizer, so lr schedulers, param groups and state_dict()/load_state_dict() work as with torch's;
the state layout (step, exp_avg, exp_avg_sq per parameter) is torch.optim.Adam's, checkpoints are interchangeable."""
import ctypes

import torch

from . import _lib

_P = ctypes.c_void_p


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = float(grad_scale)  # e.g. 1 / world_size after an all-reduce(SUM) of the gradients
        self._tables = {}

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in plist)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1]
        arr = (_lib.AdamTensor * len(plist))()
        for i, p in enumerate(plist):
            stt = self.state[p]
            arr[i] = _lib.AdamTensor(_P(p.data_ptr()), _P(p.grad.data_ptr()), _P(stt["exp_avg"].data_ptr()),
                                     _P(stt["exp_avg_sq"].data_ptr()), p.numel())
        self._tables[gi] = (key, arr)
        return arr

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            dev = plist[0].device
            step = None
            for p in plist:
                if p.device != dev or dev.type != "cuda":
                    raise _lib.NeatError("neat_b200.optim.Adam: all parameters of a group must live on one CUDA device")
                if p.dtype != torch.float32 or not p.is_contiguous() or p.grad.dtype != torch.float32 or \
                        not p.grad.is_contiguous() or p.grad.is_sparse:
                    raise _lib.NeatError("neat_b200.optim.Adam: contiguous dense fp32 parameters and gradients only")
                stt = self.state[p]
                if not stt:
                    stt["step"] = torch.zeros((), dtype=torch.float32)  # host scalar, like torch (capturable=False)
                    stt["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    stt["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                stt["step"] += 1
                s = int(stt["step"])
                if step is None:
                    step = s
                elif s != step:
                    raise _lib.NeatError("neat_b200.optim.Adam: parameters of one group must share the step count")
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                for i in range(0, len(plist), 128):
                    chunk = plist[i:i + 128]
                    arr = self._table((gi, i), chunk)
                    _lib.check(lib.neat_adam_step(arr, len(chunk), float(group["lr"]), float(b1), float(b2),
                                                  float(group["eps"]), float(group["weight_decay"]), step,
                                                  self.grad_scale, _P(torch.cuda.current_stream(dev).cuda_stream)))
            # the kernel wrote through raw pointers: tell autograd / version-keyed caches (VolSDFNetwork._sync_weights
            # decides from p._version whether the packed tcgen05 weight slabs are stale) that the parameters changed
            for p in plist:
                torch.autograd.graph.increment_version(p)
        return loss
