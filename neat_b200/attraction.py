"""Dataset-side attraction precompute on the GPU: drop-ins for `hawp.base._C.encodels`
(third-party/hawp/hawp/base/csrc/binding.cpp:3-5) and `SceneDataset.compute_point_line_attraction`
(code/datasets/scene_hawp_dataset.py:92-146)."""
import ctypes

import torch

from . import _lib

_P = ctypes.c_void_p


def _stream(dev):
    return _P(torch.cuda.current_stream(dev).cuda_stream)


def encodels(lines, input_height, input_width, height, width, num_lines):
    """Same signature and return value as hawp.base._C.encodels: (map [6,H,W], label [n,H,W] bool, tmap [1,H,W])."""
    if not lines.is_cuda:
        raise _lib.NeatError("encodels: `lines` must be a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    lines = lines.detach().float().contiguous()
    dev = lines.device
    mp = torch.empty(6, height, width, device=dev)
    label = torch.empty(num_lines, height, width, dtype=torch.bool, device=dev)
    tmap = torch.empty(1, height, width, device=dev)
    _lib.check(lib.neat_encodels(_P(lines.data_ptr()), input_height, input_width, height, width, num_lines,
                                 _P(mp.data_ptr()), _P(label.data_ptr()), _P(tmap.data_ptr()), _stream(dev)))
    return mp, label, tmap


def compute_point_line_attraction(lines, img_res, distance):
    """lines [n, >=4] (x1,y1,x2,y2,...) -> (mask [HW] bool, labels [HW] int64, proj_points [HW,2]) on lines.device;
    the reference returns mask / labels on the CPU and proj_points on the GPU -- move them as needed."""
    if not lines.is_cuda:
        raise _lib.NeatError("compute_point_line_attraction: `lines` must be a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    l4 = lines.detach()[:, :4].float().contiguous()
    dev = l4.device
    H, W = int(img_res[0]), int(img_res[1])
    mask = torch.empty(H * W, dtype=torch.bool, device=dev)
    labels = torch.empty(H * W, dtype=torch.int64, device=dev)
    proj = torch.empty(H * W, 2, device=dev)
    _lib.check(lib.neat_point_line_attraction(_P(l4.data_ptr()), l4.shape[0], H, W, ctypes.c_float(distance),
                                              _P(mask.data_ptr()), _P(labels.data_ptr()), _P(proj.data_ptr()), _stream(dev)))
    return mask, labels, proj
