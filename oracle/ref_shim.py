"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (cherubicXN/neat).

Loads ``model.networks.neat_wfr_rend_a.VolSDFNetwork`` and
``model.networks.loss_wfr.VolSDFLoss`` straight from ``/root/reference/code`` so that
golden vectors can be generated from the reference itself (``oracle/make_golden.py``)
and so that the CPU restatement in ``oracle/neat_oracle.py`` can be pinned against it.

``/root/reference`` is absent on the GPU box; there the same seven files are imported from
``oracle/_ref/neat_ref_code.zip`` (built, unmodified, by ``oracle/build_ref.py``).  Only ``bench.py``'s
reference legs (``--impl reference``, ``cpu_baseline``, ``gpu_eager_baseline``) and tests may import this file --
never the product path.

What the shim does:
  * stubs the third-party imports the hot path never calls
    (open3d, trimesh, imageio, skimage, matplotlib, pyhocon, GPUtil, plotly);
  * provides a dict-backed ``ConfigTree`` with the six getters the reference uses
    (``code/model/networks/neat_wfr_rend_a.py:258-315``);
  * makes ``Tensor.cuda`` / ``Module.cuda`` the identity when no GPU is present
    (the reference hard-codes ``.cuda()``, e.g. ``code/model/ray_sampler.py:71``).
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("NEAT_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")
# the same seven files, archived unmodified by oracle/build_ref.py (travels to the GPU box, where REF_ROOT is absent)
REF_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "neat_ref_code.zip")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_CODE, "model", "networks")) or os.path.exists(REF_ZIP)


def code_location() -> str:
    """Where `import model...` resolves: the reference tree when present, else the archive built from it."""
    return REF_CODE if os.path.isdir(os.path.join(REF_CODE, "model", "networks")) else REF_ZIP


class ConfigTree(dict):
    """Minimal stand-in for pyhocon.ConfigTree (getters with defaults)."""

    _MISSING = object()

    def _get(self, key, default=_MISSING):
        if key in self:
            return self[key]
        if default is ConfigTree._MISSING:
            raise KeyError(key)
        return default

    def get_int(self, key, default=_MISSING):
        return int(self._get(key, default))

    def get_float(self, key, default=_MISSING):
        return float(self._get(key, default))

    def get_bool(self, key, default=_MISSING):
        return bool(self._get(key, default))

    def get_string(self, key, default=_MISSING):
        return str(self._get(key, default))

    def get_list(self, key, default=_MISSING):
        return list(self._get(key, default))

    def get_config(self, key, default=_MISSING):
        v = self._get(key, default)
        return v if isinstance(v, ConfigTree) else ConfigTree(v)


def to_config(d):
    out = ConfigTree()
    for k, v in d.items():
        out[k] = to_config(v) if isinstance(v, dict) else v
    return out


_installed = False


def install():
    """Install stubs + sys.path so ``import model.networks...`` resolves to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("open3d")
    stub("trimesh")
    stub("imageio")
    stub("skimage")
    stub("GPUtil")
    stub("plotly")
    mpl = stub("matplotlib")
    plt = stub("matplotlib.pyplot")
    mpl.pyplot = plt
    stub("pyhocon", ConfigTree=ConfigTree, ConfigFactory=object)

    if not torch.cuda.is_available():
        force_cpu(True)

    loc = code_location()
    if loc not in sys.path:
        sys.path.insert(0, loc)
    _installed = True


_orig_cuda = (torch.Tensor.cuda, torch.nn.Module.cuda)


def force_cpu(on: bool):
    """The reference hard-codes `.cuda()` (e.g. code/model/ray_sampler.py:71).  on=True makes Tensor.cuda / Module.cuda
    the identity so that the unmodified classes run on the host CPU (always the case without a GPU; on the GPU box this is
    how bench.py times the reference's CPU path); on=False restores torch's own methods (the reference on the B200)."""
    if on:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    else:
        torch.Tensor.cuda, torch.nn.Module.cuda = _orig_cuda


def load_classes():
    """Return (VolSDFNetwork, VolSDFLoss, module_net, module_loss) from the reference."""
    install()
    import importlib

    net = importlib.import_module("model.networks.neat_wfr_rend_a")
    loss = importlib.import_module("model.networks.loss_wfr")
    assert net.__file__.startswith(code_location()), net.__file__
    return net.VolSDFNetwork, loss.VolSDFLoss, net, loss


class Wireframe:
    """Host-side container with the two members the model touches
    (``code/utils/hawp_util.py:7-95``): ``vertices`` [J,2] and ``line_segments()`` [E,5]."""

    def __init__(self, vertices, edges, weights=None):
        self.vertices = torch.as_tensor(vertices, dtype=torch.float32)
        self.edges = torch.as_tensor(edges, dtype=torch.long)
        self.weights = (torch.ones(len(self.edges)) if weights is None
                        else torch.as_tensor(weights, dtype=torch.float32))

    def line_segments(self, threshold=0.05):
        keep = self.weights > threshold
        e = self.edges[keep]
        return torch.cat([self.vertices[e[:, 0]], self.vertices[e[:, 1]],
                          self.weights[keep][:, None]], dim=-1)
