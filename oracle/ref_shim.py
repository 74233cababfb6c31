"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (cherubicXN/neat).

Loads ``model.networks.neat_wfr_rend_a.VolSDFNetwork`` and
``model.networks.loss_wfr.VolSDFLoss`` straight from ``/root/reference/code`` so that
golden vectors can be generated from the reference itself (``oracle/make_golden.py``)
and so that the CPU restatement in ``oracle/neat_oracle.py`` can be pinned against it.

It exists only in the build container: ``/root/reference`` is absent on the GPU box, so
nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may import this file.

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


def available() -> bool:
    return os.path.isdir(os.path.join(REF_CODE, "model", "networks"))


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
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    _installed = True


def load_classes():
    """Return (VolSDFNetwork, VolSDFLoss, module_net, module_loss) from the reference."""
    install()
    import importlib

    net = importlib.import_module("model.networks.neat_wfr_rend_a")
    loss = importlib.import_module("model.networks.loss_wfr")
    assert net.__file__.startswith(REF_CODE), net.__file__
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
