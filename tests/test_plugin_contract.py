"""CPU: the drop-in plugin keeps the reference's constructor / state_dict contract and has NO CPU fallback."""
import numpy as np
import pytest
import torch

import golden_io as G
from neat_b200 import _lib, synth
from neat_b200.loss import VolSDFLoss
from neat_b200.model import VolSDFNetwork


@pytest.mark.parametrize("name", list(G.CASES))
def test_state_dict_keys_match_reference(name):
    """synth.make_state_dict reproduces the key set / shapes of the reference module (the goldens were produced by
    loading it into the unmodified reference with strict=True); ours must load it strictly too."""
    g, conf, sd_np = G.load(name)
    model = VolSDFNetwork(conf)
    sd = model.state_dict()
    assert set(sd.keys()) == set(sd_np.keys())
    for k, v in sd_np.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd_np.items()}, strict=True)
    # parameter (not just buffer) names, as the optimizer / checkpoints see them
    names = {n for n, _ in model.named_parameters()}
    assert "density.beta" in names and "latents" in names and "implicit_network.lin0.weight_g" in names


def test_geometric_init_statistics():
    """ImplicitNetwork geometric init (neat_wfr_rend_a.py:55-69): sdf ~ |x| - bias right after construction."""
    torch.manual_seed(0)
    model = VolSDFNetwork(synth.dtu_conf())
    net = model.implicit_network
    assert float(net.lin8.bias[0]) == pytest.approx(-0.6)
    assert float(net.lin0.weight_v[:, 3:].abs().max()) == 0.0            # PE columns start at zero
    assert float(net.lin4.weight_v[:, -36:].abs().max()) == 0.0          # and so does the skip's PE part
    w = net.lin8.weight_v[0]
    assert float(w.mean()) == pytest.approx(np.sqrt(np.pi) / np.sqrt(256), rel=1e-2)


def test_no_cpu_fallback():
    model = VolSDFNetwork(synth.toy_conf())
    inp = {"intrinsics": torch.eye(4)[None], "pose": torch.eye(4)[None], "uv": torch.zeros(1, 4, 2),
           "uv_proj": torch.zeros(1, 4, 2), "wireframe": []}
    with pytest.raises(_lib.NeatError):
        model(inp)
    with pytest.raises(_lib.NeatError):
        model.implicit_network.get_sdf_vals(torch.zeros(4, 3))


def test_unsupported_confs_are_rejected():
    c = synth.toy_conf()
    c["white_bkgd"] = True
    with pytest.raises(_lib.NeatError):
        VolSDFNetwork(c)
    c = synth.toy_conf()
    c["rendering_network"]["mode"] = "nerf"
    with pytest.raises(_lib.NeatError):
        VolSDFNetwork(c)


def test_loss_cpu_path_matches_oracle():
    """VolSDFLoss on CPU tensors (the pure-torch mirror of loss_wfr.py used when outputs are not on CUDA) equals
    the oracle's restatement, which is pinned against the reference."""
    from oracle import neat_oracle as O
    rs = np.random.RandomState(0)
    R = 64
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    out = {"rgb_values": T(rs.uniform(size=(R, 3))), "lines2d": T(rs.uniform(0, 500, size=(R, 2, 2))),
           "lines2d_calib": T(rs.normal(size=(R, 2, 2))), "grad_theta": T(rs.normal(size=(2 * R, 3))),
           "K": T([[560, 0, 256], [0, 560, 256], [0, 0, 1]]), "j3d_local": torch.zeros(0, 3)}
    gt = {"rgb": T(rs.uniform(size=(1, R, 3))), "lines2d": T(np.concatenate([rs.uniform(0, 500, size=(1, R, 4)),
                                                                             rs.uniform(0.3, 1, size=(1, R, 1))], -1))}
    ours = VolSDFLoss(**synth.loss_conf())(out, gt)
    ref = O.neat_loss(out, gt["rgb"][0], gt["lines2d"][0], out["K"])
    for k in ("loss", "rgb_loss", "eikonal_loss", "line_loss", "l2d_loss"):
        assert float(ours[k]) == pytest.approx(float(ref[k]), rel=1e-6), k
    assert int(ours["count"]) == int(ref["count"])
