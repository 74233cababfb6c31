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
    c["ray_sampler"]["inverse_sphere_bg"] = True
    from neat_b200.context import sampler_config_from_conf
    with pytest.raises(_lib.NeatError):        # cannot run in the reference's own class either
        sampler_config_from_conf(c)
    VolSDFNetwork(synth.toy_white_conf())     # white_bkgd + junction_eikonal: supported (golden case toy_white_jeik)
    VolSDFNetwork(synth.toy_l3d_conf())       # use_l3d without DBSCAN: supported (golden case toy_l3d)
    c = synth.toy_conf()
    c["rendering_network"]["mode"] = "nerf"
    with pytest.raises(_lib.NeatError):
        VolSDFNetwork(c)


def test_loss_has_no_cpu_path():
    """VolSDFLoss is a CUDA-kernel path only: CPU tensors raise instead of falling back to eager PyTorch (the CPU
    restatement of loss_wfr.py is oracle.neat_oracle.neat_loss, pinned in test_oracle_golden.py)."""
    R = 8
    out = {"rgb_values": torch.zeros(R, 3), "lines2d": torch.zeros(R, 2, 2), "lines2d_calib": torch.zeros(R, 2, 2),
           "grad_theta": torch.ones(2 * R, 3), "K": torch.eye(3), "j3d_local": torch.zeros(0, 3)}
    gt = {"rgb": torch.zeros(1, R, 3), "lines2d": torch.zeros(1, R, 5)}
    with pytest.raises(_lib.NeatError):
        VolSDFLoss(**synth.loss_conf())(out, gt)


def test_same_seed_gives_the_reference_initial_weights():
    """Constructing the plugin under torch.manual_seed(s) draws the same random numbers in the same order as the
    reference constructors (neat_wfr_rend_a.py:46-72, 257-303): bit-identical initial state dicts, every shipped model
    block.  Needs the reference (tree or the archive oracle/build_ref.py makes)."""
    import pytest
    import torch
    from oracle import ref_shim
    from neat_b200 import synth
    from neat_b200.model import VolSDFNetwork
    if not ref_shim.available():
        pytest.skip("reference not available")
    Net, _, _, _ = ref_shim.load_classes()
    for conf in (synth.dtu_conf(), synth.abc_conf(), synth.toy_conf()):
        torch.manual_seed(42)
        ref = Net(conf=ref_shim.to_config(conf)).state_dict()
        torch.manual_seed(42)
        mine = VolSDFNetwork(conf).state_dict()
        assert set(ref) == set(mine)
        for k in ref:
            assert torch.equal(ref[k], mine[k]), k
