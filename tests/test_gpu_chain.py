"""One chained real-data check (VERDICT r01 missing #6): the ABC toy of the reference tree (data/abc/00075213, views 0-3)
through attraction -> DeviceScene(rng="reference-numpy") -> plugin -> VolSDFLoss -> neat_b200.optim.Adam for five optimizer
steps, against the trajectory the UNMODIFIED reference (BlenderDataset + VolSDFNetwork + VolSDFLoss + torch.optim.Adam,
oracle/make_golden_chain.py) produced from the same seeds: same pixels drawn, every loss term, beta and the PSNR per step.
Reference: code/datasets/blender_hawp_dataset.py:17-229, code/training/volsdf_train.py:297-408."""
import os

import numpy as np
import pytest
import torch

from neat_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_abc.npz")


def test_five_training_steps_on_the_abc_toy_follow_the_reference():
    from neat_b200.dataset import DeviceScene
    from neat_b200.loss import VolSDFLoss
    from neat_b200.model import VolSDFNetwork
    from neat_b200.optim import Adam
    from neat_b200.trainer import Wireframe
    g = dict(np.load(GOLD))
    dev = torch.device("cuda:0")
    H, W = (int(x) for x in g["img_res"])
    R, views = int(g["rays"]), [int(v) for v in g["views"]]
    T = lambda a: torch.from_numpy(np.asarray(a))
    scene = DeviceScene((H, W), device=dev, rng="reference-numpy")
    for i in range(4):
        # the images are not shipped: only the colours of the pixels the trajectory samples are needed
        rgb = torch.zeros(H * W, 3)
        for s, v in enumerate(views):
            if v == i:
                rgb[T(g["pixels"][s]).long()] = T(g["rgb_gt"][s])
        wf = Wireframe(g["wf%d_vertices" % i], g["wf%d_edges" % i], g["wf%d_weights" % i])
        scene.add_image(rgb, T(g["lines%d" % i]), T(g["intrinsics"][i]), T(g["pose"][i]), wireframe=wf,
                        distance_threshold=float(g["distance_threshold"]))
        assert int(scene.images[i].masked.numel()) == int(g["n_masked%d" % i])   # the attraction support region
    torch.manual_seed(42)                                     # exp_runner.py:49-51
    np.random.seed(42)
    model = VolSDFNetwork(synth.abc_conf()).to(dev).train()   # same seed => the reference's initial weights
    model.rng = "reference"                                   # replay the reference's CPU-generator draws
    loss_fn = VolSDFLoss(**synth.loss_conf())
    opt = Adam(model.parameters(), lr=5.0e-4)
    scene.change_sampling_idx(R)
    for s, v in enumerate(views):
        idx, sample, gt = scene.collate_fn([scene[v]])
        assert np.array_equal(sample["sampling_idx"][0].cpu().numpy().astype(np.int32), g["pixels"][s]), "pixels of step %d" % s
        out = model(sample)
        lo = loss_fn(out, gt)
        opt.zero_grad()
        lo["loss"].backward()
        opt.step()
        mse = torch.mean((out["rgb_values"].detach() - gt["rgb"].reshape(-1, 3)) ** 2)
        psnr = float(-10.0 * torch.log(mse) / np.log(10.0))
        for k in ("loss", "rgb_loss", "eikonal_loss", "line_loss", "l2d_loss", "j3d_loss", "j2d_loss"):
            ref = float(g["traj_" + k][s])
            assert abs(float(lo[k]) - ref) <= 1e-3 * max(1.0, abs(ref)), (s, k, float(lo[k]), ref)
        assert int(lo["count"]) == int(g["traj_count"][s])
        assert abs(float(model.density.get_beta()) - float(g["traj_beta"][s])) <= 1e-5
        assert abs(psnr - float(g["traj_psnr"][s])) <= 1e-2
