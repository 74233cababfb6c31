"""TEST INFRASTRUCTURE ONLY -- a chained real-data golden: the UNMODIFIED reference trainer loop body on the ABC toy.

    python oracle/make_golden_chain.py          # build container only (needs /root/reference); ~3 min on 8 cores

Everything below is the reference's own code (imported, not copied): datasets.blender_hawp_dataset.BlenderDataset on
data/abc/00075213 (its only unbuildable dependency, the CUDA extension hawp.base._C, is replaced by oracle/hawp_oracle's
numpy restatement, which is pinned bit-exactly against the reference binary), model.networks.neat_wfr_rend_a.VolSDFNetwork
with the model{} block of confs/abc-neat-a.conf, model.networks.loss_wfr.VolSDFLoss, torch.optim.Adam(lr = 5e-4).  The
loop is code/training/volsdf_train.py:355-374 with the seeds of exp_runner.py:49-51 (torch / numpy 42):

    dataset.change_sampling_idx(1024); for view in 0, 1, 2, 3, 0: item -> collate -> model -> loss -> zero_grad ->
    backward -> step          (the DataLoader's shuffle is replaced by this fixed order; 5 steps)

Stored in tests/golden/chain_abc.npz: the four views' cameras, wireframes and image size; per step the sampled pixel
indices and their ground-truth colours (so that the GPU replay does not need the PNGs), every loss term, beta, the PSNR
and the sampler's iteration count.  tests/test_gpu_chain.py replays it on the B200 through
attraction -> DeviceScene(rng="reference-numpy") -> plugin -> VolSDFLoss -> neat_b200.optim.Adam."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from neat_b200 import synth  # noqa: E402
from oracle import hawp_oracle, ref_shim  # noqa: E402

VIEWS = [0, 1, 2, 3, 0]
R = 1024
N_VIEWS = 4


def load_reference_dataset():
    import cv2
    ref_shim.install()
    # the CUDA extension of the 2D detector -> the numpy restatement
    hawp, base, C = types.ModuleType("hawp"), types.ModuleType("hawp.base"), types.ModuleType("hawp.base._C")

    def encodels(lines, ih, iw, h, w, n):
        m, l, t = hawp_oracle.encodels(lines.cpu().numpy(), ih, iw, h, w, n)
        return torch.from_numpy(m), torch.from_numpy(l), torch.from_numpy(t)

    C.encodels = encodels
    base._C = C
    hawp.base = base
    sys.modules.update({"hawp": hawp, "hawp.base": base, "hawp.base._C": C})
    # rend_util.load_rgb = imageio.imread + skimage.img_as_float32 (stubbed modules): the same values through cv2
    sys.modules["imageio"].imread = lambda p: cv2.cvtColor(cv2.imread(p, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    sys.modules["skimage"].img_as_float32 = lambda a: a.astype(np.float32) / np.float32(255.0)
    # code/datasets/ has no __init__ and collides with the HuggingFace `datasets` package: load it under another name
    pkg = types.ModuleType("refds")
    pkg.__path__ = [os.path.join(ref_shim.REF_CODE, "datasets")]
    sys.modules["refds"] = pkg
    mod = importlib.import_module("refds.blender_hawp_dataset")
    return mod


def main():
    mod = load_reference_dataset()
    conf = synth.abc_conf()
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_CODE)                                    # the dataset opens '../data/<data_dir>'
    try:
        # only the first N_VIEWS images are needed: keep the reference class, shrink its file list
        import utils.general as ug
        glob_all = ug.glob_imgs
        ug.glob_imgs = lambda p: sorted(glob_all(p))[:N_VIEWS]
        ds = mod.BlenderDataset("abc/00075213", [512, 512], reverse_coordinate=True)
        ug.glob_imgs = glob_all
    finally:
        os.chdir(cwd)
    assert ds.n_images == N_VIEWS
    Net, Loss, _, _ = ref_shim.load_classes()
    torch.manual_seed(42)                                          # exp_runner.py:49-51
    np.random.seed(42)
    model = Net(conf=ref_shim.to_config(conf))
    loss_fn = Loss(**synth.loss_conf())
    opt = torch.optim.Adam(model.parameters(), lr=5.0e-4)          # volsdf_train.py:178
    import utils.rend_util as rend_util
    gold = {"views": np.array(VIEWS), "rays": R, "img_res": np.array([512, 512]), "distance_threshold": ds.distance,
            "intrinsics": ds.intrinsics_all.numpy(), "pose": ds.pose_all.numpy()}
    for i, wf in enumerate(ds.wireframes):
        gold["wf%d_vertices" % i] = wf.vertices.numpy()
        gold["wf%d_edges" % i] = wf.edges.numpy()
        gold["wf%d_weights" % i] = wf.weights.numpy()
        gold["lines%d" % i] = ds.lines[i].numpy()
        gold["n_masked%d" % i] = int(ds.masks[i].sum())
    ds.change_sampling_idx(R)                                      # volsdf_train.py:355
    model.train()
    terms = ("loss", "rgb_loss", "eikonal_loss", "line_loss", "l2d_loss", "j3d_loss", "j2d_loss", "count", "jcount")
    traj = {k: [] for k in terms + ("beta", "psnr", "median")}
    pix, rgb_gt = [], []
    for step, view in enumerate(VIEWS):
        indices, model_input, ground_truth = ds.collate_fn([ds[view]])
        uv = model_input["uv"][0]
        pix.append((uv[:, 1] * 512 + uv[:, 0]).long().numpy().astype(np.int32))   # uv = (column, row)
        rgb_gt.append(ground_truth["rgb"][0].numpy())
        out = model(model_input)                                   # :366
        lo = loss_fn(out, ground_truth)                            # :367
        opt.zero_grad()
        lo["loss"].backward()
        opt.step()                                                 # :372-374
        psnr = rend_util.get_psnr(out["rgb_values"], ground_truth["rgb"].reshape(-1, 3))
        for k in terms:
            traj[k].append(float(lo[k]))
        traj["beta"].append(float(model.density.get_beta()))
        traj["psnr"].append(float(psnr))
        traj["median"].append(float(out["median"]) if "median" in out else 0.0)
        print("step", step, "view", view, {k: round(traj[k][-1], 6) for k in ("loss", "rgb_loss", "line_loss", "j3d_loss", "beta", "psnr")},
              flush=True)
    for k, v in traj.items():
        gold["traj_" + k] = np.array(v, dtype=np.float64)
    gold["pixels"] = np.stack(pix)
    gold["rgb_gt"] = np.stack(rgb_gt).astype(np.float32)
    out = os.path.join(ROOT, "tests", "golden", "chain_abc.npz")
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
