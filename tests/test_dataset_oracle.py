"""Per-step pixel sampling.  CPU: the numpy restatement of SceneDataset.__getitem__ vs goldens
from the unmodified reference method (oracle/make_golden_dataset.py), and the host evaluation of the device draw's
bijection vs its numpy mirror.  GPU: the gather / compaction kernels vs the oracle (bit-exact: index work and copies)."""
import os

import numpy as np
import pytest
import torch

from oracle import dataset_oracle as DO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tables(dist=5.0):
    g = np.load(os.path.join(GOLDEN, "hawp_abc.npz"))
    H, W = (int(v) for v in g["img_res"])
    mask = np.unpackbits(g["mask_%g" % dist])[:H * W].astype(bool)
    labels = g["labels_%g" % dist].astype(np.int64)
    att = np.zeros((H * W, 2), dtype=np.float32)
    att[g["proj_idx_%g" % dist]] = g["proj_val_%g" % dist]
    rgb = np.random.default_rng(7).random((H * W, 3), dtype=np.float32)      # as in make_golden_dataset.py
    return (H, W), g["lines"].astype(np.float32), mask, labels, att, rgb


def test_getitem_oracle_vs_reference():
    d = np.load(os.path.join(GOLDEN, "dataset_abc.npz"))
    (H, W), lines, mask, labels, att, rgb = _tables()
    R = int(d["R"])
    s, gt = DO.getitem((H, W), rgb, lines, mask, labels, att, perm=d["perm"], num_pixels=R)
    for k in ("uv", "uv_proj", "labels", "lines"):
        assert np.array_equal(s[k], d["s_" + k]), k
    assert np.array_equal(gt["rgb"], d["gt_rgb"]) and np.array_equal(gt["lines2d"], d["gt_lines2d"])
    assert mask[s["sampling_idx"]].all() and len(set(s["sampling_idx"].tolist())) == R
    full, gt_full = DO.getitem((H, W), rgb, lines, mask, labels, att)
    n = d["full_uv_head"].shape[0]
    assert np.array_equal(full["uv"][:n], d["full_uv_head"]) and np.array_equal(full["lines"][:n], d["full_lines_head"])
    assert gt_full["rgb"].shape == (H * W, 3) and "lines2d" not in gt_full


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 17, 1000, 15030, 65536, 65537])
def test_pixel_permutation_is_a_bijection_and_matches_mirror(n):
    from neat_b200 import dataset as D
    got = D.pixel_permutation(n, 1234, 7, 0, n)
    assert sorted(got) == list(range(n))
    assert np.array_equal(np.asarray(got, dtype=np.int64), DO.pixel_permutation(n, 1234, 7, 0, n))
    if n >= 1000:                                   # a different step or seed is a different permutation
        assert D.pixel_permutation(n, 1234, 8, 0, 64) != got[:64]
        assert D.pixel_permutation(n, 1235, 7, 0, 64) != got[:64]
        assert D.pixel_permutation(n, 1234, 7, 10, 20) == got[10:30]


@pytest.mark.parametrize("seed,step", [(2 ** 63 + 12345, 2 ** 40 + 7), (0, 0), (2 ** 64 - 1, 2 ** 64 - 1)])
def test_pixel_permutation_64bit_keys(seed, step):
    from neat_b200 import dataset as D
    got = np.asarray(D.pixel_permutation(100003, seed, step, 0, 5000))
    assert np.array_equal(got, DO.pixel_permutation(100003, seed, step, 0, 5000)) and np.unique(got).size == 5000


def test_pixel_permutation_prefix_is_unbiased():
    """The first 1024 of 1.92 M positions (one training step of a 1600x1200 image), over 64 steps: position deciles are
    hit uniformly (chi-square, 9 dof; 99.9 % quantile = 27.9) and no position repeats suspiciously often."""
    from neat_b200 import dataset as D
    n, R = 1920000, 1024
    pos = np.concatenate([np.asarray(D.pixel_permutation(n, 99, s, 0, R)) for s in range(1, 65)])
    hist = np.bincount(pos * 10 // n, minlength=10).astype(np.float64)
    exp = pos.size / 10.0
    assert ((hist - exp) ** 2 / exp).sum() < 27.9
    assert np.unique(pos).size > 0.98 * pos.size
    assert np.array_equal(pos[:R], DO.pixel_permutation(n, 99, 1, 0, R))


def test_pixel_permutation_rejects_bad_arguments():
    from neat_b200 import _lib, dataset as D
    with pytest.raises(_lib.NeatError):
        D.pixel_permutation(10, 0, 0, 5, 6)          # first + count > n
    with pytest.raises(_lib.NeatError):
        D.pixel_permutation(0, 0, 0, 0, 0)


def test_reference_numpy_positions_consume_the_stream_like_blender_dataset():
    """BlenderDataset draws `np.random.choice(sampling_idx, R)` (with replacement); drawing POSITIONS with the same seed
    selects the same pixels."""
    from neat_b200 import dataset as D
    masked = np.sort(np.random.RandomState(0).permutation(512 * 512)[:15030])
    np.random.seed(123)
    ref = np.random.choice(torch.from_numpy(masked), 1024)              # the reference call, on its tensor argument
    np.random.seed(123)
    pos = D.reference_numpy_positions(masked.size, 1024)
    assert np.array_equal(masked[pos], np.asarray(ref)) and np.unique(pos).size < 1024   # with replacement


def test_device_scene_has_no_cpu_path():
    from neat_b200 import _lib, dataset as D
    with pytest.raises(_lib.NeatError):
        D.DeviceScene((4, 4), device="cpu")
    with pytest.raises(_lib.NeatError):
        D.nonzero_mask(torch.ones(4, dtype=torch.bool))


# ------------------------------------------------------------------------------------------------------------ GPU
def _scene(rng, seed=0):
    from neat_b200 import dataset as D
    (H, W), lines, mask, labels, att, rgb = _tables()
    sc = D.DeviceScene((H, W), device="cuda:0", rng=rng, seed=seed)
    sc.add_image(rgb, lines, np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32), wireframe=None,
                 tables=(mask, labels, att))
    return sc, ((H, W), lines, mask, labels, att, rgb)


@pytest.mark.gpu
def test_gpu_getitem_reference_rng_matches_reference_bit_exact():
    d = np.load(os.path.join(GOLDEN, "dataset_abc.npz"))
    sc, _ = _scene("reference")
    torch.manual_seed(int(d["seed"]))
    sc.change_sampling_idx(int(d["R"]))
    idx, s, gt = sc[0]
    for k in ("uv", "uv_proj", "labels", "lines"):
        assert np.array_equal(s[k].cpu().numpy(), d["s_" + k]), k
    assert np.array_equal(gt["rgb"].cpu().numpy(), d["gt_rgb"])
    assert np.array_equal(gt["lines2d"].cpu().numpy(), d["gt_lines2d"])
    # through the DataLoader + collate_fn of volsdf_train.py:155-159: a leading batch dimension, lists for non-tensors
    # (the loader's iterator draws its base seed from the same generator first, so this is a different subset)
    loader = torch.utils.data.DataLoader(sc, batch_size=1, shuffle=False, collate_fn=sc.collate_fn)
    ind, mi, g2 = next(iter(loader))
    assert ind.tolist() == [0] and mi["uv"].shape == (1, int(d["R"]), 2) and mi["wireframe"] == [None]
    assert mi["intrinsics"].shape == (1, 4, 4) and g2["lines2d"].shape == (1, int(d["R"]), 5)
    rgb_image = np.random.default_rng(7).random((512 * 512, 3), dtype=np.float32)
    assert np.array_equal(g2["rgb"][0].cpu().numpy(), rgb_image[mi["sampling_idx"][0].cpu().numpy()])


@pytest.mark.gpu
def test_gpu_getitem_device_rng_and_full_image():
    from neat_b200 import dataset as D
    sc, ((H, W), lines, mask, labels, att, rgb) = _scene("device", seed=5)
    R = 1024
    sc.change_sampling_idx(R)
    nz = np.nonzero(mask)[0]
    assert np.array_equal(sc.images[0].masked.cpu().numpy(), nz)
    seen = []
    for step in (1, 2):
        _, s, gt = sc[0]
        pos = np.asarray(D.pixel_permutation(nz.size, 5, step, 0, R))
        o, ogt = DO.getitem((H, W), rgb, lines, mask, labels, att, perm=pos, num_pixels=R)
        assert np.array_equal(s["sampling_idx"].cpu().numpy(), o["sampling_idx"])
        for k in ("uv", "uv_proj", "labels", "lines"):
            assert np.array_equal(s[k].cpu().numpy(), o[k]), k
        assert np.array_equal(gt["rgb"].cpu().numpy(), ogt["rgb"])
        assert len(set(o["sampling_idx"].tolist())) == R and mask[o["sampling_idx"]].all()
        seen.append(o["sampling_idx"])
    assert not np.array_equal(seen[0], seen[1])
    sc.change_sampling_idx(-1)
    _, s, gt = sc[0]
    o, ogt = DO.getitem((H, W), rgb, lines, mask, labels, att)
    for k in ("uv", "uv_proj", "labels", "lines"):
        assert np.array_equal(s[k].cpu().numpy(), o[k]), k
    assert np.array_equal(gt["rgb"].cpu().numpy(), ogt["rgb"]) and "lines2d" not in gt
    cs, cgt = sc.full_image_chunk(0, 1000, 333)                   # ragged chunk of the full image
    assert np.array_equal(cs["uv"].cpu().numpy(), o["uv"][1000:1333])
    assert np.array_equal(cgt["rgb"].cpu().numpy(), ogt["rgb"][1000:1333])
    with pytest.raises(Exception):
        sc.change_sampling_idx(nz.size + 1)
        sc[0]


@pytest.mark.gpu
@pytest.mark.parametrize("n,density", [(1, 1.0), (1, 0.0), (1023, 0.5), (1025, 0.01), (1200 * 1600, 0.3),
                                       (1024 * 1030 + 7, 1.0), (5000, 0.0)])
def test_gpu_nonzero_mask(n, density):
    from neat_b200 import dataset as D
    rs = np.random.RandomState(n % 1000)
    m = rs.rand(n) < density
    got = D.nonzero_mask(torch.from_numpy(m).cuda())
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), np.nonzero(m)[0])
    got8 = D.nonzero_mask(torch.from_numpy(m.astype(np.uint8) * 3).cuda())   # any non-zero byte counts
    assert np.array_equal(got8.cpu().numpy(), np.nonzero(m)[0])


@pytest.mark.gpu
def test_gpu_scene_feeds_the_train_step():
    """The reference's loop (volsdf_train.py:355-374) on DeviceScene items: the collated item is the model input as is
    (eval forward identical to the same rays fed from host arrays), and Adam steps run on its batches."""
    from neat_b200 import dataset as D, synth, trainer as TR
    (H, W), lines, mask, labels, att, rgb = _tables(20.0)
    cam = synth.make_batch(1, seed=1, img_res=(H, W), focal=560.0)
    verts = np.unique(lines[:, :4].reshape(-1, 2), axis=0)
    sc = D.DeviceScene((H, W), device="cuda:0", rng="device", seed=11)
    sc.add_image(rgb, lines, cam["intrinsics"][0], cam["pose"][0], wireframe=TR.Wireframe(verts), tables=(mask, labels, att))
    R = 256
    sc.change_sampling_idx(R)
    loader = torch.utils.data.DataLoader(sc, batch_size=1, shuffle=True, collate_fn=sc.collate_fn)
    ts = TR.TrainStep(synth.toy_conf(), device="cuda:0", seed=0, beta=0.1, rng="device")
    _, mi, gt = next(iter(loader))
    assert mi["uv"].is_cuda and mi["uv"].shape == (1, R, 2) and gt["lines2d"].shape == (1, R, 5)
    ts.model.eval()
    with torch.no_grad():
        a = ts.model(mi)
        idx = mi["sampling_idx"][0].cpu().numpy()
        host = {"intrinsics": torch.from_numpy(cam["intrinsics"]).cuda(), "pose": torch.from_numpy(cam["pose"]).cuda(),
                "uv": torch.from_numpy(DO.uv_grid((H, W))[idx])[None].cuda(), "uv_proj": torch.from_numpy(att[idx])[None].cuda(),
                "wireframe": mi["wireframe"]}
        b = ts.model(host)
    for k in ("rgb_values", "lines3d", "lines2d", "sdf"):
        assert torch.allclose(a[k], b[k], rtol=1e-6, atol=1e-6), k
    ts.model.train()
    before = ts.model.implicit_network.lin0.weight_v.detach().clone()
    losses = []
    for _ in range(4):
        _, mi, gt = next(iter(loader))           # a fresh subset of the image every step
        losses.append(float(ts.step(mi, gt).detach()))
    assert all(np.isfinite(l) for l in losses) and bool(torch.isfinite(ts.bucket.flat).all())
    assert float((ts.model.implicit_network.lin0.weight_v.detach() - before).abs().max()) > 0
