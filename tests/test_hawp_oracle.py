"""CPU: the attraction-precompute oracle vs goldens from the unmodified reference method (oracle/make_golden_hawp.py);
GPU: the fused kernels vs the oracle (bit-exact: integer / index work and IEEE fp32 arithmetic without contraction)."""
import os

import numpy as np
import pytest
import torch

from oracle import hawp_oracle as HO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hawp_abc.npz")


@pytest.mark.parametrize("dist", [5.0, 20.0])
def test_point_line_attraction_oracle_vs_reference(dist):
    g = np.load(GOLD)
    H, W = [int(x) for x in g["img_res"]]
    mask, labels, proj = HO.point_line_attraction(g["lines"], (H, W), dist)
    ref_mask = np.unpackbits(g["mask_%g" % dist])[: H * W].astype(bool)
    assert np.array_equal(mask, ref_mask)
    assert np.array_equal(labels[ref_mask], g["labels_%g" % dist].astype(np.int64)[ref_mask])
    idx = g["proj_idx_%g" % dist]
    assert np.array_equal(proj[idx], g["proj_val_%g" % dist])
    assert not proj[~ref_mask].any()


def test_encodels_properties():
    rs = np.random.RandomState(0)
    lines = rs.uniform(0, 64, size=(7, 4)).astype(np.float32)
    mp, label, tmap = HO.encodels(lines, 64, 64, 48, 80, 7)       # output grid != input grid: coordinates are rescaled
    assert mp.shape == (6, 48, 80) and label.shape == (7, 48, 80) and tmap.shape == (1, 48, 80)
    assert label.sum(axis=0).max() <= 1                            # at most one line per pixel
    assert (tmap >= 0).all() and (tmap <= 1).all()
    m, l, t = HO.encodels(np.zeros((0, 4), np.float32), 8, 8, 8, 8, 0)   # no lines: everything stays zero
    assert not m.any() and l.shape == (0, 8, 8)


@pytest.mark.gpu
@pytest.mark.parametrize("dist", [5.0, 20.0])
def test_gpu_point_line_attraction_bit_exact(dist):
    from neat_b200 import attraction as A
    g = np.load(GOLD)
    H, W = [int(x) for x in g["img_res"]]
    mask, labels, proj = A.compute_point_line_attraction(torch.from_numpy(g["lines"]).cuda(), (H, W), dist)
    om, ol, op = HO.point_line_attraction(g["lines"], (H, W), dist)
    assert np.array_equal(mask.cpu().numpy(), om)
    assert np.array_equal(labels.cpu().numpy()[om], ol[om])
    assert np.array_equal(proj.cpu().numpy(), op)


@pytest.mark.gpu
@pytest.mark.parametrize("n,H,W,ih,iw", [(9, 512, 512, 512, 512), (1500, 96, 130, 192, 260), (1, 7, 5, 7, 5)])
def test_gpu_encodels_bit_exact(n, H, W, ih, iw):
    """hawp.base._C.encodels drop-in vs the scalar restatement, incl. > 1024 lines (two shared-memory passes),
    a rescaled grid and a ragged tiny image."""
    from neat_b200 import attraction as A
    rs = np.random.RandomState(n)
    lines = np.stack([rs.uniform(0, iw, n), rs.uniform(0, ih, n), rs.uniform(0, iw, n), rs.uniform(0, ih, n)], -1).astype(np.float32)
    mp, label, tmap = A.encodels(torch.from_numpy(lines).cuda(), ih, iw, H, W, n)
    om, ol, ot = HO.encodels(lines, ih, iw, H, W, n)
    assert np.array_equal(mp.cpu().numpy(), om)
    assert np.array_equal(label.cpu().numpy(), ol)
    assert np.array_equal(tmap.cpu().numpy(), ot)


@pytest.mark.gpu
@pytest.mark.parametrize("n,H,W,ih,iw", [(9, 512, 512, 512, 512), (1500, 96, 130, 192, 260), (1, 7, 5, 7, 5), (40, 300, 400, 1200, 1600)])
def test_gpu_encodels_vs_reference_kernel(n, H, W, ih, iw):
    """The REAL reference kernel (hawp.base._C.encodels, built from the reference's own two source files for sm_100a by
    oracle/build_ref.py into oracle/_ref/) against the product kernel and against the numpy restatement, on the same
    inputs, bit for bit.  This is what pins `encode_kernel`."""
    from neat_b200 import attraction as A
    from oracle import build_ref
    ref = build_ref.load_built()
    if ref is None:
        pytest.skip("oracle/_ref/hawp_ref_C.so is not built (python oracle/build_ref.py needs the reference tree)")
    rs = np.random.RandomState(100 + n)
    lines = np.stack([rs.uniform(0, iw, n), rs.uniform(0, ih, n), rs.uniform(0, iw, n), rs.uniform(0, ih, n)], -1).astype(np.float32)
    if n == 9:                                            # the in-repo ABC wireframe (incl. pixels equidistant to two lines)
        lines = np.load(GOLD)["lines"][:, :4].astype(np.float32)
    rm, rl, rt = ref.encodels(torch.from_numpy(lines).cuda(), ih, iw, H, W, n)
    torch.cuda.synchronize()
    mp, label, tmap = A.encodels(torch.from_numpy(lines).cuda(), ih, iw, H, W, n)
    assert torch.equal(mp, rm) and torch.equal(label, rl) and torch.equal(tmap, rt)
    om, ol, ot = HO.encodels(lines, ih, iw, H, W, n)
    assert np.array_equal(rm.cpu().numpy(), om) and np.array_equal(rl.cpu().numpy(), ol) and np.array_equal(rt.cpu().numpy(), ot)
