"""CPU: the native junction matching (csrc/junction.cpp) against scipy and a numpy restatement of the reference's
junction block (neat_wfr_rend_a.py:466-484, loss_wfr.py:104-108)."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment as scipy_lsa

from neat_b200 import build, junction, synth


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build()


@pytest.mark.parametrize("shape", [(1, 1), (5, 5), (200, 60), (60, 200), (3, 1024), (17, 1), (1, 9), (128, 128)])
def test_assignment_matches_scipy(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for trial in range(4):
        cost = rng.random(shape) * (10.0 if trial % 2 else 1.0)
        if trial == 3:
            cost = np.round(cost, 1)  # many ties: the optimum value must still agree
        r, c = junction.linear_sum_assignment(cost)
        r0, c0 = scipy_lsa(cost)
        assert len(r) == min(shape) and np.array_equal(r, np.sort(r)) and len(set(c.tolist())) == len(c)
        assert np.array_equal(r, r0)
        assert abs(cost[r, c].sum() - cost[r0, c0].sum()) <= 1e-9 * max(1.0, abs(cost[r0, c0].sum()))
        if trial < 3:
            assert np.array_equal(c, c0)


def test_assignment_edge_cases():
    r, c = junction.linear_sum_assignment(np.zeros((0, 5)))
    assert len(r) == 0 and len(c) == 0
    with pytest.raises(ValueError):
        junction.linear_sum_assignment(np.array([[np.nan, 1.0], [1.0, 2.0]]))
    with pytest.raises(ValueError):
        junction.linear_sum_assignment(np.full((2, 2), np.inf))
    # a forbidden pair (inf) is routed around
    r, c = junction.linear_sum_assignment(np.array([[np.inf, 1.0], [1.0, np.inf]]))
    assert c.tolist() == [1, 0]


def _reference_block(cent, gt, pose, K4, glob, use_median):
    RT = np.linalg.inv(pose.astype(np.float64))[:3]

    def proj(Km, X):
        x = (Km @ (RT[:, :3] @ X.T.astype(np.float64) + RT[:, 3:])).T
        den = x[:, 2:3]
        sign = np.where(den >= 0, 1.0, -1.0)
        eps = np.where(np.abs(den) < 1e-8, 1e-8, 0.0)
        return (x / (den + eps * sign))[:, :2]

    K3 = K4[:3, :3].astype(np.float64)
    j2d, j2c = proj(K3, cent), proj(np.eye(3), cent)
    jcost = np.sqrt(((j2d[None] - gt[:, None]) ** 2).sum(-1))
    a0, a1 = scipy_lsa(jcost)
    sel = jcost[a0, a1]
    thr = np.sort(sel)[(len(sel) - 1) // 2] if use_median else 10.0  # torch.median: the lower middle value
    ok = sel < thr
    j3l, j2l, j2lc = cent[a1][ok], j2d[a1][ok], j2c[a1][ok]
    gcal = proj(np.eye(3), glob)
    cost = np.abs(j3l[:, None] - glob[None]).sum(-1) + 0.1 * np.abs(j2lc[:, None] - gcal[None]).sum(-1)
    b0, b1 = scipy_lsa(cost)
    return np.concatenate([j3l, j2l, j2lc], 1), b0, b1, int((cost[b0, b1] < 10).sum()), thr


@pytest.mark.parametrize("use_median", [False, True])
def test_junction_block_matches_reference_restatement(use_median):
    rng = np.random.default_rng(7)
    b = synth.make_batch(64, seed=3)
    pose, K4 = b["pose"][0], b["intrinsics"][0]
    gt = b["wf_vertices"].astype(np.float32)
    # centroids that project near some of the ground-truth junctions: back-project gt pixels at depth ~2.5
    Kinv = np.linalg.inv(K4[:3, :3].astype(np.float64))
    n = 40
    pick = rng.choice(gt.shape[0], n, replace=False)
    pix = np.concatenate([gt[pick] + rng.normal(0, 3.0, (n, 2)), np.ones((n, 1))], 1)
    cam_pts = (Kinv @ pix.T).T * rng.uniform(2.0, 3.0, (n, 1))
    world = (pose[:3, :3].astype(np.float64) @ cam_pts.T + pose[:3, 3:].astype(np.float64)).T
    cent = np.concatenate([world, rng.uniform(-1, 1, (15, 3))], 0).astype(np.float32)
    glob = rng.uniform(-1, 1, (1024, 3)).astype(np.float32)
    local, rows, cols, close, med = junction.junction_match(cent, gt, pose, K4, glob, use_median)
    ref_local, b0, b1, ref_close, thr = _reference_block(cent.astype(np.float64), gt.astype(np.float64), pose, K4,
                                                         glob.astype(np.float64), use_median)
    assert local.shape == ref_local.shape and local.shape[0] >= (10 if not use_median else 5)
    np.testing.assert_allclose(local[:, :3], ref_local[:, :3], rtol=0, atol=0)  # centroids are copied
    np.testing.assert_allclose(local[:, 3:5], ref_local[:, 3:5], rtol=1e-5, atol=2e-3)  # pixels (~1e3 magnitude)
    np.testing.assert_allclose(local[:, 5:7], ref_local[:, 5:7], rtol=1e-5, atol=1e-6)
    assert np.array_equal(rows, b0) and np.array_equal(cols, b1) and close == ref_close
    if use_median:
        assert abs(med - thr) < 1e-3


def test_junction_block_empty_inputs():
    b = synth.make_batch(8, seed=1)
    glob = np.zeros((16, 3), np.float32)
    local, rows, cols, close, _ = junction.junction_match(np.zeros((0, 3), np.float32), b["wf_vertices"], b["pose"][0],
                                                          b["intrinsics"][0], glob)
    assert local.shape == (0, 7) and len(rows) == 0 and close == 0
    local, rows, cols, close, _ = junction.junction_match(np.ones((3, 3), np.float32) * 100, b["wf_vertices"],
                                                          b["pose"][0], b["intrinsics"][0], glob)
    assert local.shape[0] == 0  # nothing within 10 px
