"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the dataset-side attraction precompute.

  encodels(...)                 restates hawp.base._C.encodels, i.e. `encode_kernel`
                                (third-party/hawp/hawp/base/csrc/linesegment.cu:23-103), in numpy.
  point_line_attraction(...)    restates SceneDataset.compute_point_line_attraction
                                (code/datasets/scene_hawp_dataset.py:92-146).

Parity status: PINNED ON THE GPU BOX.  The reference kernel needs a GPU, so it cannot run where the reference tree
lives; instead oracle/build_ref.py compiles the reference's own two source files for sm_100a into oracle/_ref/ (which
travels to the B200 box) and tests/test_hawp_oracle.py::test_gpu_encodels_vs_reference_kernel compares this
restatement -- and the product kernel -- with that binary bit for bit.  The restatement follows the arithmetic nvcc
gives the kernel at its default flags, read off its SASS: `a*a + b*b` is fma(a, a, rn(b*b)), `x1 + t*dx` is
fma(t, dx, x1), the division runs in double (the `1e-6` literal).  `point_line_attraction` is pinned here as well: oracle/make_golden_hawp.py runs the UNMODIFIED
reference method with `_C.encodels` replaced by the restatement above and stores its outputs in
tests/golden/hawp_abc.npz.  Only tests/ may import this module."""
import numpy as np


def _fma(a, b, c):
    """float32 fma(a, b, c): the product of two float32 is exact in float64; the sum is rounded to float64 and then to
    float32 (double rounding can differ from a true fma only when the float64 sum sits exactly on a float32 tie)."""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64) + np.asarray(c, dtype=np.float64)).astype(np.float32)


def encodels(lines, input_height, input_width, height, width, num_lines):
    """lines [n,4] float32 (x1,y1,x2,y2) -> map [6,H,W] f32, label [n,H,W] bool, tmap [1,H,W] f32."""
    lines = np.asarray(lines, dtype=np.float32)
    H, W, n = int(height), int(width), int(num_lines)
    px = np.broadcast_to(np.arange(W, dtype=np.float32)[None, :], (H, W))
    py = np.broadcast_to(np.arange(H, dtype=np.float32)[:, None], (H, W))
    f32 = np.float32
    xs = f32(f32(W) / f32(input_width))
    ys = f32(f32(H) / f32(input_height))
    min_dis = np.full((H, W), 1e30, dtype=np.float32)
    minp = np.full((H, W), -1, dtype=np.int64)
    flagp = np.ones((H, W), dtype=bool)
    mp = np.zeros((6, H, W), dtype=np.float32)
    tmap = np.zeros((1, H, W), dtype=np.float32)
    for i in range(n):
        x1, y1, x2, y2 = f32(lines[i, 0] * xs), f32(lines[i, 1] * ys), f32(lines[i, 2] * xs), f32(lines[i, 3] * ys)
        dx, dy = f32(x2 - x1), f32(y2 - y1)
        ux, uy, vx, vy = x1 - px, y1 - py, x2 - px, y2 - py
        norm2 = f32(_fma(dx, dx, f32(dy * dy)))
        num = _fma((px - x1).astype(np.float32), dx, ((py - y1).astype(np.float32) * dy).astype(np.float32))
        t = (num.astype(np.float64) / (np.float64(norm2) + 1e-6)).astype(np.float32)   # `1e-6` is a double literal
        flag = (t <= 1) & (t >= 0.0)
        t = np.clip(t, 0.0, 1.0).astype(np.float32)
        ax = (_fma(t, dx, x1) - px).astype(np.float32)
        ay = (_fma(t, dy, y1) - py).astype(np.float32)
        dis = _fma(ax, ax, (ay * ay).astype(np.float32))
        upd = dis < min_dis
        min_dis = np.where(upd, dis, min_dis)
        nu2 = _fma(ux, ux, (uy * uy).astype(np.float32))
        nv2 = _fma(vx, vx, (vy * vy).astype(np.float32))
        first = nu2 < nv2
        mp[0] = np.where(upd, ax, mp[0])
        mp[1] = np.where(upd, ay, mp[1])
        mp[2] = np.where(upd, np.where(first, ux, vx), mp[2])
        mp[3] = np.where(upd, np.where(first, uy, vy), mp[3])
        mp[4] = np.where(upd, np.where(first, vx, ux), mp[4])
        mp[5] = np.where(upd, np.where(first, vy, uy), mp[5])
        minp = np.where(upd, i, minp)
        flagp = np.where(upd, flag, flagp)
        tmap[0] = np.where(upd, t, tmap[0])
    label = np.zeros((n, H, W), dtype=bool)
    hh, ww = np.nonzero(minp >= 0)
    label[minp[hh, ww], hh, ww] = flagp[hh, ww]
    return mp, label, tmap


def point_line_attraction(lines, img_res, distance):
    """SceneDataset.compute_point_line_attraction (scene_hawp_dataset.py:92-146).
    lines [n,>=4] -> mask [HW] bool, labels [HW] int64, proj_points [HW,2] float32."""
    H, W = int(img_res[0]), int(img_res[1])
    lines = np.asarray(lines, dtype=np.float32)[:, :4]
    lmap, onehot, _ = encodels(lines, H, W, H, W, lines.shape[0])
    mask = onehot.max(axis=0)
    labels = onehot.argmax(axis=0).astype(np.int64)

    def norm2(v):
        mag = np.sqrt(v[0] * v[0] + v[1] * v[1])
        return v / (mag + np.float32(1e-6))

    dismap = np.sqrt(lmap[0] ** 2 + lmap[1] ** 2)
    md = norm2(lmap[:2]).reshape(2, -1)
    st, ed = lmap[2:4].reshape(2, -1), lmap[4:].reshape(2, -1)
    # Rt = [[mx, my], [-my, mx]] applied to st / ed
    rst = np.stack([md[0] * st[0] + md[1] * st[1], -md[1] * st[0] + md[0] * st[1]])
    red = np.stack([md[0] * ed[0] + md[1] * ed[1], -md[1] * ed[0] + md[0] * ed[1]])
    swap = (rst[1] < 0) & (red[1] > 0)
    pos, neg = rst.copy(), red.copy()
    pos[:, swap], neg[:, swap] = red[:, swap], rst[:, swap]
    pos[0] = np.maximum(pos[0], 1e-9); pos[1] = np.maximum(pos[1], 1e-9)
    neg[0] = np.maximum(neg[0], 1e-9); neg[1] = np.minimum(neg[1], -1e-9)
    mask = (dismap <= distance) & mask
    pos_angle = np.arctan2(pos[1], pos[0]).reshape(H, W)
    neg_angle = np.arctan2(neg[1], neg[0]).reshape(H, W)
    mask = mask & (pos_angle > 0) & (neg_angle < 0)
    proj = np.zeros((H, W, 2), dtype=np.float32)
    hh, ww = np.nonzero(mask)
    proj[hh, ww, 0] = lmap[0][hh, ww] + ww.astype(np.float32)
    proj[hh, ww, 1] = lmap[1][hh, ww] + hh.astype(np.float32)
    return mask.reshape(-1), labels.reshape(-1), proj.reshape(-1, 2)
