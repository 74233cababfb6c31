"""Deterministic synthetic configs, weights and DTU-shaped batches (numpy only).

There is no dataset or checkpoint in the image, so benchmarks, tests and golden vectors all
use the inputs made here.  Everything is drawn from ``numpy.random.RandomState`` (MT19937),
whose streams are stable across numpy versions and machines, so the same seed gives the same
bytes in the build container (where golden vectors are produced from the reference) and on
the GPU box.

Config values follow ``/root/reference/code/confs/dtu.conf:28-87``; the toy config is
BASELINE.json configs[0] (4x128 nets, 256 rays x 64 samples; not a shipped conf).
"""
import math

import numpy as np


def dtu_conf():
    return {
        "feature_vector_size": 256,
        "scene_bounding_sphere": 3.0,
        "dbscan_enabled": True,
        "use_median": False,
        "global_junctions": {"num_junctions": 1024, "num_layers": 2, "dim_out": 3, "dim_hidden": 256},
        "implicit_network": {"d_in": 3, "d_out": 1, "dims": [256] * 8, "geometric_init": True,
                             "bias": 0.6, "skip_in": [4], "weight_norm": True, "multires": 6,
                             "sphere_scale": 20.0},
        "attraction_network": {"d_in": 9, "d_out": 6, "dims": [256] * 4, "mode": "idr",
                               "weight_norm": True},
        "rendering_network": {"mode": "idr", "d_in": 9, "d_out": 3, "dims": [256] * 4,
                              "weight_norm": True, "multires_view": 4},
        "density": {"params_init": {"beta": 0.1}, "beta_min": 0.0001},
        "ray_sampler": {"near": 0.0, "N_samples": 64, "N_samples_eval": 128, "N_samples_extra": 32,
                        "eps": 0.1, "beta_iters": 10, "max_total_iters": 5},
    }


def abc_conf():
    """model{} of code/confs/abc-neat-a.conf:28-87: the DTU nets and sampler, but every attraction end point is a
    junction candidate (dbscan_enabled = False), the match filter is the median cost, and 64 global junctions."""
    c = dtu_conf()
    c.update(dbscan_enabled=False, use_l3d=False, use_median=True)
    c["global_junctions"] = {"num_junctions": 64, "num_layers": 2, "dim_out": 3, "dim_hidden": 256}
    return c


def toy_conf():
    c = dtu_conf()
    c["feature_vector_size"] = 128
    c["global_junctions"] = {"num_junctions": 64, "num_layers": 2, "dim_out": 3, "dim_hidden": 64}
    c["implicit_network"].update(dims=[128] * 4, skip_in=[2])
    c["attraction_network"].update(dims=[128] * 4)
    c["rendering_network"].update(dims=[128] * 4)
    c["ray_sampler"].update(N_samples=40, N_samples_eval=64, N_samples_extra=22)
    return c


def toy_white_conf():
    """toy_conf with the two optional branches of the model class no shipped conf selects: white_bkgd (background colour,
    ImplicitNetwork without the sphere clamp, neat_wfr_rend_a.py:262-266, 411-413) and junction_eikonal (:524-525)."""
    c = toy_conf()
    c.update(white_bkgd=True, bg_color=[1.0, 0.9, 0.8], junction_eikonal=True)
    return c


def toy_l3d_conf():
    """toy_conf with the third junction-candidate rule of the model class (neat_wfr_rend_a.py:461-465): no DBSCAN,
    use_l3d -- the rays whose tangent-plane point agrees best with their 3D line propose the junctions."""
    c = toy_conf()
    c.update(dbscan_enabled=False, use_l3d=True)
    return c


def loss_conf():
    return {"eikonal_weight": 0.1, "line_weight": 0.01, "rgb_loss": "torch.nn.L1Loss"}


def _embed_dim(multires, d=3):
    return d + 2 * d * multires if multires > 0 else d


def sdf_layer_dims(conf):
    """(in, out) of every ImplicitNetwork layer (neat_wfr_rend_a.py:35-53)."""
    c = conf["implicit_network"]
    d0 = _embed_dim(c["multires"], c["d_in"])
    dims = [d0] + list(c["dims"]) + [c["d_out"] + conf["feature_vector_size"]]
    out = []
    for l in range(len(dims) - 1):
        o = dims[l + 1] - d0 if (l + 1) in c["skip_in"] else dims[l + 1]
        out.append((dims[l], o))
    return out


def head_layer_dims(conf, which):
    c = conf[which]
    d0 = c["d_in"] + conf["feature_vector_size"]
    mv = c.get("multires_view", 0)
    if mv > 0:
        d0 += _embed_dim(mv) - 3
    dims = [d0] + list(c["dims"]) + [c["d_out"]]
    return [(dims[l], dims[l + 1]) for l in range(len(dims) - 1)]


def make_state_dict(conf, seed=0, perturb=0.15, beta=None):
    """A full VolSDFNetwork state dict (reference key names) as numpy
    float32 arrays.  Starts from the geometric initialisation (SDF ~ |x| - bias) and perturbs
    every entry -- including the positional-encoding columns that the geometric init zeroes --
    so that no code path is multiplied by an exact zero in parity tests.  perturb=0 gives a
    pure geometric-init-like network."""
    rs = np.random.RandomState(seed)
    sd = {}
    ci = conf["implicit_network"]
    dims = sdf_layer_dims(conf)
    L = len(dims)
    d0 = dims[0][0]
    for l, (i, o) in enumerate(dims):
        if l == L - 1:
            W = rs.normal(math.sqrt(math.pi) / math.sqrt(i), 1e-4, size=(o, i))
            # only the sdf row follows the geometric init; feature rows get a generic init
            W[1:] = rs.normal(0.0, math.sqrt(2.0) / math.sqrt(i), size=(o - 1, i))
            b = np.full(o, 0.0)
            b[0] = -ci["bias"]
            b[1:] = rs.normal(0.0, 0.05, size=o - 1)
        elif l == 0:
            W = np.zeros((o, i))
            W[:, :3] = rs.normal(0.0, math.sqrt(2.0) / math.sqrt(o), size=(o, 3))
            b = np.zeros(o)
        elif l in ci["skip_in"]:
            W = rs.normal(0.0, math.sqrt(2.0) / math.sqrt(o), size=(o, i))
            W[:, -(d0 - 3):] = 0.0
            b = np.zeros(o)
        else:
            W = rs.normal(0.0, math.sqrt(2.0) / math.sqrt(o), size=(o, i))
            b = np.zeros(o)
        if perturb > 0:
            W = W * (1.0 + perturb * rs.normal(size=W.shape))
            b = b + perturb * 0.02 * rs.normal(size=b.shape)
            if l == 0 or l in ci["skip_in"]:
                # small, frequency-damped weights on the PE columns keep the SDF smooth
                npe = d0 - 3
                cols = slice(3, None) if l == 0 else slice(i - npe, None)
                freq = np.repeat(2.0 ** np.arange(ci["multires"]), 6)
                W[:, cols] = perturb * 0.2 * rs.normal(size=(o, npe)) / freq[None, :] / math.sqrt(o)
        g = np.linalg.norm(W, axis=1, keepdims=True)
        if perturb > 0:
            g = g * (1.0 + 0.05 * perturb * rs.normal(size=g.shape))
        sd[f"implicit_network.lin{l}.bias"] = b
        sd[f"implicit_network.lin{l}.weight_g"] = g
        sd[f"implicit_network.lin{l}.weight_v"] = W
    for which, name in (("rendering_network", "rendering_network"), ("attraction_network", "attraction_network")):
        hd = head_layer_dims(conf, which)
        for l, (i, o) in enumerate(hd):
            k = 1.0 / math.sqrt(i)
            W = rs.uniform(-k, k, size=(o, i))
            b = rs.uniform(-k, k, size=o)
            if l == len(hd) - 1 and which == "attraction_network":
                W *= 0.3            # keep the attraction offsets in a plausible range
            g = np.linalg.norm(W, axis=1, keepdims=True) * (1.0 + 0.05 * perturb * rs.normal(size=(o, 1)))
            sd[f"{name}.lin{l}.bias"] = b
            sd[f"{name}.lin{l}.weight_g"] = g
            sd[f"{name}.lin{l}.weight_v"] = W
    sd["density.beta"] = np.array(conf["density"]["params_init"]["beta"] if beta is None else beta)
    cj = conf["global_junctions"]
    H = cj["dim_hidden"]
    sd["latents"] = rs.normal(size=(cj["num_junctions"], H))
    k = 1.0 / math.sqrt(H)
    for idx, o in zip((0, 2, 4), (H, H, 3)):
        sd[f"ffn.{idx}.weight"] = rs.uniform(-k, k, size=(o, H))
        sd[f"ffn.{idx}.bias"] = rs.uniform(-k, k, size=o)
    return {k_: np.asarray(v, dtype=np.float32) for k_, v in sd.items()}


def look_at_pose(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """Camera-to-world 4x4 (x right, y down, z forward) -- the `pose` convention of
    rend_util.get_camera_params (code/utils/rend_util.py:55-81)."""
    eye = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, dtype=np.float64))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    P = np.eye(4)
    P[:3, 0], P[:3, 1], P[:3, 2], P[:3, 3] = r, d, f, eye
    return P.astype(np.float32)


def make_wireframe(seed, n_junctions, n_edges, width, height):
    rs = np.random.RandomState(seed)
    v = np.stack([rs.uniform(0.15 * width, 0.85 * width, n_junctions),
                  rs.uniform(0.15 * height, 0.85 * height, n_junctions)], -1).astype(np.float32)
    e = np.stack([rs.randint(0, n_junctions, n_edges), rs.randint(0, n_junctions, n_edges)], -1)
    e[:, 1] = np.where(e[:, 0] == e[:, 1], (e[:, 1] + 1) % n_junctions, e[:, 1])
    w = rs.uniform(0.3, 1.0, n_edges).astype(np.float32)
    return v, e.astype(np.int64), w


def make_batch(R, seed=1, img_res=(1200, 1600), focal=2900.0, cam_dist=2.5, n_junctions=200,
               n_edges=300, pose=None, K=None):
    """One synthetic DTU-shaped training batch: a camera at distance
    2.5 looking at the origin, R random pixels, noisy uv_proj, uniform rgb, a random 2D
    wireframe and one GT line (+weight) per ray."""
    rs = np.random.RandomState(seed)
    H, W = img_res
    if K is None:
        K = np.eye(4, dtype=np.float32)
        K[0, 0] = K[1, 1] = focal
        K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    if pose is None:
        a = 0.3 + 0.7 * seed
        pose = look_at_pose((cam_dist * math.cos(a) * 0.9, cam_dist * math.sin(a) * 0.9, cam_dist * 0.436))
    uv = np.stack([rs.uniform(0, W - 1, R), rs.uniform(0, H - 1, R)], -1).astype(np.float32)
    uv_proj = (uv + rs.normal(size=uv.shape)).astype(np.float32)
    rgb = rs.uniform(0, 1, size=(R, 3)).astype(np.float32)
    verts, edges, ew = make_wireframe(seed + 1000, n_junctions, n_edges, W, H)
    pick = rs.randint(0, len(edges), R)
    lines = np.concatenate([verts[edges[pick, 0]], verts[edges[pick, 1]], ew[pick, None]], -1).astype(np.float32)
    return {
        "intrinsics": K[None].astype(np.float32), "pose": np.asarray(pose, np.float32)[None],
        "uv": uv[None], "uv_proj": uv_proj[None], "rgb": rgb[None], "lines2d": lines[None],
        "wf_vertices": verts, "wf_edges": edges, "wf_weights": ew,
    }
