"""Helpers shared by the CPU and GPU parity tests: load a golden case and rebuild its inputs."""
import os

import numpy as np
import torch

from neat_b200 import synth
from oracle import neat_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"toy_beta0.1": synth.toy_conf, "dtu_beta0.1": synth.dtu_conf, "dtu_beta0.01": synth.dtu_conf,
         "abc_beta0.1": synth.abc_conf, "toy_white_jeik": synth.toy_white_conf,
         "toy_l3d": synth.toy_l3d_conf}


def load(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    conf = CASES[name]()
    sd_np = synth.make_state_dict(conf, seed=int(g["seed_w"]), perturb=float(g["perturb"]), beta=float(g["beta"]))
    return g, conf, sd_np


def oracle_params(conf, sd_np, dtype=torch.float32, track=False):
    sd = {k: torch.from_numpy(v.copy()).to(dtype) for k, v in sd_np.items()}
    if track:
        for v in sd.values():
            v.requires_grad_(True)
    ci = conf["implicit_network"]
    P = O.params_from_state_dict(sd, skip_in=tuple(ci["skip_in"]), multires=ci["multires"],
                                 multires_view=conf["rendering_network"]["multires_view"],
                                 sphere_radius=conf["scene_bounding_sphere"], sphere_scale=ci["sphere_scale"],
                                 beta_min=conf["density"]["beta_min"], track=track)
    P.dbscan_enabled, P.use_median = bool(conf.get("dbscan_enabled", True)), bool(conf.get("use_median", False))
    if conf.get("white_bkgd", False):
        P.bg_color = torch.tensor([float(v) for v in conf.get("bg_color", [1.0, 1.0, 1.0])], dtype=dtype)
    P.junction_eikonal = bool(conf.get("junction_eikonal", False))
    P.use_l3d = bool(conf.get("use_l3d", False))
    return P, sd


def sampler_conf(conf):
    c = conf["ray_sampler"]
    return O.SamplerConf(near=c["near"], N_samples=c["N_samples"], N_samples_eval=c["N_samples_eval"],
                         N_samples_extra=c["N_samples_extra"], eps=c["eps"], beta_iters=c["beta_iters"],
                         max_total_iters=c["max_total_iters"])


def train_randoms(g):
    t = lambda a: torch.from_numpy(np.asarray(a))
    return O.TrainRandoms(O.SamplerRandoms(t(g["rnd_t_rand"]), t(g["rnd_u_final"]), t(g["rnd_extra_idx"]).long(),
                                           t(g["rnd_eik_idx"]).long()), t(g["rnd_eik_uniform"]))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
