"""Measured errors behind tests/test_gpu_parity.py::test_train_step_vs_reference (training forward INCLUDING the sampler,
replayed CPU-generator draws, against the unmodified reference's goldens).  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io as G
import test_gpu_parity as TP

for name in TP.CASES:
    g, conf, sd_np, model, out, lo = TP.run_train(name)
    errs = {k: G.rel_err(out[k].detach().cpu(), g["train_" + k]) for k in ("rgb_values", "lines3d", "lines2d", "lines2d_calib", "j3d_global", "j3d_local")}
    le = {k: abs(float(lo[k]) - float(g["loss_" + k])) / max(1.0, abs(float(g["loss_" + k])))
          for k in ("rgb_loss", "line_loss", "l2d_loss", "j3d_loss", "j2d_loss", "eikonal_loss", "loss")}
    print(name, {k: float("%.2e" % v) for k, v in errs.items()}, {k: float("%.2e" % v) for k, v in le.items()}, flush=True)
