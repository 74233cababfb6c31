"""The benchmarked batch through BOTH implementations on the same GPU: the drop-in plugin with model.rng = "reference" (the
reference's own CPU-generator calls in the reference's order) and the UNMODIFIED reference classes (oracle/_ref/
neat_ref_code.zip or /root/reference, oracle/ref_shim.py), built under the same seed -- identical initial weights, batch
and random draws, so the first training forward + loss must agree term by term (code/training/volsdf_train.py:366-367).
Measured on B200 (scripts/step0_vs_reference.py): every loss term within 1e-6 relative, rgb_values 7e-7, lines3d 5e-6."""
import pytest
import torch

from neat_b200 import synth
from neat_b200 import trainer as TR

pytestmark = pytest.mark.gpu


def test_first_step_of_the_benchmark_batch_equals_the_unmodified_reference():
    from oracle import ref_bench, ref_shim
    if not ref_shim.available():
        pytest.skip("the reference is not available (neither /root/reference nor oracle/_ref/neat_ref_code.zip)")
    R = 1024
    dev = torch.device("cuda", 0)
    hb = TR.host_batch(R, seed=1)
    inp, gt = TR.to_device(hb, dev)
    ps = TR.TrainStep(synth.dtu_conf(), device=dev, seed=42, beta=0.1, rng="reference")
    out = ps.model(inp)
    lo = ps.loss_fn(out, gt)
    try:
        ref_shim.install()
        ref_shim.force_cpu(False)
        model, loss_fn = ref_bench._build(synth.dtu_conf(), 42, 0.1, "cuda:0")
        model.train()
        w = hb["wireframe"][0]
        wf = ref_shim.Wireframe(w.vertices.numpy(), w.edges.numpy(), w.weights.numpy())
        mi = {k: hb[k].to(dev) for k in ("intrinsics", "uv", "pose", "uv_proj")}
        mi["wireframe"] = [wf]
        o2 = model(mi)
        l2 = loss_fn(o2, {"rgb": hb["rgb"], "lines2d": hb["lines2d"]})
    finally:
        ref_shim.force_cpu(not torch.cuda.is_available())
    assert int(out["j3d_local"].shape[0]) == int(o2["j3d_local"].shape[0])
    for k in ("loss", "rgb_loss", "eikonal_loss", "line_loss", "l2d_loss", "j3d_loss", "j2d_loss", "j2d_stat"):
        a, b = float(lo[k]), float(l2[k])
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (k, a, b)
    assert int(lo["count"]) == int(l2["count"]) and int(lo["jcount"]) == int(l2["jcount"])
    for k in ("rgb_values", "lines3d", "lines2d_calib"):
        a, b = out[k].detach().float().cpu(), o2[k].detach().float().cpu()
        assert float((a - b).abs().max() / b.abs().max()) < 1e-4, k
