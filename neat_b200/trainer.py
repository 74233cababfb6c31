"""One training step as the reference trainer runs it (code/training/volsdf_train.py:361-374):
model(input) -> loss(...) -> zero_grad -> backward -> [all-reduce] -> optimizer.step."""
import torch

from . import synth
from .loss import VolSDFLoss
from .model import VolSDFNetwork
from .optim import Adam
from .parallel import GradBucket


class Wireframe:
    """The members of utils.hawp_util.WireframeGraph the path touches (code/utils/hawp_util.py:7-95): `vertices` [J,2]
    (the model's junction block, the loss) and `line_segments()` [E,5] (dataset lines, finalisation)."""

    def __init__(self, vertices, edges=None, weights=None):
        self.vertices = torch.as_tensor(vertices, dtype=torch.float32)
        self.edges = None if edges is None else torch.as_tensor(edges, dtype=torch.long)
        self.weights = None if weights is None else torch.as_tensor(weights, dtype=torch.float32)

    def line_segments(self, threshold=0.05, device=None):
        """(x1, y1, x2, y2, weight) of the edges with weight > threshold (hawp_util.py:56-69)."""
        ok = self.weights > threshold
        lines = torch.cat((self.vertices[self.edges[ok, 0]], self.vertices[self.edges[ok, 1]], self.weights[ok][:, None]), dim=-1)
        return lines if device is None else lines.to(device)


def host_batch(R, seed, pinned=True, camera_seed=1):
    """A synthetic DTU-shaped batch as HOST tensors, shaped like the reference's
    dataloader output (scene_hawp_dataset.py:148-194).  `seed` draws the pixels / colours / wireframe; the camera
    is the one of `camera_seed` so that every data-parallel rank sees statistically identical work."""
    import math
    a = 0.3 + 0.7 * camera_seed
    pose = synth.look_at_pose((2.5 * math.cos(a) * 0.9, 2.5 * math.sin(a) * 0.9, 2.5 * 0.436))
    b = synth.make_batch(R, seed=seed, pose=pose)
    t = {k: torch.from_numpy(b[k]) for k in ("intrinsics", "pose", "uv", "uv_proj", "rgb", "lines2d")}
    if pinned and torch.cuda.is_available():
        t = {k: v.pin_memory() for k, v in t.items()}
    t["wireframe"] = [Wireframe(b["wf_vertices"], b["wf_edges"], b["wf_weights"])]
    return t


def to_device(hb, dev):
    inp = {k: hb[k].to(dev, non_blocking=True) for k in ("intrinsics", "pose", "uv", "uv_proj")}
    inp["wireframe"] = hb["wireframe"]
    gt = {"rgb": hb["rgb"].to(dev, non_blocking=True), "lines2d": hb["lines2d"].to(dev, non_blocking=True)}
    return inp, gt


def h2d_bytes(hb):
    return sum(hb[k].numel() * hb[k].element_size() for k in ("intrinsics", "pose", "uv", "uv_proj", "rgb", "lines2d"))


class TrainStep:
    def __init__(self, conf=None, device="cuda:0", seed=42, beta=None, lr=5.0e-4, rng="device"):
        conf = conf or synth.dtu_conf()
        torch.manual_seed(seed)
        self.model = VolSDFNetwork(conf)
        if beta is not None:
            with torch.no_grad():
                self.model.density.beta.fill_(beta)
        self.model = self.model.to(device).train()
        self.model.rng = rng
        self.loss_fn = VolSDFLoss(**synth.loss_conf())
        self.bucket = GradBucket(self.model.parameters())
        # same update rule as the reference's torch.optim.Adam(lr) (volsdf_train.py:178), all tensors in one launch
        self.opt = Adam(self.model.parameters(), lr=lr)

    def step(self, inp, gt):
        out = self.model(inp)
        lo = self.loss_fn(out, gt)
        self.bucket.zero()
        lo["loss"].backward()
        self.opt.grad_scale = self.bucket.all_reduce_sum()  # the 1/world scale rides in the Adam kernel
        self.opt.step()
        return lo["loss"]
