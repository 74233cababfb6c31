"""Workload for the ncu captures of the dataset-side / finalisation kernels at DTU image size:
encodels, point_line_attraction, mask_compact, sample_pixels, line_vote, line_visibility, line_junction_graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from neat_b200 import attraction, dataset, parsing, synth
H, W = 1200, 1600
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
verts, edges, ew = synth.make_wireframe(3, 200, 300, W, H)
lines = torch.from_numpy(np.concatenate([verts[edges[:, 0]], verts[edges[:, 1]], ew[:, None]], -1).astype(np.float32)).to(dev)
for _ in range(2):
    mp, label, tmap = attraction.encodels(lines[:, :4].contiguous(), H, W, H, W, lines.shape[0])
    mask, labels, proj = attraction.compute_point_line_attraction(lines, (H, W), 5.0)
scene = dataset.DeviceScene((H, W), device=dev, rng="device", seed=1)
rgb = torch.from_numpy(rs.rand(H * W, 3).astype(np.float32))
b = synth.make_batch(8, seed=1)
scene.add_image(rgb, lines, b["intrinsics"][0], b["pose"][0], wireframe=None, tables=(mask, labels, proj))
scene.change_sampling_idx(1024)
for _ in range(3):
    idx, sample, gt = scene[0]
N, G = 65536, 300
g = torch.Generator().manual_seed(0)
gt_l = lines[:, :4].cpu()
pick = torch.randint(0, G, (N,), generator=g)
l2 = (gt_l[pick] + torch.randn(N, 4, generator=g) * 1.5).to(dev)
l3 = torch.randn(N, 2, 3, generator=g).to(dev)
p3 = l3.mean(1) + 0.01
for _ in range(2):
    parsing.vote_lines(l2, l3, p3, gt_l.to(dev), 10.0)
    parsing.line_visibility(l3[:4096].contiguous(), torch.from_numpy(b["pose"][0]).to(dev), torch.from_numpy(b["intrinsics"][0]).to(dev),
                            gt_l.to(dev), 25.0)
    parsing.wireframe_from_lines_and_junctions(l3[:4096].contiguous(), torch.randn(1024, 3, generator=g).to(dev))
torch.cuda.synchronize()
print("aux workload done: masked pixels", int(scene.images[0].masked.numel()))
