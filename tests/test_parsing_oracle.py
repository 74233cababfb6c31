"""CPU: the finalisation oracle (oracle/parsing_oracle.py) against a brute-force evaluation of the same definition."""
import numpy as np
import torch

from oracle import parsing_oracle as PO


def test_vote_lines_oracle_matches_bruteforce():
    g = torch.Generator().manual_seed(0)
    N, Gn = 200, 17
    gt = torch.rand(Gn, 4, generator=g) * 100
    pick = torch.randint(0, Gn, (N,), generator=g)
    l2 = gt[pick] + torch.randn(N, 4, generator=g)
    l2[::3] = l2[::3][:, [2, 3, 0, 1]]
    l2[::7] += 80
    l3 = torch.randn(N, 2, 3, generator=g)
    p3 = torch.randn(N, 3, generator=g)
    labels, mean, scores, counts = PO.vote_lines(l2, l3, p3, gt, 10.0)
    acc = {}
    for i in range(N):
        for sw in (False, True):
            a = l2[i][[2, 3, 0, 1]] if sw else l2[i]
            d = ((a[None] - gt) ** 2).sum(-1)
            k = int(d.argmin())
            if float(d[k]) < 10.0:
                acc.setdefault(k, []).append((l3[i][[1, 0]] if sw else l3[i], p3[i]))
    assert sorted(acc) == labels.tolist()
    for j, k in enumerate(labels.tolist()):
        m = torch.stack([v[0] for v in acc[k]]).mean(0)
        assert torch.allclose(m, mean[j], atol=1e-6)
        assert counts[j] == len(acc[k])
        dis = [float(torch.linalg.norm(torch.linalg.cross(p - m[0], p - m[1])) / torch.linalg.norm(m[1] - m[0]).clamp_min(1e-6))
               for _, p in acc[k]]
        assert abs(np.mean(dis) - float(scores[j])) < 1e-5


def test_match_endpoints_oracle():
    gj = torch.tensor([[0.0, 0, 0], [1, 0, 0], [5, 5, 5]])
    lines = torch.tensor([[[0.01, 0, 0], [1.0, 0.02, 0]]])
    assert PO.match_endpoints(gj, lines, 0.05) == [(0, 0), (1, 1)]
