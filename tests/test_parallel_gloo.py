"""CPU, world_size 2 over gloo: the data-parallel gradient bucket (one all-reduce per step) gives every rank
the mean of the per-rank gradients, and rides extra loss scalars in the same message."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neat_b200.parallel import GradBucket


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    bucket = GradBucket(net.parameters(), extra=2)
    x = torch.full((4, 5), float(rank + 1))
    bucket.zero()
    loss = net(x).pow(2).sum()
    loss.backward()
    # grads must have accumulated INTO the flat buffer (views), not replaced it
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
    local = bucket.flat[:bucket.n].clone()
    bucket.scalars()[0] = float(loss)
    bucket.scalars()[1] = 1.0
    summed = bucket.flat.clone()
    bucket.all_reduce_mean()
    # the SUM-only variant (the 1/world scale is applied inside the Adam kernel, neat_b200.optim.Adam.grad_scale)
    keep = bucket.flat.clone()
    bucket.flat.copy_(summed)
    scale = bucket.all_reduce_sum()
    assert abs(scale - 1.0 / world) < 1e-12
    assert torch.allclose(bucket.flat * scale, keep, atol=1e-6)
    bucket.flat.copy_(keep)
    q.put((rank, local, bucket.flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mean = (res[0][1] + res[1][1]) / 2
    for r, _, flat in res:
        assert torch.allclose(flat[:mean.numel()], mean, atol=1e-6)
        assert abs(float(flat[-1]) - 1.0) < 1e-6          # mean of the two "count" scalars
    assert torch.equal(res[0][2], res[1][2])


def test_bucket_single_process_is_noop():
    net = torch.nn.Linear(3, 2)
    b = GradBucket(net.parameters())
    b.zero()
    net(torch.ones(1, 3)).sum().backward()
    before = b.flat.clone()
    b.all_reduce_mean()
    assert torch.equal(before, b.flat)
    assert b.all_reduce_sum() == 1.0 and torch.equal(before, b.flat)
