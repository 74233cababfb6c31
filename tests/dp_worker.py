"""Worker of tests/test_gpu_parallel.py (one process per GPU under torchrun, NCCL): the data-parallel step through the REAL
plugin.  Checks, on every rank:
  1. bucket path (p.grad = views of the flat bucket, zeroed, the backward ADDS into them) == plain path (p.grad = None,
     the backward overwrites) for this rank's shard;
  2. all_reduce(SUM)(bucket) / world == mean over shards of the single-rank gradients, every shard recomputed locally
     (semantics of SURVEY section 8e: the reference run once per image, gradients averaged;
     the per-image count normalisation of code/model/networks/loss_wfr.py:44 stays per shard);
  3. after neat_b200.optim.Adam (grad_scale = 1 / world) every rank holds bit-identical parameters.
Prints one JSON line per rank; exit code != 0 on failure."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

from neat_b200 import synth
from neat_b200 import trainer as TR


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    R = int(os.environ.get("DP_RAYS", "256"))
    ts = TR.TrainStep(synth.dtu_conf(), device=dev, seed=42, beta=0.1)
    names = [n for n, p in ts.model.named_parameters() if p.requires_grad]
    params = dict(ts.model.named_parameters())

    def shard_batch(s):
        return TR.to_device(TR.host_batch(R, seed=1 + s, camera_seed=1 + s), dev)

    def run(s):
        inp, gt = shard_batch(s)
        ts.model.seed_draws(100 + s)                    # the device-side draws of shard s, whoever computes it
        lo = ts.loss_fn(ts.model(inp), gt)
        lo["loss"].backward()
        return float(lo["loss"])

    # ---- plain path, every shard (the single-rank references)
    views = {n: params[n].grad for n in names}          # the bucket views TrainStep bound
    ref = None
    mine_plain = None
    losses = []
    for s in range(world):
        for n in names:
            params[n].grad = None
        losses.append(run(s))
        flat = torch.cat([params[n].grad.reshape(-1) if params[n].grad is not None else torch.zeros_like(params[n]).reshape(-1)
                          for n in names])
        ref = flat.clone() if ref is None else ref + flat
        if s == rank:
            mine_plain = flat.clone()
    ref /= world
    # ---- bucket path, this rank's shard
    for n in names:
        params[n].grad = views[n]
    ts.bucket.zero()
    run(rank)
    flat_bucket = torch.cat([params[n].grad.reshape(-1) for n in names])
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    e_bucket = rel(flat_bucket, mine_plain)
    scale = ts.bucket.all_reduce_sum()
    reduced = torch.cat([params[n].grad.reshape(-1) for n in names]) * scale
    e_mean = rel(reduced, ref)
    # ---- optimizer step: identical parameters everywhere
    ts.opt.grad_scale = scale
    ts.opt.step()
    chk = torch.stack([p.detach().double().sum() for p in ts.model.parameters()] +
                      [p.detach().double().abs().sum() for p in ts.model.parameters()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(bool(torch.equal(allc[0], c)) for c in allc)
    ok = e_bucket < 2e-5 and e_mean < 2e-5 and same and scale == 1.0 / world
    print(json.dumps({"rank": rank, "world": world, "rays_per_rank": R, "shard_losses": losses,
                      "bucket_vs_plain_rel_err": e_bucket, "allreduce_vs_mean_of_shards_rel_err": e_mean,
                      "params_identical_after_adam": same, "ok": ok}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
