"""Data parallelism on real GPUs (SURVEY section 4 "Distributed", VERDICT r01 missing #4): 2 ranks under torchrun over NCCL,
the real plugin on every rank.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box, where bench.py's own
`dp_check` (printed in every multi-GPU bench line) covers the same property."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_allreduce_equals_mean_of_shard_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 2, (r.stdout[-2000:], r.stderr[-2000:])
    for l in lines:
        assert l["ok"], l
        assert l["bucket_vs_plain_rel_err"] < 2e-5 and l["allreduce_vs_mean_of_shards_rel_err"] < 2e-5
        assert l["params_identical_after_adam"]
