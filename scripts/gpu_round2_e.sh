#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
tail -3 gpurun_out/bench_fused.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fused.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","gpu_launches","kernel_ms_per_step","step_minus_big_kernels_ms","host_junction_block","plugin_path"):
    print(k, d.get(k))
print(d["e2e"])
for c in d["configs"]:
    print(c["name"][:60], round(c["value"]), round(c["ms_per_step"],3), c.get("sampler_k_mean"), c.get("plugin_path_ms_per_step"))
PY
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_step.py 33 > gpurun_out/sanitizer_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|train step|eval forward" gpurun_out/sanitizer_racecheck.log
