#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python scripts/measure_parity.py > gpurun_out/measure_parity.log 2>&1
grep -v Warn gpurun_out/measure_parity.log | cut -c1-1500
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cut -c1-600 gpurun_out/bench_quick.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d["kernel_ms_per_step"], d["step_minus_big_kernels_ms"])
PY
