#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python scripts/measure_parity.py > gpurun_out/measure_parity.log 2>&1
grep -v -i Warn gpurun_out/measure_parity.log | grep "^eval" | cut -c1-900
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cut -c1-400 gpurun_out/bench_default.json
