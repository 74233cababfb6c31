#!/bin/bash
# quick check: parity tests + the two bench sizes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_1024.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --rays 8192 --no-cpu-baseline > gpurun_out/bench_8192.json 2>> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
