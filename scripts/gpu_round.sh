#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_1024.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --mode eval --rays 65536 --steps 10 > gpurun_out/bench_eval_65536.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --mode eval --rays 2048 --steps 20 > gpurun_out/bench_eval_2048.json 2>> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
cat gpurun_out/bench_eval_65536.json gpurun_out/bench_eval_2048.json
