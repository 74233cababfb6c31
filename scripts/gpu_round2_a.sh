#!/bin/bash
# Round-2 first GPU pass: parity at the benchmarked sizes, measured error tables, bench lines.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt; free -g | head -2 >> gpurun_out/smi.txt
timeout 900 python scripts/measure_parity.py > gpurun_out/measure_parity.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json
