#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/b_p0.json 2>> gpurun_out/bench.err
NEAT_L2_PERSIST=50 timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/b_p50.json 2> gpurun_out/p50.err
NEAT_L2_PERSIST=100 timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/b_p100.json 2>> gpurun_out/bench.err
grep "neat\]" gpurun_out/p50.err
