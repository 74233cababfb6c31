#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/b_a.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --rays 8192 --no-cpu-baseline > gpurun_out/b_a_8192.json 2>> gpurun_out/bench.err
