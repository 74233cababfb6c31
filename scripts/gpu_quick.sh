#!/bin/bash
set -x
mkdir -p gpurun_out
for s in 16 24 32 48; do NEAT_WGRAD_SPLIT=$s timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/b_s$s.json 2>> gpurun_out/bench.err; done
for s in 32 64 96; do NEAT_WGRAD_SPLIT=$s timeout 600 python bench.py --steps 10 --rays 8192 --no-cpu-baseline > gpurun_out/b_s${s}_8192.json 2>> gpurun_out/bench.err; done
