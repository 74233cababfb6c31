#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for pf in 1 2 3; do
  NEAT_L2_PREFETCH=$pf timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_pf$pf.json 2>> gpurun_out/bench.err
done
NEAT_L2_PREFETCH=2 timeout 600 python bench.py --steps 20 --rays 8192 --no-cpu-baseline > gpurun_out/bench_8192_pf2.json 2>> gpurun_out/bench.err
timeout 300 python scripts/time_query.py 1024 20 0 > gpurun_out/time_query.log 2>&1
timeout 300 python scripts/time_query.py 1024 20 1 >> gpurun_out/time_query.log 2>&1
cat gpurun_out/time_query.log
