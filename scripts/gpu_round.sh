#!/bin/bash
# One GPU session: parity tests, bench variants (L2 prefetch mask), ncu full captures of the step's MLP kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for pf in 1 0 3; do
  NEAT_L2_PREFETCH=$pf timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_pf$pf.json 2>> gpurun_out/bench.err
done
timeout 600 python bench.py --steps 20 --rays 8192 --no-cpu-baseline > gpurun_out/bench_8192.json 2>> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sdf_query|sdf_render|sdf_bwd|head_fwd|head_bwd|wgrad" --launch-skip 45 --launch-count 15 -o gpurun_out/step_full -f python scripts/profile_step.py 1024 4 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
