#!/bin/bash
# Final evidence run of the round: parity tests, smoke, bench lines, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_1024_20.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --rays 8192 --no-cpu-baseline > gpurun_out/bench_8192.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --beta 0.01 --no-cpu-baseline > gpurun_out/bench_beta001.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --mode eval --rays 65536 --steps 10 > gpurun_out/bench_eval_65536.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 1024 4 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sdf_query|sdf_render|sdf_bwd|head_fwd|head_bwd|wgrad" --launch-skip 45 --launch-count 15 -o gpurun_out/step_full -f python scripts/profile_step.py 1024 4 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/bench.err
