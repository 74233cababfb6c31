#!/bin/bash
# Round-2 evidence: compute-sanitizer logs, ncu launch list, ncu --set full of the tile-MLP kernels, HBM metrics of the per-ray /
# dataset kernels.  Outputs under gpurun_out/ (summarised into profiles/ by scripts/make_profiles_r02.py).
set -x
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size"
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_step.py 33 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|train step|eval forward|exit" gpurun_out/sanitizer_$tool.log | tail -5
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 1024 4 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sdf_query|sdf_render|sdf_bwd|head_fwd|head_bwd|wgrad" --launch-skip 30 --launch-count 15 -o gpurun_out/step_full -f python scripts/profile_step.py 1024 4 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"sampler_|composite_|db_|loss_|camera_rays|line_geometry|pose_inverse|weight_norm|adam_|train_draws|junction_|gemm_f32|colsum|pack_|project_" --csv --log-file gpurun_out/perray_metrics.csv python scripts/profile_step.py 1024 3 0.01 > gpurun_out/ncu_perray.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"encodels|point_line_attraction|mask_|sample_pixels|line_vote|line_visibility|line_junction" --csv --log-file gpurun_out/aux_metrics.csv python scripts/profile_aux.py > gpurun_out/ncu_aux.log 2>&1
tail -2 gpurun_out/ncu_aux.log
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
# the benchmarked path (FusedTrainStep: two CUDA-graph replays per step): launch list of the last step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fused_launches.csv python scripts/profile_fused.py 1024 5 > gpurun_out/fused_ncu.log 2>&1
tail -1 gpurun_out/fused_ncu.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 60 python scripts/sanitize_step.py 64 > gpurun_out/sanitizer_initcheck.log 2>&1
grep -E "ERROR SUMMARY" gpurun_out/sanitizer_initcheck.log
# the driver's two arms
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> /dev/null
