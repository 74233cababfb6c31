#!/bin/bash
# Retry a short GPU validation call until the pod has a free slot (exit code 3 = nothing charged).
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout 170 -- 'mkdir -p gpurun_out; timeout 110 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu_final.log; timeout 40 python __graft_entry__.py smoke 2>&1 | tail -1' > /root/repo/gpurun_out/final_try.log 2>&1
  rc=$?
  echo "attempt $i rc=$rc $(date)" >> /root/repo/gpurun_out/final_try_attempts.log
  if ! grep -q "status=transient" /root/repo/gpurun_out/final_try.log; then break; fi
  sleep 200
done
