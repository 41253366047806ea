#!/bin/bash
# Runs the GPU tests with process isolation (a kernel trap poisons the CUDA context) and keeps the
# logs under gpurun_out/.  Usage (on the GPU box): bash scripts/gpu_check.sh [isolate|fast] [files...]
mode=${1:-isolate}; shift
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
files=${@:-$(ls tests/test_gpu_*.py)}
rc=0
for f in $files; do
  name=$(basename $f .py)
  if [ "$mode" = "isolate" ]; then
    ids=$(python -m pytest $f --collect-only -q -m gpu 2>/dev/null | grep "::")
    : > gpurun_out/$name.log
    for id in $ids; do
      timeout 300 python -m pytest "$id" -q -m gpu -x --tb=short -s >> gpurun_out/$name.log 2>&1
      r=$?
      echo "$id exit $r" | tee -a gpurun_out/summary.txt
      [ $r -ne 0 ] && rc=1
    done
  else
    timeout 900 python -m pytest $f -q -m gpu -x --tb=short -s > gpurun_out/$name.log 2>&1
    r=$?
    echo "$name exit $r" | tee -a gpurun_out/summary.txt
    [ $r -ne 0 ] && rc=1
  fi
done
grep -E "Error|error|assert|FAILED" gpurun_out/*.log | head -60
exit $rc
