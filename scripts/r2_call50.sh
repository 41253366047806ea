#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu --tb=short -x -k "graphed or dropin or momentum" > gpurun_out/r2_c50_tests.log 2>&1; echo "tests exit $?"; tail -n 25 gpurun_out/r2_c50_tests.log | cut -c1-300
timeout 300 python scripts/train_times.py 32 416 mish 10 2>&1 | tail -4
