#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/r2_c41_tests.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/r2_c41_tests.log | cut -c1-300
timeout 300 python scripts/train_times.py 32 416 mish 10 > gpurun_out/r2_c41_train_times.txt 2>&1; tail -4 gpurun_out/r2_c41_train_times.txt
timeout 300 python scripts/train_times.py 32 416 leaky_relu 10 > gpurun_out/r2_c41_train_times_leaky.txt 2>&1; tail -3 gpurun_out/r2_c41_train_times_leaky.txt
