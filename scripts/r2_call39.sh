#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train.py -q -m gpu --tb=short -x > gpurun_out/r2_c39_tests.log 2>&1; echo "tests exit $?"; tail -n 6 gpurun_out/r2_c39_tests.log | cut -c1-300
timeout 300 python scripts/train_times.py 32 416 mish 10 > gpurun_out/r2_c39_train_times.txt 2>&1; tail -4 gpurun_out/r2_c39_train_times.txt
timeout 300 python scripts/train_times.py 32 416 leaky_relu 10 > gpurun_out/r2_c39_train_times_leaky.txt 2>&1; tail -3 gpurun_out/r2_c39_train_times_leaky.txt
