#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_train_kernels.py -q -m gpu --tb=short -x > gpurun_out/r2_c57_tests.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/r2_c57_tests.log | cut -c1-300
timeout 300 python scripts/layer_times.py > gpurun_out/r2_c57_lt.txt 2>&1; grep -E "^layers.(0|1|3|5|7) |layers.2.layers.0|layers.4.layers.0|layers.6.layers.0|layers.8.layers.0|layers.10.layers.0" gpurun_out/r2_c57_lt.txt | cut -c1-120; tail -6 gpurun_out/r2_c57_lt.txt
