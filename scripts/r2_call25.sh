#!/bin/bash
python scripts/box_index.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_train_kernels.py -q -m gpu --tb=short > gpurun_out/r2_c25_tests.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/r2_c25_tests.log | cut -c1-200
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c25_lt.txt 2>&1; tail -n 6 gpurun_out/r2_c25_lt.txt
grep -E "layers.6.layers.0|layers.8.layers.0|layers.10.layers.0|layers.2.layers|layers.4.layers.0" gpurun_out/r2_c25_lt.txt
