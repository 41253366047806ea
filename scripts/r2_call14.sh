#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_train_kernels.py tests/test_gpu_train.py -q -m gpu --tb=short > gpurun_out/r2_c17_tests.log 2>&1; echo "tests exit $?"; tail -n 25 gpurun_out/r2_c17_tests.log | cut -c1-300
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c17_lt.txt 2>&1
paste <(awk '{print $1, $(NF-1)}' gpurun_out/r2_c17_lt.txt) | head -14
tail -n 6 gpurun_out/r2_c17_lt.txt
grep -E "layers.16|layers.23 " gpurun_out/r2_c17_lt.txt
