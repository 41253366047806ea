#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_train_kernels.py tests/test_gpu_train.py -q -m gpu --tb=short > gpurun_out/r2_c15_tests.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/r2_c15_tests.log
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c15_lt.txt 2>&1
paste <(awk '{print $1, $(NF-1)}' gpurun_out/r2_c15_lt.txt) | head -14
tail -n 6 gpurun_out/r2_c15_lt.txt
CMD="python scripts/one_layer.py 64 104 64 128 3 1 0 0 1"
$CMD > gpurun_out/r2_c15_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_conv_v2 -s 2 -c 1 -o gpurun_out/r2_c15_row -f $CMD > gpurun_out/r2_c15_ncu.log 2>&1
echo "ncu exit $?"
