#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1 | tee gpurun_out/r2_c8_box.txt
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --tb=short > gpurun_out/r2_c8_train.log 2>&1; echo "train tests exit $?"; tail -n 25 gpurun_out/r2_c8_train.log
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c8_layer_times_warm.txt 2>&1; tail -n 7 gpurun_out/r2_c8_layer_times_warm.txt
python scripts/box_index.py 2>&1 | tail -1
