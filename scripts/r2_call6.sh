#!/bin/bash
mkdir -p gpurun_out
L=layers.4.layers.0.1,layers.6.layers.0.0,layers.6.layers.0.1,layers.8.layers.0.0,layers.8.layers.0.1,layers.10.layers.0.1
for b in 0 1 2 3; do echo "box $b"; timeout 300 python scripts/conv_trace.py --box $b --layers $L 2>&1 | grep -o "^layers[^ ]*\|epi box.*tile [0-9. ]*"; done > gpurun_out/r2_c5_trace.txt 2>&1
cat gpurun_out/r2_c5_trace.txt
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_eval_boxes.py -q -m gpu --tb=short > gpurun_out/r2_c6_tests.log 2>&1; echo "tests exit $?"
tail -n 40 gpurun_out/r2_c6_tests.log
