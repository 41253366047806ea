#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_eval_boxes.py -q -m gpu -x --tb=short > gpurun_out/r2_c3_tests.log 2>&1; echo "conv+model tests exit $?"
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_kernels.py -q -m gpu -x --tb=short > gpurun_out/r2_c3_train.log 2>&1; echo "train tests exit $?"
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c3_layer_times_warm.txt 2>&1; echo "lt exit $?"
timeout 300 python scripts/conv_trace.py > gpurun_out/r2_c3_trace.txt 2>&1; echo "trace exit $?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_c3_bench.json 2> gpurun_out/r2_c3_bench.err; echo "bench exit $?"
tail -n 3 gpurun_out/r2_c3_tests.log gpurun_out/r2_c3_train.log
tail -n 7 gpurun_out/r2_c3_layer_times_warm.txt
cat gpurun_out/r2_c3_bench.json
