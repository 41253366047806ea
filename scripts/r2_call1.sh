#!/bin/bash
# round 2, GPU call 1: parity of the PDL / tail-split changes, then timings with each switched off
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -x --tb=short > gpurun_out/r2_c1_conv.log 2>&1; echo "conv tests exit $?"
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x --tb=short > gpurun_out/r2_c1_model.log 2>&1; echo "model tests exit $?"
timeout 900 python -m pytest tests -q -m gpu --tb=short --deselect tests/test_gpu_conv.py --deselect tests/test_gpu_model.py > gpurun_out/r2_c1_rest.log 2>&1; echo "other tests exit $?"
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c1_layer_times_warm.txt 2>&1; echo "lt exit $?"
timeout 300 python scripts/layer_times.py --warm --no-pdl --no-split > gpurun_out/r2_c1_layer_times_warm_nopdl_nosplit.txt 2>&1
timeout 300 python scripts/layer_times.py --warm --no-split > gpurun_out/r2_c1_layer_times_warm_nosplit.txt 2>&1
timeout 300 python scripts/conv_trace.py > gpurun_out/r2_c1_trace.txt 2>&1; echo "trace exit $?"
timeout 600 python bench.py > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r2_c1_conv.log gpurun_out/r2_c1_model.log gpurun_out/r2_c1_rest.log
tail -8 gpurun_out/r2_c1_layer_times_warm.txt; tail -8 gpurun_out/r2_c1_layer_times_warm_nopdl_nosplit.txt | head -3
cat gpurun_out/r2_c1_bench.json
