#!/bin/bash
for sz in 64 416 608; do timeout 120 python scripts/stem_debug.py $sz 0 2>&1 | tail -1 | cut -c1-200; done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_conv.py -q -m gpu --tb=short -x > gpurun_out/r2_c18_tests.log 2>&1; echo "tests exit $?"; tail -n 12 gpurun_out/r2_c18_tests.log | cut -c1-300
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c18_lt.txt 2>&1
head -3 gpurun_out/r2_c18_lt.txt; grep -E "layers.16|layers.23 " gpurun_out/r2_c18_lt.txt; tail -n 6 gpurun_out/r2_c18_lt.txt
