#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q -m gpu --tb=short > gpurun_out/r2_c12_tests.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/r2_c12_tests.log
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c12_lt_row.txt 2>&1
timeout 300 python scripts/layer_times.py --warm --no-row > gpurun_out/r2_c12_lt_norow.txt 2>&1
paste <(awk '{print $1, $(NF-1)}' gpurun_out/r2_c12_lt_row.txt) <(awk '{print $(NF-1)}' gpurun_out/r2_c12_lt_norow.txt) | head -12
tail -n 6 gpurun_out/r2_c12_lt_row.txt; tail -n 6 gpurun_out/r2_c12_lt_norow.txt | head -3
timeout 300 python scripts/conv_trace.py --layers layers.1,layers.2.layers.0.1,layers.4.layers.0.1 2>&1 | cut -c1-420
