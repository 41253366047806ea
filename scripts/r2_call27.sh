#!/bin/bash
python scripts/box_index.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -x -k "multicast" 2>&1 | tail -n 12 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q -m gpu --tb=short -x > gpurun_out/r2_c27_tests.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/r2_c27_tests.log | cut -c1-300
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c27_lt_mc.txt 2>&1
timeout 300 python scripts/layer_times.py --warm --no-mc > gpurun_out/r2_c27_lt_nomc.txt 2>&1
paste <(awk '{print $1, $(NF-1)}' gpurun_out/r2_c27_lt_mc.txt) <(awk '{print $(NF-1)}' gpurun_out/r2_c27_lt_nomc.txt) | grep -E "layers.(5|7|9|12|19|26) |layers.(6|8|10).layers.[01].1|pred_block.0"
tail -n 6 gpurun_out/r2_c27_lt_mc.txt; tail -n 6 gpurun_out/r2_c27_lt_nomc.txt | head -3
