#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -k "row_window or rectangular" > gpurun_out/r2_c11_row.log 2>&1; echo "row tests exit $?"; tail -n 40 gpurun_out/r2_c11_row.log | cut -c1-400
