#!/bin/bash
mkdir -p gpurun_out
for id in "tests/test_gpu_conv.py::test_row_window_mode[shifted_start_plus_base_offset]" "tests/test_gpu_conv.py::test_row_window_mode[shifted_start]" "tests/test_gpu_conv.py::test_rectangular_geometry_pair_folded_stride2[2-16-24]" "tests/test_gpu_conv.py::test_rectangular_geometry_pair_folded_stride2[1-8-256]"; do
  echo "=== $id"
  timeout 300 python -m pytest "$id" -q -m gpu --tb=short -x 2>&1 | grep -E "passed|failed|Error|error|assert|bad " | head -8 | cut -c1-500
done
