#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sort_nms.py tests/test_gpu_model.py tests/test_gpu_eval_boxes.py tests/test_gpu_decode_iou_map.py -q -m gpu --tb=short -x > gpurun_out/r2_c46_tests.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/r2_c46_tests.log | cut -c1-300
python scripts/nms_segments_stats.py 64 416 0.5 2>&1 | tail -1; python scripts/nms_segments_stats.py 32 608 0.01 2>&1 | tail -1
python scripts/nms_sweep.py > gpurun_out/r2_c46_sweep.txt 2>&1; head -24 gpurun_out/r2_c46_sweep.txt
