#!/bin/bash
timeout 300 python scripts/layer_times.py > gpurun_out/r2_c22_lt.txt 2>&1; grep -E "decode" gpurun_out/r2_c22_lt.txt
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_eval_boxes.py -q -m gpu --tb=short 2>&1 | tail -n 2
python scripts/decode_ab.py > gpurun_out/r2_c22_ab.txt 2>&1; tail -4 gpurun_out/r2_c22_ab.txt
