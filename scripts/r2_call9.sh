#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1 | tee gpurun_out/r2_c9_box.txt
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_train_kernels.py tests/test_gpu_eval_boxes.py -q -m gpu --tb=short > gpurun_out/r2_c9_tests.log 2>&1; echo "tests exit $?"; tail -n 6 gpurun_out/r2_c9_tests.log
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c9_layer_times_warm.txt 2>&1; tail -n 7 gpurun_out/r2_c9_layer_times_warm.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_c9_bench.json 2> gpurun_out/r2_c9_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_c9_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['ms_per_step_conv'], d['roofline']['frac_of_burst_peak'], d['clocks'])
PY
