#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_conv.py tests/test_gpu_eval_boxes.py tests/test_gpu_sort_nms.py -q -m gpu --tb=short > gpurun_out/r2_c21_tests.log 2>&1; echo "tests exit $?"; tail -n 15 gpurun_out/r2_c21_tests.log | cut -c1-300
timeout 300 python scripts/layer_times.py --warm > gpurun_out/r2_c21_lt.txt 2>&1; tail -n 6 gpurun_out/r2_c21_lt.txt
timeout 900 python bench.py --no-cpu-baseline --no-extra-stages > gpurun_out/r2_c21_bench.json 2> gpurun_out/r2_c21_bench.err; echo "bench exit $?"; tail -n 3 gpurun_out/r2_c21_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c21_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['clocks'])
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','ms_per_step_conv')}, r['sustained']['frac'], r['sustained']['ms_per_step_conv'])
print(d['stages']['decode']['in_timed_step'], d['stages']['nms']['ms_per_step'])
PY
