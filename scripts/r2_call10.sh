#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_c10_bench.json 2> gpurun_out/r2_c10_bench.err; echo "bench exit $?"; tail -n 5 gpurun_out/r2_c10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c10_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step')}, d['e2e'], d['roofline']['ms_per_step_conv'], d['roofline']['frac'], d['clocks'])
print(d['timing']); print(d['stages']['nms']); print(d['stages'].get('map_gather')); print(d['stages'].get('train_step')); print(d.get('cpu_baseline'))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_c10_ref.json 2> gpurun_out/r2_c10_ref.err; echo "ref exit $?"; cat gpurun_out/r2_c10_ref.json | cut -c1-600
