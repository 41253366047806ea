#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c34_bench.json 2> gpurun_out/r2_c34_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c34_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['roofline']['frac'], d['roofline']['sustained']['ms_per_step_conv'], d['clocks'])
print(d['stages']['e2e_uint8']); print(d['stages']['train_step']); print(d['cpu_baseline'])
PY
tail -3 gpurun_out/r2_c34_bench.err
