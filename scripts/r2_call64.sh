#!/bin/bash
# final validation of the round: every GPU test, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2_c64_tests.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/r2_c64_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_c64_bench.json 2> gpurun_out/r2_c64_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c64_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['sustained']['ms_per_step_conv'], d['clocks'])
print(d['stages']['nms']['ms_per_step'], d['stages']['e2e_uint8'].get('value'), d['stages']['train_step']); print(d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
