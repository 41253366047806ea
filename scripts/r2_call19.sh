#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2_c19_tests.log 2>&1; echo "tests exit $?"; tail -n 6 gpurun_out/r2_c19_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_c19_bench.json 2> gpurun_out/r2_c19_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c19_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['clocks'])
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','ms_per_step_conv')}, r['sustained'])
print(d['timing']['ms_per_conv_pass_burst'])
print(d['stages']['train_step']); print(d['stages']['map_gather']); print(d['cpu_baseline'])
PY
