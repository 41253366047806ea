#!/bin/bash
mkdir -p gpurun_out
python scripts/box_index.py 2>&1 | tail -1
timeout 900 python bench.py --no-cpu-baseline --no-extra-stages > gpurun_out/r2_c16_bench.json 2> gpurun_out/r2_c16_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c16_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['ms_per_step_conv'], d['roofline']['frac'], d['clocks'])
print(d['timing'])
PY
timeout 900 python bench.py --no-cpu-baseline --no-extra-stages --lanes 1 > gpurun_out/r2_c16_bench_l1.json 2> gpurun_out/r2_c16_bench_l1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c16_bench_l1.json').read().strip().splitlines()[-1])
print("lanes1", {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['ms_per_step_conv'], d['roofline']['frac'], d['clocks'])
PY
