#!/bin/bash
# probe: images/s vs per-step batch (L2 residency of the early layers vs tile quantisation of the late ones)
mkdir -p gpurun_out
for b in 16 32 128; do
timeout 600 python bench.py --steps 20 --warmup 5 --batch $b --no-extra-stages --no-cpu-baseline > gpurun_out/r2_c31_bench_b$b.json 2> gpurun_out/r2_c31_bench_b$b.err; echo "bench b=$b exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c31_bench_b$b.json').read().strip().splitlines()[-1])
print($b, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['sustained']['ms_per_step_conv'], d['clocks'])
PY
done
timeout 300 python scripts/layer_times.py --batch 16 > gpurun_out/r2_c31_lt_b16.txt 2>&1; tail -8 gpurun_out/r2_c31_lt_b16.txt
