#!/bin/bash
# where do the 0.2 ms between the step and the conv pass go?  conf 0.99 (NMS nearly empty) and lane counts
mkdir -p gpurun_out
for args in "--conf 0.5 --lanes 4" "--conf 0.99 --lanes 4" "--conf 0.5 --lanes 2" "--conf 0.5 --lanes 6" "--conf 0.5 --lanes 8"; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra-stages --no-cpu-baseline $args > gpurun_out/r2_c37.json 2> gpurun_out/r2_c37.err; echo "bench $args exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c37.json').read().strip().splitlines()[-1])
print("$args", {k:round(d[k],3) for k in ('value','ms_per_step')}, round(d['e2e']['value']), 'conv sustained', round(d['roofline']['sustained']['ms_per_step_conv'],3), 'burst', round(d['roofline']['ms_per_step_conv'],3), 'nms', round(d['stages']['nms']['ms_per_step'],3), d['clocks']['sm_mhz'])
PY
done
