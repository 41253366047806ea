#!/bin/bash
# restored-container sanity: all GPU tests, then the bench with and without weight-tile multicast (sustained regime A/B)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/r2_c30_tests.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/r2_c30_tests.log | cut -c1-300
for mc in 0 2; do
YOLO_B200_MC=$mc timeout 600 python bench.py --steps 20 --warmup 5 --no-extra-stages > gpurun_out/r2_c30_bench_mc$mc.json 2> gpurun_out/r2_c30_bench_mc$mc.err; echo "bench mc=$mc exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c30_bench_mc$mc.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['sustained']['ms_per_step_conv'], d['clocks'])
print(d['timing'])
PY
done
