#!/bin/bash
# DRAM traffic per conv launch with caches left as the previous launch left them (ncu --cache-control none):
# how much of a residual pair's hand-off (1x1 output -> 3x3 input) actually comes from DRAM
mkdir -p gpurun_out
python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c74_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none --csv --log-file gpurun_out/r2_v10_launches_warm.csv python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c74_ncu.log 2>&1
echo "exit $?"
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r2_c74_train.json 2> gpurun_out/r2_c74_train.err; echo "train bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_c74_train.json').read().strip().splitlines()[-1])
print(d['metric'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['launch_mode'], d['roofline']['frac'])
"
