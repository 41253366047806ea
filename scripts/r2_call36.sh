#!/bin/bash
mkdir -p gpurun_out
python scripts/nms_segments_stats.py 64 416 0.5 2>&1 | tail -6
python scripts/nms_segments_stats.py 32 608 0.01 2>&1 | tail -6
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_c36_launches_608.csv python scripts/step_for_ncu.py 32 608 0.01 > gpurun_out/r2_c36_ncu608.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_c36_launches_608.csv 2>/dev/null | grep -v conv | head -20 | cut -c1-120
