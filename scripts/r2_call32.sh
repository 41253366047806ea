#!/bin/bash
# training step: per-launch times with WARM caches (ncu --cache-control none) to see which BatchNorm/activation launches
# sit far from their HBM bound in situ; plus the plain step time
mkdir -p gpurun_out
timeout 300 python scripts/train_times.py 32 416 mish 10 > gpurun_out/r2_c32_train_times.txt 2>&1; tail -4 gpurun_out/r2_c32_train_times.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_c32_train_launches_warm.csv python scripts/train_for_ncu.py > gpurun_out/r2_c32_ncu.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/r2_c32_train_launches_warm.csv | head -24 | cut -c1-170
