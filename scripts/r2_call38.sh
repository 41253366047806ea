#!/bin/bash
# ncu evidence for the final round-2 kernels: launch list with DRAM bytes of one 416 step, and --set full on the big 3x3 conv
mkdir -p gpurun_out
python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c38_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_v7_launches.csv python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c38_ncu1.log 2>&1
echo "launch list exit $?"
python scripts/summarize_launches.py gpurun_out/r2_v7_launches.csv 2>/dev/null | head -14 | cut -c1-170
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_v2 -s 30 -c 3 -o gpurun_out/r2_v7_conv_full -f python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c38_ncu2.log 2>&1
echo "full exit $?"; ls -la gpurun_out/*.ncu-rep
