#!/bin/bash
mkdir -p gpurun_out
python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c66_plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_nms_ -c 2 -o gpurun_out/r2_v10_nms_full -f python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c66_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2_v10_nms_full.ncu-rep
