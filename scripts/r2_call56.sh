#!/bin/bash
# ncu --set full on the fused stem kernel (first conv launch of a step) and on one decode-fused head conv
mkdir -p gpurun_out
python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c56_plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_v2 -c 1 -o gpurun_out/r2_v9_stem_full -f python scripts/step_for_ncu.py 64 416 0.5 > gpurun_out/r2_c56_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2_v9_stem_full.ncu-rep
