#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/one_layer.py 64 104 64 128 3 1 0 0 1"
$CMD > gpurun_out/r2_c13_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_conv_v2 -s 2 -c 2 -o gpurun_out/r2_c13_row -f $CMD > gpurun_out/r2_c13_ncu.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/r2_c13_ncu.log; ls -la gpurun_out/*.ncu-rep
