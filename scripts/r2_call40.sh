#!/bin/bash
mkdir -p gpurun_out
for v in "" f b fb; do
echo "== YOLO_B200_TRAIN_NO_PDL='$v'"
YOLO_B200_TRAIN_NO_PDL=$v timeout 300 python scripts/train_times.py 32 416 mish 10 2>&1 | tail -3 | head -2
YOLO_B200_TRAIN_NO_PDL=$v timeout 300 python scripts/train_times.py 32 416 leaky_relu 10 2>&1 | tail -3 | head -2
done
