#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train.py -q -m gpu --tb=short -x > gpurun_out/r2_c42_tests.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/r2_c42_tests.log | cut -c1-300
timeout 300 python scripts/train_times.py 32 416 mish 10 > gpurun_out/r2_c42_train_times.txt 2>&1; tail -4 gpurun_out/r2_c42_train_times.txt
timeout 300 python scripts/train_times.py 32 416 leaky_relu 10 > gpurun_out/r2_c42_train_times_leaky.txt 2>&1; tail -3 gpurun_out/r2_c42_train_times_leaky.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_c42_train_launches_warm.csv python scripts/train_for_ncu.py > gpurun_out/r2_c42_ncu.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/r2_c42_train_launches_warm.csv 2>/dev/null | head -12 | cut -c1-170
