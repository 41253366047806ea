#!/bin/bash
timeout 300 python scripts/conv_trace.py > gpurun_out/r2_c24_trace.txt 2>&1; cut -c1-200 gpurun_out/r2_c24_trace.txt | head -80
