#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_c20_bench2.json 2> gpurun_out/r2_c20_bench2.err; echo "bench2 exit $?"; tail -n 3 gpurun_out/r2_c20_bench2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c20_bench2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
print(d['stages']['train_step']); print(d['stages']['map_gather'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/train_dp_check.py > gpurun_out/r2_c20_dp_check.txt 2>&1; echo "dp check exit $?"; tail -n 3 gpurun_out/r2_c20_dp_check.txt
timeout 600 python -m pytest tests/test_gpu_train_dp.py -q -m gpu --tb=short 2>&1 | tail -n 3
