#!/bin/bash
# strong scaling of configs[1]: 64 images in total per step, i.e. batch 64/N per GPU
N=${1:-8}
mkdir -p gpurun_out
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --batch $((64 / N)) --steps 20 --warmup 5 --no-extra-stages > gpurun_out/r2_c78_strong_${N}gpu.json 2> gpurun_out/r2_c78_strong_${N}gpu.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c78_strong_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config'], d['e2e']['value'], d['roofline']['frac'])
PY
