#!/bin/bash
# multi-GPU records: bench (inference + train_step + map_gather stages) and the configs[4] NMS sweep on N ranks
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_c35_bench_${N}gpu.json 2> gpurun_out/r2_c35_bench_${N}gpu.err; echo "bench$N exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c35_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['stages']['e2e_uint8'].get('value'), d['roofline']['frac'], d['clocks'])
t=d['stages']['train_step']; print({k:t.get(k) for k in ('ms_per_step','images_per_sec','ms_per_step_without_allreduce','exposed_comm_ms','replicas_bit_identical','nccl_kernels_one_step')}); print(d['stages']['map_gather'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 scripts/nms_sweep.py > gpurun_out/r2_c35_nms_sweep_${N}gpu.txt 2> gpurun_out/r2_c35_nms_sweep_${N}gpu.err; echo "sweep$N exit $?"; head -30 gpurun_out/r2_c35_nms_sweep_${N}gpu.txt
