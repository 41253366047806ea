#!/bin/bash
# N-rank bench with a HARD time limit (a teardown hang must not burn the GPU budget)
N=${1:-2}
mkdir -p gpurun_out
t0=$(date +%s)
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_c54_bench_${N}gpu.json 2> gpurun_out/r2_c54_bench_${N}gpu.err; echo "bench$N exit $? after $(( $(date +%s) - t0 )) s"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c54_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['stages']['e2e_uint8'].get('value'), d['roofline']['frac'])
t=d['stages']['train_step']; print({k:t.get(k) for k in ('ms_per_step','ms_per_step_eager','launch_mode','images_per_sec','ms_per_step_without_allreduce','exposed_comm_ms','replicas_bit_identical')})
PY
tail -3 gpurun_out/r2_c54_bench_${N}gpu.err | cut -c1-200
