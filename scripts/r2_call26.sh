#!/bin/bash
mkdir -p gpurun_out
for n in 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_c26_bench$n.json 2> gpurun_out/r2_c26_bench$n.err; echo "bench$n exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c26_bench$n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['sustained']['frac'])
t=d['stages']['train_step']; print({k:t.get(k) for k in ('ms_per_step','images_per_sec','ms_per_step_without_allreduce','exposed_comm_ms','replicas_bit_identical')}); print(d['stages']['map_gather'])
PY
done
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 | cut -c1-300
