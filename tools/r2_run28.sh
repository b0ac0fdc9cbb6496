#!/bin/bash
# round 2, GPU call 28+: bench.py --gpus $1 under torchrun with the final build of the round (one 65 536-problem batch per
# step sharded by index, scaling: strong)
n=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n > gpurun_out/r2_scale27_${n}gpu.json 2> gpurun_out/r2_scale27_${n}gpu.err; echo "bench N=$n rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/r2_scale27_${n}gpu.json'))
print('N=$n', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), 'streams', d['timed']['streams'], 'multi', {k: (round(v,2) if isinstance(v,float) else v) for k,v in (d.get('one_process_multi_gpu') or {}).items() if k in ('value','p50_batch_latency_ms','p99_batch_latency_ms','solved_fraction')}, 'weak', round((d.get('weak_scaling') or {}).get('value',0)/1e6,2), 'per-rank ms', [round(x,3) for x in d.get('per_rank_ms_per_step',[])])
P
tail -2 gpurun_out/r2_scale27_${n}gpu.err
