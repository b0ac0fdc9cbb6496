#!/bin/bash
# round 2, GPU call 14 (8 GPUs): bench.py under torchrun at N=8 and N=4 (BASELINE configs[3], strong scaling)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n > gpurun_out/r2_scale_${n}gpu.json 2> gpurun_out/r2_scale_${n}gpu.err; echo "bench N=$n rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2_scale_${n}gpu.json'))
print('N=$n', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3),'streams',d['timed']['streams'],'multi',{k:(round(v,2) if isinstance(v,float) else v) for k,v in (d.get('one_process_multi_gpu') or {}).items() if k in ('value','p50_batch_latency_ms','p99_batch_latency_ms','solved_fraction')},'weak',round((d.get('weak_scaling') or {}).get('value',0)/1e6,2), 'per-rank ms', [round(x,3) for x in d['per_rank_ms_per_step']])
P
done
tail -3 gpurun_out/r2_scale_8gpu.err
