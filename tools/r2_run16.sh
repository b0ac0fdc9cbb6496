#!/bin/bash
# round 2, GPU call 16: default bench line with the horizon sweep key; the 8-GPU shard size with 32 hardware queues
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench16.json'))
print(round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3),'lone',round(d['lone_caller']['value']/1e6,3),'p99',round(d['e2e']['p99_batch_latency_ms'],2))
for r in d['horizon_sweep']['rows']: print(r)
P
for c in 8 32; do
CUDA_DEVICE_MAX_CONNECTIONS=$c python bench.py --batch 8192 --no-cpu-baseline --no-sweep --latency-reps 20 > gpurun_out/r2_b8192_c$c.json 2>> gpurun_out/r2_bench16.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b8192_c$c.json')); print('B=8192 connections $c:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), 'streams', d['timed']['streams'])"
done
CUDA_DEVICE_MAX_CONNECTIONS=8 python bench.py --no-cpu-baseline --no-sweep --latency-reps 20 > gpurun_out/r2_b64k_c8.json 2>> gpurun_out/r2_bench16.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b64k_c8.json')); print('B=65536 connections 8:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3))"
tail -3 gpurun_out/r2_bench16.err
