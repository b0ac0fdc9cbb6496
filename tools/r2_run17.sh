#!/bin/bash
# round 2, GPU call 17: hand-over threshold of the cooperative kernel at the 8-GPU shard size (8192 problems per GPU and step)
mkdir -p gpurun_out
for hv in 1184 592 296 128 0; do
B200MPC_HANDOVER=$hv python bench.py --batch 8192 --no-cpu-baseline --no-sweep --latency-reps 30 > gpurun_out/r2_b8192_h$hv.json 2>> gpurun_out/r2_run17.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b8192_h$hv.json')); print('B=8192 handover $hv:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), 'lone', round(d['lone_caller']['value']/1e6,3), 'p50', round(d['e2e']['p50_batch_latency_ms'],2), 'p99', round(d['e2e']['p99_batch_latency_ms'],2))"
done
for hv in 592 296; do
B200MPC_HANDOVER=$hv python bench.py --no-cpu-baseline --no-sweep --latency-reps 30 > gpurun_out/r2_b64k_h$hv.json 2>> gpurun_out/r2_run17.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b64k_h$hv.json')); print('B=65536 handover $hv:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), 'lone', round(d['lone_caller']['value']/1e6,3), 'p50', round(d['e2e']['p50_batch_latency_ms'],2), 'p99', round(d['e2e']['p99_batch_latency_ms'],2))"
done
tail -3 gpurun_out/r2_run17.err
