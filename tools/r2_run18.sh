#!/bin/bash
mkdir -p gpurun_out
for hv in -1 512 256 128 64; do
python bench.py --handover $hv --no-cpu-baseline --no-sweep --latency-reps 30 > gpurun_out/r2_b64k_ho$hv.json 2>> gpurun_out/r2_run18.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b64k_ho$hv.json')); print('B=65536 handover $hv:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), 'lone', round(d['lone_caller']['value']/1e6,3), 'p50', round(d['e2e']['p50_batch_latency_ms'],2), 'p99', round(d['e2e']['p99_batch_latency_ms'],2))"
done
for hv in 256 64 32; do
python bench.py --batch 8192 --handover $hv --no-cpu-baseline --no-sweep --latency-reps 30 > gpurun_out/r2_b8192_ho$hv.json 2>> gpurun_out/r2_run18.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b8192_ho$hv.json')); print('B=8192 handover $hv:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), 'p50', round(d['e2e']['p50_batch_latency_ms'],2))"
done
tail -3 gpurun_out/r2_run18.err
