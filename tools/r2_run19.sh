#!/bin/bash
mkdir -p gpurun_out
for S in 4 6 8 12; do
python bench.py --streams $S --no-cpu-baseline --no-sweep --latency-reps 10 > gpurun_out/r2_b64k_s$S.json 2>> gpurun_out/r2_run19.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b64k_s$S.json')); print('B=65536 streams $S:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3))"
done
for c in "0.8,3" "0.6,4" "0.85,4" "0.7,6"; do
B200MPC_COMPACT=$c python bench.py --no-cpu-baseline --no-sweep --latency-reps 10 > gpurun_out/r2_b64k_c$c.json 2>> gpurun_out/r2_run19.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b64k_c$c.json')); print('B=65536 compact $c:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3))"
done
tail -3 gpurun_out/r2_run19.err
