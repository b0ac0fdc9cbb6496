#!/bin/bash
# round 2, GPU call 31: compaction threshold / first round and hand-over threshold re-checked with the tuned kernels
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py --no-cpu-baseline --no-sweep --latency-reps 5 "$@" > gpurun_out/r2_retune_$tag.json 2>> gpurun_out/r2_run31.err
python -c "
import json; d=json.load(open('gpurun_out/r2_retune_$tag.json')); print('$tag:', round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3), 'lone', round(d['lone_caller']['value']/1e6,3))"; }
run base
B200MPC_COMPACT=0.65,4 run c065_4
B200MPC_COMPACT=0.75,4 run c075_4
B200MPC_COMPACT=0.7,5 run c07_5
B200MPC_COMPACT=0.7,3 run c07_3
run ho32 --handover 32
run ho128 --handover 128
run base2
tail -2 gpurun_out/r2_run31.err
