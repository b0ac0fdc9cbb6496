#!/bin/bash
# round 2, GPU call 23: what bounds the pipelined N = 100 solves at ~380 k solves/s: cooperative-kernel launches inside
# the tails (200 KB shared memory per block), the launch count of the tails, or the bulks themselves
mkdir -p gpurun_out
O=gpurun_out/r2_pipe_bound.jsonl; : > $O
export B200MPC_PIPE_DETACH=1
P="timeout 200 python tools/pipe_detach_probe.py 100 65536"
B200MPC_NO_COOP=1 $P nocoop32 32 4096 1 1 >> $O 2>> gpurun_out/r2_run23.err
B200MPC_NO_COOP=1 $P nocoop16 16 4096 1 1 >> $O 2>> gpurun_out/r2_run23.err
B200MPC_TAIL=1,32 $P tail1_32 32 4096 1 1 >> $O 2>> gpurun_out/r2_run23.err
B200MPC_NO_COOP=1 B200MPC_TAIL=100 $P nocoop_tail100 32 4096 1 1 >> $O 2>> gpurun_out/r2_run23.err
PROBE_MAX_ITER=22 $P maxiter22 32 4096 1 1 >> $O 2>> gpurun_out/r2_run23.err
PROBE_MAX_ITER=60 $P maxiter60 32 4096 1 1 >> $O 2>> gpurun_out/r2_run23.err
PROBE_MAX_ITER=150 $P maxiter150 32 4096 1 1 >> $O 2>> gpurun_out/r2_run23.err
cat $O | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['label'], round(d['solves_per_s']/1e3,1), 'k/s', round(d['ms_per_batch'],1), 'ms/batch', d['all_solved'], d['max_iters'], 'warm', round(d['warm_s'],1))"
tail -5 gpurun_out/r2_run23.err
