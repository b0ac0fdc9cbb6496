#!/bin/bash
# round 2, GPU call 24: pipelined N = 100 solves -- tail schedules under which the cooperative kernel really takes the
# deep tail over (enough tail rounds for the live count to fall below the take-over threshold)
mkdir -p gpurun_out
O=gpurun_out/r2_pipe_tail_sched.jsonl; : > $O
export B200MPC_PIPE_DETACH=1
P="timeout 200 python tools/pipe_detach_probe.py 100 65536"
B200MPC_TAIL=800,128 $P t800_128 32 4096 1 1 >> $O 2>> gpurun_out/r2_run24.err
B200MPC_TAIL=800,296 $P t800_296 32 4096 1 1 >> $O 2>> gpurun_out/r2_run24.err
B200MPC_TAIL=1200,64 $P t1200_64 32 4096 1 1 >> $O 2>> gpurun_out/r2_run24.err
B200MPC_TAIL=600,296 $P t600_296 32 4096 1 1 >> $O 2>> gpurun_out/r2_run24.err
B200MPC_TAIL=800,128 $P t800_128_d64 64 4096 1 1 >> $O 2>> gpurun_out/r2_run24.err
B200MPC_TAIL=800,128 $P t800_128_d16 16 4096 1 1 >> $O 2>> gpurun_out/r2_run24.err
cat $O | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['label'], round(d['solves_per_s']/1e3,1), 'k/s', round(d['ms_per_batch'],1), 'ms/batch', d['all_solved'], d['max_iters'], 'warm', round(d['warm_s'],1))"
tail -5 gpurun_out/r2_run24.err
