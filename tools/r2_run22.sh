#!/bin/bash
# round 2, GPU call 22: pipelined N = 100 solves -- callers' streams waiting for their tails against detached tails
mkdir -p gpurun_out
O=gpurun_out/r2_pipe_detach.jsonl; : > $O
P="timeout 300 python tools/pipe_detach_probe.py 100 65536"
$P base32 32 4096 32 >> $O 2>> gpurun_out/r2_run22.err
B200MPC_PIPE_DETACH=1 $P det32 32 4096 1 >> $O 2>> gpurun_out/r2_run22.err
B200MPC_PIPE_DETACH=1 $P det31 31 4096 1 >> $O 2>> gpurun_out/r2_run22.err
B200MPC_PIPE_DETACH=1 $P det16 16 4096 1 >> $O 2>> gpurun_out/r2_run22.err
B200MPC_PIPE_DETACH=1 $P det64 64 4096 1 >> $O 2>> gpurun_out/r2_run22.err
B200MPC_PIPE_DETACH=1 $P det128 128 2048 1 >> $O 2>> gpurun_out/r2_run22.err
B200MPC_PIPE_DETACH=1 B200MPC_TAIL=200,256 $P det32_t200_256 32 4096 1 >> $O 2>> gpurun_out/r2_run22.err
B200MPC_PIPE_DETACH=1 B200MPC_TAIL=600,128 $P det32_t600_128 32 4096 1 >> $O 2>> gpurun_out/r2_run22.err
cat $O | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['label'], round(d['solves_per_s']/1e3,1), 'k/s', round(d['ms_per_batch'],1), 'ms/batch', d['all_solved'], d['max_iters'])"
tail -5 gpurun_out/r2_run22.err
