#!/bin/bash
# round 2, GPU call 6: (a) per-thread cp.async staging of the factor sweep (variant "async") against the in-tree build;
# (b) long-horizon pipelined sweep with enough hardware queues for the overlapped streams
mkdir -p gpurun_out
bash tools/ab.sh r2_async "" async "" async
B200MPC_LIB=$PWD/udacitympc_b200/lib/libb200mpc_async.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_problems or full_size or compaction" 2>&1 | tail -3
export CUDA_DEVICE_MAX_CONNECTIONS=32
for cfg in "100 8 0" "100 16 16" "100 32 32"; do
  set -- $cfg
  timeout 600 python bench_sweep.py --horizons $1 --batches 65536 --streams $2 --pipeline $3 --reps 2 > gpurun_out/r2_sweepc_N$1_s$2_p$3.json 2>> gpurun_out/r2_run6.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_sweepc_N$1_s$2_p$3.json'))['rows'][0]; print('sweep (32 connections)', d['N'], 'streams', d['streams'], 'pipe', d['pipeline_depth'], round(d['solves_per_s']/1e3,1), 'k/s', round(d['ms_per_batch'],1), 'ms', d['max_iters'], d['status_hist'])"
done
tail -5 gpurun_out/r2_run6.err
