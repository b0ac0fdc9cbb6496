#!/bin/bash
# round 2, GPU call 7: long tails of pipelined solves stay with the compacted sweeps
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined" 2>&1 | tail -3
for cfg in "100 16 16" "100 32 32" "100 24 24" "50 8 8"; do
  set -- $cfg
  timeout 600 python bench_sweep.py --horizons $1 --batches 65536 --streams $2 --pipeline $3 --reps 2 > gpurun_out/r2_sweepd_N$1_s$2_p$3.json 2>> gpurun_out/r2_run7.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_sweepd_N$1_s$2_p$3.json'))['rows'][0]; print('sweep', d['N'], 'streams', d['streams'], 'pipe', d['pipeline_depth'], round(d['solves_per_s']/1e3,1), 'k/s', round(d['ms_per_batch'],1), 'ms', d['max_iters'], d['status_hist'])"
done
tail -5 gpurun_out/r2_run7.err
