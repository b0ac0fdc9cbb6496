#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
for t in "300,296" "150,296" "600,128"; do
  B200MPC_TAIL=$t timeout 600 python bench_sweep.py --horizons 100 --batches 65536 --streams 32 --pipeline 32 --reps 2 > gpurun_out/r2_sweepe_$t.json 2>> gpurun_out/r2_run10.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_sweepe_$t.json'))['rows'][0]; print('tail $t: sweep', d['N'], 'streams', d['streams'], 'pipe', d['pipeline_depth'], round(d['solves_per_s']/1e3,1), 'k/s', round(d['ms_per_batch'],1), 'ms', d['max_iters'], d['status_hist'])"
done
tail -3 gpurun_out/r2_run10.err
