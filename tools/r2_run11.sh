#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined" 2>&1 | tail -3
for t in "0,0" "300,296"; do
  B200MPC_TAIL=$t timeout 600 python bench_sweep.py --horizons 100 --batches 65536 --streams 32 --pipeline 32 --reps 2 > gpurun_out/r2_sweepf_$t.json 2>> gpurun_out/r2_run11.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_sweepf_$t.json'))['rows'][0]; print('prio tail $t: sweep', d['N'], 'streams', d['streams'], 'pipe', d['pipeline_depth'], round(d['solves_per_s']/1e3,1), 'k/s', round(d['ms_per_batch'],1), 'ms', d['max_iters'], d['status_hist'])"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-reps 5 --pipeline 6 --streams 6 > gpurun_out/r2_pipeprio_p6_s6.json 2>> gpurun_out/r2_run11.err
python -c "
import json; d=json.load(open('gpurun_out/r2_pipeprio_p6_s6.json')); print('N=25 pipeline 6 streams 6 (priority tails):', round(d['value']/1e6,3), 'M/s')"
timeout 600 python bench_sweep.py --horizons 50 --batches 65536 --streams 8 --pipeline 8 --reps 2 2>> gpurun_out/r2_run11.err | python -c "
import json,sys; d=json.load(sys.stdin)['rows'][0]; print('N=50', round(d['solves_per_s']/1e3,1), 'k/s')"
tail -3 gpurun_out/r2_run11.err
