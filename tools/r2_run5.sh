#!/bin/bash
# round 2, GPU call 5: pipelined solves (tail hand-off) -- parity test, N=25 bench with one handle, long-horizon sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or compaction or deterministic" > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests5.log
tail -15 gpurun_out/r2_tests5.log
for cfg in "0 6" "6 6" "8 8" "12 12"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --latency-reps 5 --pipeline $1 --streams $2 > gpurun_out/r2_pipe_p$1_s$2.json 2>> gpurun_out/r2_run5.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_pipe_p$1_s$2.json')); print('N=25 pipeline $1 streams $2:', round(d['value']/1e6,3), 'M/s')"
done
for cfg in "100 4 0" "100 16 16" "100 32 32" "50 8 8"; do
  set -- $cfg
  timeout 600 python bench_sweep.py --horizons $1 --batches 65536 --streams $2 --pipeline $3 --reps 2 > gpurun_out/r2_sweep_N$1_s$2_p$3.json 2>> gpurun_out/r2_run5.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_sweep_N$1_s$2_p$3.json'))['rows'][0]; print('sweep', d['N'], 'streams', d['streams'], 'pipe', d['pipeline_depth'], round(d['solves_per_s']/1e3,1), 'k/s', round(d['ms_per_batch'],1), 'ms', d['max_iters'], d['status_hist'])"
done
tail -5 gpurun_out/r2_run5.err
