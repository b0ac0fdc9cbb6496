#!/bin/bash
# needs a library built with -DMPC_FUSE_FACTOR=1 (udacitympc_b200.build.build_variant + B200MPC_LIB, or the default build of that commit)
# round 2, GPU call 25: STEP sweep with the next iteration's Riccati factorisation riding on it (mpc_stepfactor_kernel):
# parity tests with it on, then A/B of the default bench against the separate kernels (B200MPC_FUSE=0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1 or config3 or roadmap or full_size or horizons or parameters or pipelined or compaction" > gpurun_out/r2_fuse_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_fuse_tests.log
for f in 1 0 1 0; do
B200MPC_FUSE=$f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sweep --latency-reps 5 > gpurun_out/r2_fuse_$f.json 2>> gpurun_out/r2_run25.err
python -c "
import json; d=json.load(open('gpurun_out/r2_fuse_$f.json')); print('fuse=$f', round(d['value']/1e6,3), 'M/s  e2e', round(d['e2e']['value']/1e6,3), ' lone', round(d['lone_caller']['value']/1e6,3), ' solved', d['solved_fraction'], 'iters', round(d['roofline']['mean_ip_iters'],4), 'launches', d['gpu_launches'])"
done
tail -3 gpurun_out/r2_run25.err
