#!/bin/bash
# round 2, GPU call 29: asynchronous staging in the forward sweep: parity tests, then A/B against the L1-prefetch version
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_fa_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_fa_tests.log
bash tools/ab.sh r2_fa "" nofa "" nofa
