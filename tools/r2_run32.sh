#!/bin/bash
# round 2, GPU call 32: asynchronous staging in the step sweep: parity tests, then A/B against the L1-prefetch version
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_restoration.py -m gpu -x -q > gpurun_out/r2_sa_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_sa_tests.log
bash tools/ab.sh r2_sa "" nosa "" nosa
