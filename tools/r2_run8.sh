#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python tools/pipe_probe.py 100 65536 > gpurun_out/r2_pipe_probe_N100.log 2>&1
cat gpurun_out/r2_pipe_probe_N100.log | head -20
