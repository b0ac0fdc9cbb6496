#!/bin/bash
# round 2, GPU call 20: compute-sanitizer (memcheck, then racecheck) over every kernel incl. the round-2 paths
mkdir -p gpurun_out
python tools/smoke_workload.py > gpurun_out/r2_smoke_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2_smoke_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/smoke_workload.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/r2_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/smoke_workload.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/r2_sanitizer_racecheck.log
