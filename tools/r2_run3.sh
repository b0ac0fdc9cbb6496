#!/bin/bash
# round 2, GPU call 3: K6 fused into the host entry points -- full GPU test suite + default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests3.log
tail -5 gpurun_out/r2_tests3.log
python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench3.json'))
print(round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3),'lone',round(d['lone_caller']['value']/1e6,3),'p99',round(d['e2e']['p99_batch_latency_ms'],2),'p50',round(d['e2e']['p50_batch_latency_ms'],2))
P
