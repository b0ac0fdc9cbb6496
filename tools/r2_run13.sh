#!/bin/bash
# round 2, GPU call 13: full GPU suite + default bench line with the strided cooperative kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests13.log
tail -4 gpurun_out/r2_tests13.log
python bench.py > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench13.json'))
print(round(d['value']/1e6,3),'M/s e2e',round(d['e2e']['value']/1e6,3),'lone',round(d['lone_caller']['value']/1e6,3),'p99',round(d['e2e']['p99_batch_latency_ms'],2),'p50',round(d['e2e']['p50_batch_latency_ms'],2), 'frac', d['roofline']['frac'], 'hbm', d['roofline']['hbm_frac'], 'fexec', d['roofline']['frac_executed'])
P
python tools/latency.py 2>&1 | tail -6
