#!/bin/bash
# measures the internal batch split of one solve call against caller-side overlap: args are "streams:split" pairs
mkdir -p gpurun_out
for pair in "$@"; do
  st=${pair%%:*}; sp=${pair##*:}
  timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --latency-reps 20 --streams $st --split $sp > gpurun_out/split_${st}_${sp}.json 2> gpurun_out/split_${st}_${sp}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/split_${st}_${sp}.json")); print("streams $st split $sp", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["solved_fraction"], {k: v for k, v in d.items() if "lat" in k})
except Exception as e: print("streams $st split $sp failed", e)
PY
done
