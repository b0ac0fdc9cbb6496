#!/bin/bash
# measures compaction settings: args are "<max live fraction>,<first round>:<rounds>" (0:<rounds> = compaction off)
mkdir -p gpurun_out
for pair in "$@"; do
  sc=${pair%%:*}; rd=${pair##*:}
  tag=$(echo "$sc" | tr ',' '_')_$rd
  B200MPC_COMPACT=$sc timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --latency-reps 5 --rounds $rd $CS_ARGS > gpurun_out/cs_$tag.json 2> gpurun_out/cs_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/cs_$tag.json")); print("compact $sc rounds $rd", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["solved_fraction"], d["gpu_launches"])
except Exception as e: print("compact $sc rounds $rd failed", e)
PY
done
