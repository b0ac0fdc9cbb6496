#!/bin/bash
# A/B of kernel variants: usage tools/ab.sh <reps> name1 name2 ...  ("base" = the default library); alternates the
# variants rep times so clock / box drift hits all of them alike
reps=$1; shift
mkdir -p gpurun_out
for r in $(seq 1 $reps); do for v in "$@"; do
  if [ "$v" = base ]; then lib=$PWD/udacitympc_b200/lib/libb200mpc.so; else lib=$PWD/udacitympc_b200/lib/libb200mpc_$v.so; fi
  B200MPC_LIB=$lib timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --latency-reps 5 $AB_ARGS > gpurun_out/ab_${v}_$r.json 2> gpurun_out/ab_${v}_$r.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${v}_$r.json")); print("$v rep $r", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["solved_fraction"])
except Exception as e: print("$v rep $r failed", e)
PY
done; done
