#!/bin/bash
# usage: tools_variants.sh name1 name2 ... ; runs bench for each variant lib, prints value/ms
mkdir -p gpurun_out
for v in "$@"; do
  B200MPC_LIB=$PWD/udacitympc_b200/lib/libb200mpc_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --latency-reps 5 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$v.json")); print("$v", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["solved_fraction"])
except Exception as e: print("$v failed", e)
PY
done
