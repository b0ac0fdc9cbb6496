#!/bin/bash
# A/B of library builds on the default bench: tools/ab.sh <tag> <variant suffix>...   ("" = the in-tree build)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/udacitympc_b200/lib/libb200mpc${v:+_$v}.so
  B200MPC_LIB=$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sweep --latency-reps 5 > gpurun_out/${tag}_${v:-base}.json 2>> gpurun_out/${tag}.err
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_${v:-base}.json')); print('variant [${v:-base}]', round(d['value']/1e6,3), 'M/s  lone', round(d['lone_caller']['value']/1e6,3), ' solved', d['solved_fraction'], 'iters', round(d['roofline']['mean_ip_iters'],4))"
done
