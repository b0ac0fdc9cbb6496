# max rounds before the hand-over to the cooperative kernel at long horizons, with overlapped batches (4 streams)
for cfg in "100 1184,120" "100 600,200" "50 1184,40"; do
  set -- $cfg
  echo "N=$1 HANDOVER=$2"
  B200MPC_HANDOVER=$2 timeout 40 python bench_sweep.py --horizons $1 --batches 65536 --reps 2 --streams 4 2>&1 | grep '^{"N"' | python -c "import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ', d['ms_per_batch'], d['solves_per_s'], d['mean_iters'], d['max_iters'], d['status_hist'])"
done
