"""Iteration counts of the default bench workload (4 cycled 65 536-problem roadmap sets) per 8-way index shard, computed on
the CPU with the host build of the solver core (tests/hostsim): which rank of the 8-GPU strong-scaling run holds a straggler.
usage: python tools/iters_by_shard.py profiles/r2_iters_by_shard.json"""
import sys, os, numpy as np, json, time
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
import bench
from conftest import _HostSim
hs = _HostSim()
B = 65536
out = {}
t0 = time.time()
for s in range(4):
    st, cf = bench.make_workload("roadmap", B, seed_shift=s, mpc=None)
    its = np.zeros(B, dtype=np.int32)
    for c in range(0, B, 4096):
        r = hs.batch_interleaved(st[c:c+4096], cf[c:c+4096], compact=False)
        its[c:c+4096] = r["iters"]
    per = [int(its[k*8192:(k+1)*8192].max()) for k in range(8)]
    mean = [round(float(its[k*8192:(k+1)*8192].mean()), 3) for k in range(8)]
    top = np.argsort(-its)[:6]
    out[s] = dict(max_per_shard=per, mean_per_shard=mean, top=[(int(i), int(its[i]), int(i // 8192)) for i in top])
    print(s, out[s], round(time.time() - t0), flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "/tmp/iters_by_shard.json", "w"), indent=1)
