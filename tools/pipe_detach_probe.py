#!/usr/bin/env python
"""Pipelined long-horizon solves: does the tail latency under load come from the stream choreography?  Variants of the
N-horizon, B-problem pipelined run: callers' streams waiting for their tails (the library's stream-ordered contract)
against detached tails (B200MPC_PIPE_DETACH=1: one caller stream, completion by device synchronisation), at several
depths.  usage: pipe_detach_probe.py N B label depth slots caller_streams [calls_factor]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import udacitympc_b200 as mp
from udacitympc_b200 import synth
N, B, label, depth, slots, S = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
factor = int(sys.argv[7]) if len(sys.argv) > 7 else 2
dev = torch.device("cuda", 0)
with mp.MPC(device=0) as m0:
    xs, ys = synth.roadmap_windows(B)
    fit = mp.polyfit_batch(xs, ys, 3, mpc=m0)
st = synth.roadmap_problems(B, fit)
st_d = torch.from_numpy(np.ascontiguousarray(st.T)).to(dev); cf_d = torch.from_numpy(np.ascontiguousarray(fit.T)).to(dev)
nout = max(depth, S, 1)
outs = [dict(out8=torch.empty((8, B), dtype=torch.float64, device=dev), status=torch.full((B,), -99, dtype=torch.int32, device=dev),
             iters=torch.empty(B, dtype=torch.int32, device=dev)) for _ in range(nout)]
streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
def call(m, i):
    o = outs[i % nout]
    m.solve_batch_device(B, st_d.data_ptr(), cf_d.data_ptr(), 4, o["out8"].data_ptr(), 0, 0, o["status"].data_ptr(), o["iters"].data_ptr(), streams[i % S].cuda_stream)
def timed(m, calls):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(calls): call(m, i)
    torch.cuda.synchronize(); return time.perf_counter() - t0
kw = dict(max_iter=int(os.environ["PROBE_MAX_ITER"])) if os.environ.get("PROBE_MAX_ITER") else {}
with mp.MPC(device=0, N=N, **kw) as m:
    m.set_batch_split(1)
    if depth: m.set_pipeline(depth, slots)
    tw = timed(m, nout)                 # graphs are captured here
    t = timed(m, factor * nout)
    ok = all(int((o["status"] == 0).sum()) == B for o in outs) or bool(kw)
    it = outs[-1]["iters"].cpu().numpy()
    r = dict(label=label, N=N, B=B, depth=depth, slots=slots, caller_streams=S, detach=os.environ.get("B200MPC_PIPE_DETACH", ""), tail=os.environ.get("B200MPC_TAIL", ""), no_coop=os.environ.get("B200MPC_NO_COOP", ""), max_iter=os.environ.get("PROBE_MAX_ITER", ""),
             warm_s=tw, calls=factor * nout, seconds=t, ms_per_batch=1e3 * t / (factor * nout), solves_per_s=B * factor * nout / t, all_solved=ok,
             mean_iters=float(it.mean()), max_iters=int(it.max()))
print(json.dumps(r), flush=True)
