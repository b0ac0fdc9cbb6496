#!/usr/bin/env python
"""Probe of the pipelined long-horizon path: iteration histogram of one N-horizon batch, latency of one pipelined call,
and throughput at a few depths.  usage: pipe_probe.py [N] [B]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import udacitympc_b200 as mp
from udacitympc_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = torch.device("cuda", 0)
with mp.MPC(device=0) as m0:
    xs, ys = synth.roadmap_windows(B)
    fit = mp.polyfit_batch(xs, ys, 3, mpc=m0)
st = synth.roadmap_problems(B, fit)
st_d = torch.from_numpy(np.ascontiguousarray(st.T)).to(dev); cf_d = torch.from_numpy(np.ascontiguousarray(fit.T)).to(dev)
out = {}
def mk():
    return dict(out8=torch.empty((8, B), dtype=torch.float64, device=dev), status=torch.empty(B, dtype=torch.int32, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
def call(m, o, s):
    m.solve_batch_device(B, st_d.data_ptr(), cf_d.data_ptr(), 4, o["out8"].data_ptr(), 0, 0, o["status"].data_ptr(), o["iters"].data_ptr(), s.cuda_stream)
def timed(m, outs, streams, calls):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(calls): call(m, outs[i % len(outs)], streams[i % len(streams)])
    torch.cuda.synchronize(); return (time.perf_counter() - t0)
for label, depth, slots, S in (("plain", 0, 0, 1), ("pipe1", 1, 4096, 1), ("pipe8", 8, 4096, 8), ("pipe32", 32, 4096, 32), ("pipe32_1k", 32, 1024, 32), ("pipe32_8k", 32, 8192, 32)):
    with mp.MPC(device=0, N=N) as m:
        m.set_batch_split(1)
        if depth: m.set_pipeline(depth, slots)
        streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
        outs = [mk() for _ in range(S)]
        timed(m, outs, streams, S)           # graphs
        t = timed(m, outs, streams, 2 * S)
        out[label] = dict(depth=depth, slots=slots, streams=S, ms_per_batch=1e3 * t / (2 * S), solves_per_s=B * 2 * S / t)
        print(label, out[label], flush=True)
        if label == "plain":
            it = outs[0]["iters"].cpu().numpy()
            out["iters_hist"] = {str(k): int((it > k).sum()) for k in (10, 15, 20, 30, 50, 100, 200, 400, 600, 800, 1000)}
            out["iters_mean_max"] = [float(it.mean()), int(it.max())]
            print(out["iters_hist"], out["iters_mean_max"], flush=True)
print(json.dumps(out))
