#!/usr/bin/env python
"""One pipelined solve at a time (depth 1): prints its wall time; under ncu it gives the launch list of bulk + tail."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import udacitympc_b200 as mp
from udacitympc_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
with mp.MPC(device=0) as m0:
    xs, ys = synth.roadmap_windows(B)
    fit = mp.polyfit_batch(xs, ys, 3, mpc=m0)
st = synth.roadmap_problems(B, fit)
st_d = torch.from_numpy(np.ascontiguousarray(st.T)).to(dev); cf_d = torch.from_numpy(np.ascontiguousarray(fit.T)).to(dev)
o = dict(out8=torch.empty((8, B), dtype=torch.float64, device=dev), status=torch.empty(B, dtype=torch.int32, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
s = torch.cuda.Stream(device=dev)
with mp.MPC(device=0, N=N) as m:
    m.set_batch_split(1)
    m.set_pipeline(1, 4096)
    for i in range(calls):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.solve_batch_device(B, st_d.data_ptr(), cf_d.data_ptr(), 4, o["out8"].data_ptr(), 0, 0, o["status"].data_ptr(), o["iters"].data_ptr(), s.cuda_stream)
        torch.cuda.synchronize()
        print("call", i, "ms", 1e3 * (time.perf_counter() - t0), "launches", m.launch_count(), flush=True)
