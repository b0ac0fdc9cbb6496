#!/usr/bin/env python
"""bench_sweep.py -- BASELINE.json configs[4]: horizon / batch sweep (N in 10..100, batch 1K..1M), degree-3 reference
from roadmap windows, one B200.  Device-resident inputs, CUDA events.  For every point: solves/s, mean / max
interior-point iterations and the status histogram (a failed line search, where Ipopt enters its restoration phase, is
followed by the library's restoration step -- b200mpc_set_restoration, on by default; status -2 only with it off).
`python bench_sweep.py > profiles/r1_sweep.json`"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--horizons", default="10,25,50,100")
    ap.add_argument("--batches", default="1024,16384,65536,262144,1048576")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--streams", type=int, default=1, help="solver handles / streams whose batches overlap (as bench.py does); "
                    "1 = one batch at a time, where the slowest problem of a batch sets its time")
    ap.add_argument("--pipeline", type=int, default=0, help="> 0: the overlapped batches go to ONE solver handle in pipelined mode with this many tail "
                    "contexts (b200mpc_set_pipeline): memory of one full-size workspace + the tail contexts instead of one workspace per stream")
    ap.add_argument("--pipeline-slots", type=int, default=0, help="problems a tail context holds (0 = max(1024, batch / 16))")
    ap.add_argument("--rounds", type=int, default=0, help="per-pass rounds before the cooperative finisher (0 = library default)")
    args = ap.parse_args()
    import torch
    import udacitympc_b200 as mp
    from udacitympc_b200 import synth
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    Bmax = max(int(b) for b in args.batches.split(","))
    nb = min(Bmax, 262144)
    with mp.MPC(device=0) as m0:
        xs, ys = synth.roadmap_windows(nb)
        fit = mp.polyfit_batch(xs, ys, 3, mpc=m0)
    st = synth.roadmap_problems(nb, fit)
    rows = []
    for N in [int(n) for n in args.horizons.split(",")]:
        for B in [int(b) for b in args.batches.split(",")]:
            if B * 83 * (N + 2) * 8 > 60e9:
                continue
            t = (B + nb - 1) // nb
            st_d = torch.from_numpy(np.ascontiguousarray(np.tile(st, (t, 1))[:B].T)).to(dev)
            cf_d = torch.from_numpy(np.ascontiguousarray(np.tile(fit, (t, 1))[:B].T)).to(dev)
            S = max(1, min(args.streams, int(150e9 // (2 * B * 83 * (N + 2) * 8)))) if args.pipeline <= 0 else max(1, args.streams)
            streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(S - 1)]
            out8 = [torch.empty((8, B), dtype=torch.float64, device=dev) for _ in range(S)]
            status = torch.empty((S, B), dtype=torch.int32, device=dev)
            iters = torch.empty((S, B), dtype=torch.int32, device=dev)
            own = [mp.MPC(device=0, N=N) for _ in range(S if args.pipeline <= 0 else 1)]
            handles = own if args.pipeline <= 0 else own * S
            try:
                for m in own:
                    m.set_solver_mode(0, args.rounds, -1)
                    if S > 1:
                        m.set_batch_split(1)   # the caller overlaps the batches itself
                    if args.pipeline > 0 and S > 1:
                        m.set_pipeline(args.pipeline, args.pipeline_slots if args.pipeline_slots > 0 else max(1024, B // 16))

                def step(k):
                    handles[k].solve_batch_device(B, st_d.data_ptr(), cf_d.data_ptr(), 4, out8[k].data_ptr(), 0, 0,
                                                  status[k].data_ptr(), iters[k].data_ptr(), streams[k].cuda_stream)
                for k in range(S):
                    step(k)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for s_ in streams[1:]:
                    s_.wait_event(e0)
                for _ in range(args.reps):
                    for k in range(S):
                        step(k)
                for s_ in streams[1:]:
                    ev = torch.cuda.Event()
                    ev.record(s_)
                    stream.wait_event(ev)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / (args.reps * S)
            finally:
                for m in own:
                    m.close()
            status, iters = status[0], iters[0]
            sc = status.cpu().numpy()
            it = iters.cpu().numpy()
            hist = {int(k): int(v) for k, v in zip(*np.unique(sc, return_counts=True))}
            rows.append(dict(N=N, batch=B, rounds=args.rounds, streams=S, pipeline_depth=args.pipeline if S > 1 else 0, solver_handles=len(own), ms_per_batch=ms, solves_per_s=B / (ms * 1e-3), mean_iters=float(it.mean()),
                             max_iters=int(it.max()), status_hist=hist, solved_fraction=float((sc == 0).mean())))
            print(json.dumps(rows[-1]), file=sys.stderr)
            del st_d, cf_d, out8
            torch.cuda.empty_cache()
    print(json.dumps(dict(workload="degree-3 reference from roadmap windows (SURVEY 8d config 5), dt=0.05", rows=rows), indent=1))


if __name__ == "__main__":
    main()
