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


def run_point(mp, torch, N, B, st, fit, streams=1, pipeline=0, pipeline_slots=0, reps=3, rounds=0, device=0):
    """One point of the sweep: B problems at horizon N (inputs tiled from the (nb, 6) / (nb, 4) arrays st / fit),
    `streams` overlapped batches on as many handles -- or on ONE handle in pipelined mode (pipeline > 0)."""
    dev = torch.device("cuda", device)
    nb = len(st)
    t = (B + nb - 1) // nb
    st_d = torch.from_numpy(np.ascontiguousarray(np.tile(st, (t, 1))[:B].T)).to(dev)
    cf_d = torch.from_numpy(np.ascontiguousarray(np.tile(fit, (t, 1))[:B].T)).to(dev)
    S = max(1, min(streams, int(150e9 // (2 * B * 83 * (N + 2) * 8)))) if pipeline <= 0 else max(1, streams)
    strs = [torch.cuda.Stream(device=dev) for _ in range(S)]
    out8 = [torch.empty((8, B), dtype=torch.float64, device=dev) for _ in range(S)]
    status = torch.empty((S, B), dtype=torch.int32, device=dev)
    iters = torch.empty((S, B), dtype=torch.int32, device=dev)
    own = [mp.MPC(device=device, N=N) for _ in range(S if pipeline <= 0 else 1)]
    handles = own if pipeline <= 0 else own * S
    try:
        for m in own:
            m.set_solver_mode(0, rounds, -1)
            if S > 1:
                m.set_batch_split(1)   # the caller overlaps the batches itself
            if pipeline > 0 and S > 1:
                m.set_pipeline(pipeline, pipeline_slots if pipeline_slots > 0 else max(1024, B // 16))

        def step(k):
            handles[k].solve_batch_device(B, st_d.data_ptr(), cf_d.data_ptr(), 4, out8[k].data_ptr(), 0, 0,
                                          status[k].data_ptr(), iters[k].data_ptr(), strs[k].cuda_stream)
        for k in range(S):
            step(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(strs[0])
        for s_ in strs[1:]:
            s_.wait_event(e0)
        for _ in range(reps):
            for k in range(S):
                step(k)
        for s_ in strs[1:]:
            ev = torch.cuda.Event()
            ev.record(s_)
            strs[0].wait_event(ev)
        e1.record(strs[0])
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * S)
    finally:
        for m in own:
            m.close()
    sc = status[0].cpu().numpy()
    it = iters[0].cpu().numpy()
    hist = {int(k): int(v) for k, v in zip(*np.unique(sc, return_counts=True))}
    row = dict(N=N, batch=B, rounds=rounds, streams=S, pipeline_depth=pipeline if S > 1 else 0, solver_handles=len(own),
               ms_per_batch=ms, solves_per_s=B / (ms * 1e-3), mean_iters=float(it.mean()), max_iters=int(it.max()), status_hist=hist,
               solved_fraction=float((sc == 0).mean()))
    del st_d, cf_d, out8
    torch.cuda.empty_cache()
    return row


def sweep_inputs(mp, nb):
    from udacitympc_b200 import synth
    with mp.MPC(device=0) as m0:
        xs, ys = synth.roadmap_windows(nb)
        fit = mp.polyfit_batch(xs, ys, 3, mpc=m0)
    return synth.roadmap_problems(nb, fit), fit


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
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per overlapped stream
    import torch
    import udacitympc_b200 as mp
    Bmax = max(int(b) for b in args.batches.split(","))
    st, fit = sweep_inputs(mp, min(Bmax, 262144))
    rows = []
    for N in [int(n) for n in args.horizons.split(",")]:
        for B in [int(b) for b in args.batches.split(",")]:
            if B * 83 * (N + 2) * 8 > 60e9:
                continue
            rows.append(run_point(mp, torch, N, B, st, fit, streams=args.streams, pipeline=args.pipeline, pipeline_slots=args.pipeline_slots,
                                  reps=args.reps, rounds=args.rounds))
            print(json.dumps(rows[-1]), file=sys.stderr)
    print(json.dumps(dict(workload="degree-3 reference from roadmap windows (SURVEY 8d config 5), dt=0.05", rows=rows), indent=1))


if __name__ == "__main__":
    main()
