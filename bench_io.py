#!/usr/bin/env python
"""bench_io.py -- HBM-roofline measurement of the batch I/O-style kernels (BASELINE.json configs[1]):
  K4 polyfit   degree-3 fits of 6-point roadmap windows      128 B per fit  (16m + 8(d+1))
  K5 rollout   bicycle-model steps, H = 1 and H = 25         80 B per step; 32+16H in, 32H out per rollout
  K6 batch I/O [B][K] <-> [K][B] transposes through the host-buffer entry points (timed inside them by nothing here;
               reported through the end-to-end polyfit call)
Device-resident, field-major inputs; CUDA events on the launching stream; inputs larger than the 126 MB L2 so every
iteration streams from HBM.  Prints one JSON object; `python bench_io.py > profiles/r1_io_kernels.json`."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4 * 1048576)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    import udacitympc_b200 as mp
    from udacitympc_b200 import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench_io.py needs a CUDA device")
    dev = torch.device("cuda", 0)
    B = args.batch
    mpc = mp.MPC(device=0)
    lib = mp.load_library()
    stream = torch.cuda.Stream(device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps

    def chk(rc):
        if rc:
            raise RuntimeError(lib.b200mpc_last_error().decode())

    out = dict(batch=B, reps=args.reps, hbm_peak_gbs=hbm, peak_source=peak_src, kernels={})
    vp = ctypes.c_void_p

    # ---- K4: the 1M-window workload tiled up to B fits
    nb = min(B, 1048576)
    xs, ys = synth.roadmap_windows(nb)
    reps_t = (B + nb - 1) // nb
    xs_d = torch.from_numpy(np.ascontiguousarray(np.tile(xs, (reps_t, 1))[:B].T)).to(dev)
    ys_d = torch.from_numpy(np.ascontiguousarray(np.tile(ys, (reps_t, 1))[:B].T)).to(dev)
    cf_d = torch.empty((4, B), dtype=torch.float64, device=dev)
    ms = timeit(lambda: chk(lib.b200mpc_polyfit_batch_device(mpc.handle, B, vp(xs_d.data_ptr()), vp(ys_d.data_ptr()), 6, 3,
                                                             vp(cf_d.data_ptr()), vp(stream.cuda_stream))))
    byt = 128.0 * B
    out["kernels"]["polyfit_kernel<6,4>"] = dict(ms=ms, fits_per_s=B / (ms * 1e-3), bytes_per_unit=128, achieved_gbs=byt / (ms * 1e-3) / 1e9,
                                                 frac=byt / (ms * 1e-3) / 1e9 / hbm)
    # spot check against numpy
    ref = np.linalg.lstsq(np.stack([xs[0] ** i for i in range(4)], axis=1), ys[0], rcond=None)[0]
    assert np.allclose(cf_d[:, 0].cpu().numpy(), ref, atol=1e-9)

    # ---- K5: single step and 25-step rollouts
    for H in (1, 25):
        nbh = B if H == 1 else max(B // 8, 65536)
        st, act = synth.kinematic_inputs(min(nbh, 262144), H=H)
        t = (nbh + len(st) - 1) // len(st)
        st_d = torch.from_numpy(np.ascontiguousarray(np.tile(st, (t, 1))[:nbh].T)).to(dev)
        act_d = torch.from_numpy(np.ascontiguousarray(np.tile(act.reshape(len(st), -1), (t, 1))[:nbh].T)).to(dev)
        o_d = torch.empty((4 * H, nbh), dtype=torch.float64, device=dev)
        ms = timeit(lambda: chk(lib.b200mpc_rollout_batch_device(mpc.handle, nbh, H, vp(st_d.data_ptr()), vp(act_d.data_ptr()),
                                                                 ctypes.c_double(0.3), ctypes.c_double(2.0), vp(o_d.data_ptr()),
                                                                 vp(stream.cuda_stream))))
        per = 32 + 16 * H + 32 * H
        byt = float(per) * nbh
        out["kernels"][f"rollout_kernel H={H}"] = dict(ms=ms, batch=nbh, rollouts_per_s=nbh / (ms * 1e-3), steps_per_s=nbh * H / (ms * 1e-3),
                                                       bytes_per_unit=per, achieved_gbs=byt / (ms * 1e-3) / 1e9,
                                                       frac=byt / (ms * 1e-3) / 1e9 / hbm)

    # ---- end to end (host buffers, K6 transposes + copies inside): 1M fits
    import time
    xs1, ys1 = xs[:nb], ys[:nb]
    mp.polyfit_batch(xs1, ys1, 3, mpc=mpc)
    t0 = time.perf_counter()
    for _ in range(5):
        mp.polyfit_batch(xs1, ys1, 3, mpc=mpc)
    dt = (time.perf_counter() - t0) / 5
    out["e2e_polyfit_host_buffers"] = dict(batch=nb, ms=dt * 1e3, fits_per_s=nb / dt, note="pageable numpy buffers, H2D + 3 transposes + fit + D2H")
    print(json.dumps(out, indent=1))
    mpc.close()


if __name__ == "__main__":
    main()
