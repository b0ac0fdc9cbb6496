"""Determinism stress: the same device-resident batch is solved `reps` times on `streams` overlapped handles; every
result must be bit-identical to the first one.  usage: python tools/stress_repeat.py [reps] [batch] [streams] [input set]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    shift = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # bench.py's input set number
    import torch
    import udacitympc_b200 as m
    from udacitympc_b200 import synth
    dev = torch.device("cuda", 0)
    mpcs = [m.MPC(device=0) for _ in range(S)]
    for h in mpcs:
        h.set_batch_split(1 if S > 1 else 4)
    xs, ys = synth.roadmap_windows(B, synth.MT19937_64(synth.SEED + 1000 * shift))
    fit = m.polyfit_batch(xs, ys, 3, mpc=mpcs[0])
    st = synth.roadmap_problems(B, fit, synth.MT19937_64(synth.SEED + 1 + 1000 * shift))
    st_d = torch.from_numpy(np.ascontiguousarray(st.T)).to(dev)
    cf_d = torch.from_numpy(np.ascontiguousarray(fit.T)).to(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    outs = [dict(out8=torch.empty((8, B), dtype=torch.float64, device=dev), obj=torch.empty(B, dtype=torch.float64, device=dev),
                 status=torch.empty(B, dtype=torch.int32, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
            for _ in range(reps)]
    torch.cuda.synchronize()
    for i in range(reps):
        o = outs[i]
        mpcs[i % S].solve_batch_device(B, st_d.data_ptr(), cf_d.data_ptr(), 4, o["out8"].data_ptr(), 0, o["obj"].data_ptr(),
                                       o["status"].data_ptr(), o["iters"].data_ptr(), streams[i % S].cuda_stream)
    torch.cuda.synchronize()
    bad = 0
    for i in range(reps):
        o = outs[i]
        nz = int((o["status"] != 0).sum().item())
        d8 = int((o["out8"] != outs[0]["out8"]).any(dim=0).sum().item())
        di = int((o["iters"] != outs[0]["iters"]).sum().item())
        if nz or d8 or di:
            bad += 1
            idx = torch.nonzero((o["out8"] != outs[0]["out8"]).any(dim=0) | (o["status"] != 0)).flatten()[:5].tolist()
            print(f"rep {i}: status!=0 {nz}, out8 differs {d8}, iters differ {di}, e.g. problems {idx}, status {[int(o['status'][j]) for j in idx]}, iters {[int(o['iters'][j]) for j in idx]} vs {[int(outs[0]['iters'][j]) for j in idx]}")
    print(f"set {shift} max iters {int(outs[0]['iters'].max())} {os.environ.get('B200MPC_LIB', 'default lib')}: {reps} reps x {B} problems on {S} streams: {bad} reps differ")
