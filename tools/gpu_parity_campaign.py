"""TEST INFRASTRUCTURE: parity campaign ON THE GPU BOX.  n random problems of each workload are solved by the CUDA
library through the C ABI (default settings: per-pass kernels, batch compaction, internal split, cooperative finisher)
and by THE REFERENCE ITSELF (oracle/_ref: the reference's Ipopt 3.12.7 + MUMPS binaries, one process per host core);
prints / writes the mismatch statistics.

    python tools/gpu_parity_campaign.py 16384 [out.json]
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_bindings as ob  # noqa: E402
from udacitympc_b200 import synth  # noqa: E402


def ref(a):
    r = ob.ref_solve(a[0], a[1], trace=True)
    return r["x"], r["obj"], r["status"], r["iters"], int((r["trace"][:, 9] >= 100).any())


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    out = sys.argv[2] if len(sys.argv) > 2 else None
    import udacitympc_b200 as m
    rows = []
    with m.MPC(device=0) as mpc:
        st1, cf1 = synth.line_problems(n, synth.MT19937_64(4242))
        xs, ys = synth.roadmap_windows(n, synth.MT19937_64(4243))
        fit = m.polyfit_batch(xs, ys, 3, mpc=mpc)
        st3 = synth.roadmap_problems(n, fit, synth.MT19937_64(4244))
        for name, S, C in (("line (degree 1)", st1, cf1), ("roadmap (degree 3)", st3, fit)):
            g = mpc.solve_batch(S, C, want_traj=True)
            t = time.time()
            with mp.get_context("fork").Pool(os.cpu_count()) as pool:
                res = pool.map(ref, [(S[b], C[b]) for b in range(n)], chunksize=64)
            X = np.array([r[0] for r in res]); obj = np.array([r[1] for r in res])
            status = np.array([r[2] for r in res]); iters = np.array([r[3] for r in res]); resto = np.array([r[4] for r in res])
            ok = (status == 0) & (resto == 0)
            row = dict(workload=name, n=n, horizon_N=25, reference_failed_or_used_restoration=int((~ok).sum()),
                       status_mismatches=int((g["status"][ok] != status[ok]).sum()),
                       iteration_count_mismatches=int((g["iters"][ok] != iters[ok]).sum()),
                       max_iters=int(iters.max()),
                       max_abs_dx=float(np.abs(g["traj"][ok] - X[ok]).max()),
                       max_abs_d_actuators=float(np.abs(g["out8"][ok][:, 6:] - X[ok][:, [6 * 25, 7 * 25 - 1]]).max()),
                       max_rel_dobj=float((np.abs(g["cost"][ok] - obj[ok]) / np.abs(obj[ok])).max()),
                       reference_seconds=round(time.time() - t, 1), host_cores=os.cpu_count())
            rows.append(row)
            print(json.dumps(row), flush=True)
    if out:
        json.dump(dict(what="CUDA library (C ABI, default settings) vs the reference's Ipopt+MUMPS binaries on the same seeded inputs",
                       tolerances=dict(actuators=1e-5, trajectory=1e-5, objective_rel=1e-6), rows=rows), open(out, "w"), indent=1)
