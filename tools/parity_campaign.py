"""TEST INFRASTRUCTURE: statistical parity campaign in the build container (no GPU): the solver core compiled for the
host (tests/hostsim: mode 0 = thread-per-problem path, mode 2 = cooperative warp-per-problem path) against THE
REFERENCE ITSELF (oracle/_ref: the reference's Ipopt 3.12.7 + MUMPS binaries) on n random problems of each workload.

    python tools/parity_campaign.py 65536 [modes, e.g. 0,2]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_bindings as ob  # noqa: E402
from conftest import _HostSim  # noqa: E402
from udacitympc_b200 import synth  # noqa: E402

hs = None


def work(a):
    global hs
    if hs is None:
        hs = _HostSim()
    st, cf, mode = a
    r = ob.ref_solve(st, cf, trace=True)
    h = hs.solve(st, cf, mode=mode)
    resto = int((r["trace"][:, 9] >= 100).any())
    return (r["status"], h["status"], r["iters"], h["iters"], float(np.abs(r["x"] - h["x"]).max()),
            abs(r["obj"] - h["obj"]) / abs(r["obj"]), resto)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    st1, cf1 = synth.line_problems(n, synth.MT19937_64(777))
    xs, ys = synth.roadmap_windows(n, synth.MT19937_64(778))
    V = np.stack([xs ** i for i in range(4)], axis=2)
    fit = np.stack([np.linalg.lstsq(V[b], ys[b], rcond=None)[0] for b in range(n)])
    st3 = synth.roadmap_problems(n, fit, synth.MT19937_64(779))
    for name, S, C in (("line", st1, cf1), ("roadmap", st3, fit)):
        for mode in modes:
            t = time.time()
            with mp.get_context("fork").Pool(os.cpu_count()) as pool:
                res = np.array(pool.map(work, [(S[b], C[b], mode) for b in range(n)], chunksize=32))
            ok = (res[:, 0] == 0) & (res[:, 6] == 0)
            print(f"{name} mode {mode} n {n}: reference failed/used restoration {int((~ok).sum())}, status mismatches "
                  f"{int((res[ok, 0] != res[ok, 1]).sum())}, iteration-count mismatches {int((res[ok, 2] != res[ok, 3]).sum())}, "
                  f"max iters {int(res[:, 2].max())}, max |dx| {res[ok, 4].max():.2e}, max rel dobj {res[ok, 5].max():.2e}, "
                  f"{time.time() - t:.0f} s", flush=True)
