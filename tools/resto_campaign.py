"""TEST INFRASTRUCTURE: how often does the restoration step (Solver::do_resto) end in the local minimum the reference's
Ipopt reaches through its own restoration phase?  Solver core compiled for the host (tests/hostsim) against THE
REFERENCE ITSELF (oracle/_ref binaries) on n random roadmap problems at horizon N.

    python tools/resto_campaign.py N n [mode] [--wild] [--soft]   (reference answers are cached in /tmp; --wild: initial
                                        states far outside the benchmark distribution; --soft: with Ipopt's soft
                                        restoration phase restated before the restoration step, Params::resto = 2)
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


def ref_work(a):
    st, cf, N = a
    r = ob.ref_solve(st, cf, N=N, trace=True)
    return np.concatenate([[r["status"], r["iters"], r["obj"], int((r["trace"][:, 9] >= 100).any())], r["out8"]])


def host_work(a):
    global hs
    if hs is None:
        hs = _HostSim()
        hs.lib.hostsim_set_restoration(2 if "--soft" in sys.argv else 1)
    st, cf, N, mode = a
    h = hs.solve(st, cf, mode=mode, N=N)
    return np.concatenate([[h["status"], h["iters"], h["obj"]], h["out8"]])


def line_problems_wild(n, seed=0):
    """Degree-1 reference y = -1 (the reference's own mpc_to_line setting) from far-off states: x +-50 m, lateral offset
    +-40 m, heading +-2 rad, speed 0.5..70 m/s."""
    u = synth.MT19937_64(881 + seed).uniform(4 * n).reshape(n, 4)
    x = -50.0 + 100.0 * u[:, 0]
    y = -1.0 - 40.0 + 80.0 * u[:, 1]
    psi = -2.0 + 4.0 * u[:, 2]
    v = 0.5 + 69.5 * u[:, 3]
    st = np.ascontiguousarray(np.stack([x, y, psi, v, -1.0 - y, psi], axis=1))
    return st, np.tile(np.array([-1.0, 0.0]), (n, 1))


def problems(n, seed=0, wild=False):
    xs, ys = synth.roadmap_windows(n, synth.MT19937_64(878 + seed))
    V = np.stack([xs ** i for i in range(4)], axis=2)
    fit = np.stack([np.linalg.lstsq(V[b], ys[b], rcond=None)[0] for b in range(n)])
    st = synth.roadmap_problems(n, fit, synth.MT19937_64(879 + seed))
    if wild:   # far outside SURVEY 8d config 4: lateral offset +-20 m, heading error +-1.5 rad, speed 1..60 m/s
        u = synth.MT19937_64(880 + seed).uniform(3 * n).reshape(n, 3)
        c0, c1 = fit[:, 0], fit[:, 1]
        if wild == 2:   # --wilder: +-80 m, +-3 rad, 0.1..80 m/s (objective scaling branch, |cte| > 50)
            y = -80.0 + 160.0 * u[:, 0]
            psi = np.arctan(c1) - 3.0 + 6.0 * u[:, 1]
            v = 0.1 + 79.9 * u[:, 2]
        else:
            y = -20.0 + 40.0 * u[:, 0]
            psi = np.arctan(c1) - 1.5 + 3.0 * u[:, 1]
            v = 1.0 + 59.0 * u[:, 2]
        st = np.ascontiguousarray(np.stack([np.zeros(n), y, psi, v, c0 - y, psi - np.arctan(c1)], axis=1))
    return st, fit


if __name__ == "__main__":
    N = int(sys.argv[1]); n = int(sys.argv[2]); mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    wild = 2 if "--wilder" in sys.argv else int("--wild" in sys.argv)
    S, C = problems(n, wild=wild)
    cache = f"/tmp/resto_ref_N{N}_{n}{['', '_wild', '_wilder'][wild]}.npy"
    if "--line" in sys.argv:
        S, C = line_problems_wild(n)
        cache = f"/tmp/resto_ref_N{N}_{n}_line.npy"
    ctx = mp.get_context("fork")
    if os.path.exists(cache):
        ref = np.load(cache)
    else:
        t = time.time()
        with ctx.Pool(os.cpu_count()) as pool:
            ref = np.array(pool.map(ref_work, [(S[b], C[b], N) for b in range(n)], chunksize=4))
        np.save(cache, ref)
        print(f"reference: {time.time() - t:.0f} s")
    t = time.time()
    with ctx.Pool(os.cpu_count()) as pool:
        got = np.array(pool.map(host_work, [(S[b], C[b], N, mode) for b in range(n)], chunksize=4))
    used = ref[:, 3] == 1
    ok_ref = ref[:, 0] == 0
    relobj = np.abs(got[:, 2] - ref[:, 2]) / np.abs(ref[:, 2])
    dact = np.abs(got[:, 9:11] - ref[:, 10:12]).max(axis=1)
    same = (got[:, 0] == 0) & ok_ref & (relobj < 1e-6) & (dact < 1e-5)
    ours_resto = (got[:, 0] == 0) & ok_ref & ~used & (got[:, 1] != ref[:, 1])
    print(f"  reference status histogram {np.unique(ref[:, 0], return_counts=True)}; status differs on {int((got[:, 0] != ref[:, 0]).sum())}")
    print(f"N {N} n {n} mode {mode}: reference ok {int(ok_ref.sum())}, reference used restoration {int(used.sum())}; "
          f"ours status!=0 {int((got[:, 0] != 0).sum())} {np.unique(got[:, 0], return_counts=True)}")
    print(f"  no-restoration problems: same solution {int((same & ~used).sum())} / {int((~used & ok_ref).sum())}, "
          f"iteration-count mismatches {int(ours_resto.sum())}")
    print(f"  restoration problems: same solution {int((same & used).sum())} / {int((used & ok_ref).sum())}; "
          f"ours better {int(((got[:, 0] == 0) & used & ~same & (got[:, 2] < ref[:, 2])).sum())}, "
          f"ours worse {int(((got[:, 0] == 0) & used & ~same & (got[:, 2] > ref[:, 2])).sum())}; "
          f"mean iters ours {got[used, 1].mean():.1f} ref {ref[used, 1].mean():.1f}; {time.time() - t:.0f} s")
    bad = np.nonzero(used & ~same)[0]
    for b in bad[:12]:
        print(f"    #{b}: ours status {int(got[b, 0])} iters {int(got[b, 1])} obj {got[b, 2]:.6f} | ref iters {int(ref[b, 1])} obj {ref[b, 2]:.6f}")
