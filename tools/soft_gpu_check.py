"""One-off GPU check of the soft restoration phase (b200mpc_set_restoration(h, 2)) against the golden sets."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import udacitympc_b200 as mp
G = os.path.join(ROOT, "tests", "golden")
for name, N in (("soft_N50_3.npz", 50), ("resto_N50_14.npz", 50), ("resto_N25_wild_32.npz", 25), ("resto_N100_48.npz", 100)):
    g = np.load(os.path.join(G, name))
    for reps in (1, 100):
        st, cf = np.tile(g["states"], (reps, 1)), np.tile(g["coeffs"], (reps, 1))
        with mp.MPC(N=N) as m:
            m.set_restoration(2)
            if reps > 1:
                m.set_solver_mode(0, 14, 0)
            r = m.solve_batch(st, cf)
        n = len(g["obj"])
        same = np.abs(r["cost"][:n] - g["obj"]) <= 1e-6 * np.abs(g["obj"])
        print(name, "reps", reps, "status", np.unique(r["status"]), "same", int(same.sum()), "/", n, "iters", r["iters"][:3], g["iters"][:3],
              "copies agree", bool(np.abs(r["out8"].reshape(reps, n, 8) - r["out8"][:n]).max() < 1e-7), flush=True)
