"""-m "not gpu": the CPU oracle (oracle/mpc_oracle.c, the "port") pinned against
  (a) the reference's own in-source known answers, and
  (b) golden vectors produced by the reference itself (tests/golden/make_golden.py: the reference's prebuilt
      Ipopt 3.12.7 + MUMPS binaries on the MPC.cpp NLP; the reference's helpers.h / globalKinematic)."""
import numpy as np
import pytest

import oracle_bindings as ob
from conftest import golden

# polyfit/solution/main.cpp:18-20 (inputs), :33-54 (expected polyeval at x = 0..20, 6 significant digits)
QUIZ_X = [9.261977, -2.06803, -19.6663, -36.868, -51.6263, -66.3482]
QUIZ_Y = [5.17, -2.25, -15.306, -29.46, -42.85, -57.6116]
QUIZ_EXPECT = [-0.905562, -0.226606, 0.447594, 1.11706, 1.7818, 2.44185, 3.09723, 3.74794, 4.39402, 5.03548, 5.67235,
               6.30463, 6.93236, 7.55555, 8.17423, 8.7884, 9.3981, 10.0033, 10.6041, 11.2005, 11.7925]


def sig6(v):
    return float(f"{v:.6g}")


def test_polyfit_known_answer():
    c = ob.port_polyfit(QUIZ_X, QUIZ_Y, 3)
    got = [sig6(ob.port_polyeval(c, float(x))) for x in range(21)]
    assert got == QUIZ_EXPECT


def test_kinematic_known_answer():
    # global_kinematic_model/solution/main.cpp:27-30
    nxt = ob.port_kinematic([0, 0, np.deg2rad(45), 1], [np.deg2rad(5), 1], 0.3, 2.0)
    assert [sig6(v) for v in nxt] == [0.212132, 0.212132, 0.798488, 1.3]


def test_polyfit_matches_reference_fixtures():
    g = golden("polyfit_shapes.npz")
    for key in [k for k in g.files if k.startswith("fit_")]:
        _, m, order = key.split("_")
        xs, ys, fit = g[f"xs_{m}_{order}"], g[f"ys_{m}_{order}"], g[key]
        for b in range(xs.shape[0]):
            np.testing.assert_allclose(ob.port_polyfit(xs[b], ys[b], int(order)), fit[b], rtol=0, atol=1e-10 * max(1, np.abs(fit[b]).max()))
    r = golden("roadmap_256.npz")
    for b in range(0, 256, 8):
        np.testing.assert_allclose(ob.port_polyfit(r["xs"][b], r["ys"][b], 3), r["fit"][b], rtol=0, atol=1e-10)


def test_polyfit_rejects_bad_order():
    with pytest.raises(ValueError):
        ob.port_polyfit([0.0, 1.0], [0.0, 1.0], 2)   # helpers.h:26 assert order <= m-1
    with pytest.raises(ValueError):
        ob.port_polyfit([0.0, 1.0, 2.0], [0.0, 1.0, 2.0], 0)


def test_kinematic_matches_reference_fixtures():
    g = golden("kinematic_256.npz")
    for b in range(256):
        np.testing.assert_allclose(ob.port_kinematic(g["states"][b], g["act"][b, 0], 0.3, 2.0), g["one_step"][b], rtol=0, atol=1e-12)
    for b in range(0, 256, 16):
        s = g["states"][b].copy()
        for h in range(25):
            s = ob.port_kinematic(s, g["act"][b, h], 0.3, 2.0)
            np.testing.assert_allclose(s, g["rollout"][b, h], rtol=0, atol=1e-12)


def test_mpc_config1_closed_loop_prefix():
    g = golden("config1_closed_loop.npz")
    for k in (0, 1, 5, 24, 49):
        r = ob.port_solve(g["states"][k], g["coeffs"])
        assert r["status"] == 0 and r["iters"] == g["iters"][k]
        np.testing.assert_allclose(r["out8"], g["out8"][k], rtol=0, atol=1e-9)
        np.testing.assert_allclose(r["x"], g["x"][k], rtol=0, atol=1e-9)
        assert abs(r["obj"] - g["cost"][k]) <= 1e-10 * abs(g["cost"][k])


@pytest.mark.parametrize("name,step", [("line_256.npz", 16), ("roadmap_256.npz", 16)])
def test_mpc_random_problems(name, step):
    g = golden(name)
    cf = g["coeffs"] if "coeffs" in g.files else g["fit"]
    for b in range(0, 256, step):
        r = ob.port_solve(g["states"][b], cf[b])
        assert r["status"] == g["status"][b] == 0
        assert r["iters"] == g["iters"][b]
        np.testing.assert_allclose(r["x"], g["x"][b], rtol=0, atol=1e-9)
        assert abs(r["obj"] - g["obj"][b]) <= 1e-10 * abs(g["obj"][b])


def test_mpc_other_parameters():
    g = golden("line_params_64.npz")
    N, dt, Lf, ref_v, dmax, amax = g["params"]
    p = ob.default_params(N=int(N), dt=dt, Lf=Lf, ref_v=ref_v, delta_max=dmax, a_max=amax)
    n = 0
    for b in range(0, 64, 3):
        if g["status"][b] != 0 or g["used_restoration"][b]:
            continue   # the port does not restate the restoration phase (documented gap): it returns -2 there
        r = ob.port_solve(g["states"][b], g["coeffs"][b], params=p)
        assert r["status"] == 0 and r["iters"] == g["iters"][b]
        np.testing.assert_allclose(r["x"], g["x"][b], rtol=0, atol=1e-9)
        n += 1
    assert n >= 10


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built (no /root/reference)")
def test_port_derivatives_match_reference_tnlp():
    rng = np.random.default_rng(3)
    c = np.array([0.3, -0.2, 0.01, 2e-4])
    x = rng.normal(size=198); lam = rng.normal(size=150)
    a = ob.port_eval(x, lam, c, sigma=0.7)
    b = ob.ref_eval(x, lam, c, sigma=0.7)
    for u, v in zip(a, b):
        np.testing.assert_allclose(u, v, rtol=0, atol=1e-12)


def test_port_soft_restoration_phase_matches_reference():
    """Problems on which the reference's Ipopt takes soft-restoration steps (tests/golden/soft_N50_3.npz): the port
    restates that phase (IpBacktrackingLineSearch.cpp:426-448, 498-530, 595-603, 1043-1140) and reproduces the
    reference's iteration counts and solutions."""
    g = golden("soft_N50_3.npz")
    for b in range(3):
        o = ob.port_solve(g["states"][b], g["coeffs"][b], params=ob.default_params(N=50))
        assert o["status"] == 0 and o["iters"] == g["iters"][b]
        np.testing.assert_allclose(o["x"], g["x"][b], rtol=0, atol=1e-10)
        assert abs(o["obj"] - g["obj"][b]) <= 1e-12 * abs(g["obj"][b])


def test_frontend_fixture_fit_is_reproduced_by_the_port():
    """frontend_256.npz: the fit stored there is the reference's polyfit; the port's QR restatement agrees to 1e-10."""
    g = golden("frontend_256.npz")
    cl = g["centerline"]
    for b in range(0, 256, 5):
        x, y, p = g["poses"][b, :3]
        d = (x - cl[:, 0]) * (x - cl[:, 0]) + (y - cl[:, 1]) * (y - cl[:, 1])
        i0 = int(np.argmin(d))
        assert i0 == g["nearest"][b]
        i0 = min(i0, len(cl) - 6)
        dx, dy = cl[i0:i0 + 6, 0] - x, cl[i0:i0 + 6, 1] - y
        c = ob.port_polyfit(np.cos(p) * dx + np.sin(p) * dy, np.cos(p) * dy - np.sin(p) * dx, 3)
        np.testing.assert_allclose(c, g["coeffs"][b], rtol=0, atol=1e-10 * max(1.0, np.abs(c).max()))
