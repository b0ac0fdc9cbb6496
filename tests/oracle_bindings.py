"""ctypes bindings for the TEST-ONLY oracles (oracle/libmpc_oracle.so = C port, oracle/_ref = the
reference's own Ipopt/MUMPS binaries and helpers).  Importable only from tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke(); never from the product package."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


class OracleParams(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("dt", ctypes.c_double), ("Lf", ctypes.c_double), ("ref_v", ctypes.c_double),
                ("w_cte", ctypes.c_double), ("w_epsi", ctypes.c_double), ("w_v", ctypes.c_double),
                ("w_delta", ctypes.c_double), ("w_a", ctypes.c_double), ("w_ddelta", ctypes.c_double),
                ("w_da", ctypes.c_double), ("delta_max", ctypes.c_double), ("a_max", ctypes.c_double),
                ("tol", ctypes.c_double), ("max_iter", ctypes.c_int)]


_port = None
_ref = None
_refh = None


def port():
    global _port
    if _port is None:
        L = ctypes.CDLL(os.path.join(ORACLE_DIR, "libmpc_oracle.so"))
        L.oracle_mpc_solve.restype = ctypes.c_int
        L.oracle_mpc_solve.argtypes = [ctypes.POINTER(OracleParams), dp, dp, ctypes.c_int, dp, dp, dp, ip, dp, dp,
                                       ctypes.c_int, ip]
        L.oracle_mpc_eval.restype = None
        L.oracle_mpc_eval.argtypes = [ctypes.POINTER(OracleParams), dp, ctypes.c_int, dp, dp, ctypes.c_double, dp, dp,
                                      dp, dp, dp]
        L.oracle_polyeval.restype = ctypes.c_double
        L.oracle_polyeval.argtypes = [dp, ctypes.c_int, ctypes.c_double]
        L.oracle_polyfit.restype = ctypes.c_int
        L.oracle_polyfit.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, dp]
        L.oracle_global_kinematic.restype = None
        L.oracle_global_kinematic.argtypes = [dp, dp, ctypes.c_double, ctypes.c_double, dp]
        L.oracle_default_params.argtypes = [ctypes.POINTER(OracleParams)]
        _port = L
    return _port


def default_params(**kw):
    p = OracleParams()
    port().oracle_default_params(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def port_solve(state6, coeffs, params=None, trace=False):
    p = params or default_params()
    n = 8 * p.N - 2
    st = np.ascontiguousarray(state6, dtype=np.float64)
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    x = np.zeros(n); o8 = np.zeros(8); obj = ctypes.c_double(); it = ctypes.c_int()
    lam = np.zeros(6 * p.N); tr = np.zeros((400, 10)); nr = ctypes.c_int()
    rc = port().oracle_mpc_solve(ctypes.byref(p), _p(st), _p(c), len(c), _p(x), _p(o8), ctypes.byref(obj),
                                 ctypes.byref(it), _p(lam), _p(tr) if trace else None, 400, ctypes.byref(nr))
    out = dict(status=rc, x=x, out8=o8, obj=obj.value, iters=it.value, lam=lam)
    if trace:
        out["trace"] = tr[:min(nr.value, 400)].copy()
    return out


def port_eval(x, lam, coeffs, sigma=1.0, params=None):
    p = params or default_params()
    n, m = 8 * p.N - 2, 6 * p.N
    x = np.ascontiguousarray(x, dtype=np.float64); lam = np.ascontiguousarray(lam, dtype=np.float64)
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    f = ctypes.c_double(); grad = np.zeros(n); g = np.zeros(m); J = np.zeros((m, n)); H = np.zeros((n, n))
    port().oracle_mpc_eval(ctypes.byref(p), _p(c), len(c), _p(x), _p(lam), sigma, ctypes.byref(f), _p(grad), _p(g),
                           _p(J), _p(H))
    return f.value, grad, g, J, H


def port_polyfit(xs, ys, order):
    xs = np.ascontiguousarray(xs, dtype=np.float64); ys = np.ascontiguousarray(ys, dtype=np.float64)
    out = np.zeros(order + 1)
    rc = port().oracle_polyfit(_p(xs), _p(ys), len(xs), order, _p(out))
    if rc != 0:
        raise ValueError("polyfit: order must satisfy 1 <= order <= m-1")
    return out


def cpu_fit_and_step_batch(xs, ys, order, state4, act2, dt, Lf=2.0):
    """B polyfit calls + B globalKinematic calls in one native loop each (bench_io.py's CPU arm): the reference's own
    helpers.h / globalKinematic (oracle/_ref, Lf = 2 as in its source) when built, else the C port.  Returns
    (coeffs (B, order+1), next4 (B, 4), kind)."""
    xs = np.ascontiguousarray(xs, dtype=np.float64); ys = np.ascontiguousarray(ys, dtype=np.float64)
    st = np.ascontiguousarray(state4, dtype=np.float64); ac = np.ascontiguousarray(act2, dtype=np.float64).reshape(len(st), 2)
    B, m = xs.shape
    cf = np.zeros((B, order + 1)); nx = np.zeros((len(st), 4))
    if ref_available() and hasattr(ref_helpers(), "ref_polyfit_batch") and Lf == 2.0:
        ref_helpers().ref_polyfit_batch(_p(xs), _p(ys), B, m, order, _p(cf))
        ref_helpers().ref_global_kinematic_batch(_p(st), _p(ac), len(st), dt, _p(nx))
        return cf, nx, "reference"
    L = port()
    L.oracle_polyfit_batch.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
    L.oracle_global_kinematic_batch.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, dp]
    L.oracle_polyfit_batch(_p(xs), _p(ys), B, m, order, _p(cf))
    L.oracle_global_kinematic_batch(_p(st), _p(ac), len(st), dt, Lf, _p(nx))
    return cf, nx, "port"


def port_polyeval(coeffs, x):
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    return port().oracle_polyeval(_p(c), len(c), float(x))


def port_kinematic(state4, act2, dt, Lf):
    s = np.ascontiguousarray(state4, dtype=np.float64); a = np.ascontiguousarray(act2, dtype=np.float64)
    out = np.zeros(4)
    port().oracle_global_kinematic(_p(s), _p(a), dt, Lf, _p(out))
    return out


# ------------------------------------------------------------------ the reference itself (oracle/_ref)
def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libmpc_ref.so"))


def ref():
    global _ref
    if _ref is None:
        L = ctypes.CDLL(os.path.join(ORACLE_DIR, "_ref", "libmpc_ref.so"))
        L.ref_mpc_solve.restype = ctypes.c_int
        L.ref_mpc_solve.argtypes = [ctypes.c_int] + [ctypes.c_double] * 5 + [dp, dp, ctypes.c_int, ctypes.c_char_p, dp,
                                                                             dp, dp, ip, dp, dp, dp, dp, ctypes.c_int, ip]
        L.ref_mpc_solve_w.restype = ctypes.c_int
        L.ref_mpc_solve_w.argtypes = [ctypes.c_int] + [ctypes.c_double] * 5 + [dp, dp, dp, ctypes.c_int, ctypes.c_char_p, dp,
                                                                               dp, dp, ip, dp, dp, dp, dp, ctypes.c_int, ip]
        L.ref_mpc_eval.restype = ctypes.c_int
        L.ref_mpc_eval.argtypes = [ctypes.c_int] + [ctypes.c_double] * 3 + [dp, ctypes.c_int, dp, dp, ctypes.c_double,
                                                                             dp, dp, dp, dp, dp]
        _ref = L
    return _ref


def ref_helpers():
    global _refh
    if _refh is None:
        H = ctypes.CDLL(os.path.join(ORACLE_DIR, "_ref", "libhelpers_ref.so"))
        H.ref_polyfit.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, dp]
        H.ref_polyeval.restype = ctypes.c_double
        H.ref_polyeval.argtypes = [dp, ctypes.c_int, ctypes.c_double]
        H.ref_global_kinematic.argtypes = [dp, dp, ctypes.c_double, dp]
        if hasattr(H, "ref_polyfit_batch"):
            H.ref_polyfit_batch.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
            H.ref_global_kinematic_batch.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double, dp]
        _refh = H
    return _refh


def ref_solve(state6, coeffs, N=25, dt=0.05, Lf=2.67, ref_v=40.0, delta_max=0.436332, a_max=1.0, opts="", trace=False,
              weights=None):
    """weights = (w_cte, w_epsi, w_v, w_delta, w_a, w_ddelta, w_da) on the seven cost terms of MPC.cpp:57-76 (None = 1)."""
    n = 8 * N - 2
    st = np.ascontiguousarray(state6, dtype=np.float64); c = np.ascontiguousarray(coeffs, dtype=np.float64)
    x = np.zeros(n); o8 = np.zeros(8); obj = ctypes.c_double(); it = ctypes.c_int()
    lam = np.zeros(6 * N); zl = np.zeros(n); zu = np.zeros(n); tr = np.zeros((3100, 10)); nr = ctypes.c_int()
    cwd = os.getcwd()
    os.chdir(os.path.join(ORACLE_DIR, "_ref"))  # a directory without ipopt.opt
    try:
        wv = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        assert wv is None or wv.shape == (7,)
        rc = ref().ref_mpc_solve_w(N, dt, Lf, ref_v, delta_max, a_max, _p(wv), _p(st), _p(c), len(c), opts.encode(), _p(x),
                                   _p(o8), ctypes.byref(obj), ctypes.byref(it), _p(lam), _p(zl), _p(zu), _p(tr), 3100,
                                   ctypes.byref(nr))
    finally:
        os.chdir(cwd)
    out = dict(status=rc, x=x, out8=o8, obj=obj.value, iters=it.value, lam=lam, zl=zl, zu=zu)
    if trace:
        out["trace"] = tr[:min(nr.value, 3100)].copy()
    return out


def ref_eval(x, lam, coeffs, sigma=1.0, N=25, dt=0.05, Lf=2.67, ref_v=40.0):
    n, m = 8 * N - 2, 6 * N
    x = np.ascontiguousarray(x, dtype=np.float64); lam = np.ascontiguousarray(lam, dtype=np.float64)
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    f = ctypes.c_double(); grad = np.zeros(n); g = np.zeros(m); J = np.zeros((m, n)); H = np.zeros((n, n))
    ref().ref_mpc_eval(N, dt, Lf, ref_v, _p(c), len(c), _p(x), _p(lam), sigma, ctypes.byref(f), _p(grad), _p(g), _p(J),
                       _p(H))
    return f.value, grad, g, J, H


def ref_polyfit(xs, ys, order):
    xs = np.ascontiguousarray(xs, dtype=np.float64); ys = np.ascontiguousarray(ys, dtype=np.float64)
    out = np.zeros(order + 1)
    ref_helpers().ref_polyfit(_p(xs), _p(ys), len(xs), order, _p(out))
    return out


def ref_polyeval(coeffs, x):
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    return ref_helpers().ref_polyeval(_p(c), len(c), float(x))


def ref_kinematic(state4, act2, dt):
    s = np.ascontiguousarray(state4, dtype=np.float64); a = np.ascontiguousarray(act2, dtype=np.float64)
    out = np.zeros(4)
    ref_helpers().ref_global_kinematic(_p(s), _p(a), dt, _p(out))
    return out
