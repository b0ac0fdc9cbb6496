"""ctypes binding of include/b200mpc.h and the reference-shaped Python wrappers around it."""
import ctypes
import os
import threading

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p


class B200MPCError(RuntimeError):
    pass


class MPCParams(ctypes.Structure):
    """b200mpc_params: the reference's hard-coded configuration (MPC.cpp:14-31, 57-76, 194-203) + Ipopt tol/max_iter."""
    _fields_ = [("N", ctypes.c_int), ("dt", ctypes.c_double), ("Lf", ctypes.c_double), ("ref_v", ctypes.c_double),
                ("w_cte", ctypes.c_double), ("w_epsi", ctypes.c_double), ("w_v", ctypes.c_double),
                ("w_delta", ctypes.c_double), ("w_a", ctypes.c_double), ("w_ddelta", ctypes.c_double),
                ("w_da", ctypes.c_double), ("delta_max", ctypes.c_double), ("a_max", ctypes.c_double),
                ("tol", ctypes.c_double), ("max_iter", ctypes.c_int)]


# every symbol include/b200mpc.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b200mpc_default_params": (None, [ctypes.POINTER(MPCParams)]),
    "b200mpc_create": (ctypes.c_int, [ctypes.POINTER(MPCParams), ctypes.c_int, ctypes.POINTER(_vp)]),
    "b200mpc_destroy": (None, [_vp]),
    "b200mpc_last_error": (ctypes.c_char_p, []),
    "b200mpc_num_vars": (ctypes.c_int, [_vp]),
    "b200mpc_solve_batch": (ctypes.c_int, [_vp, ctypes.c_int, _dp, _dp, ctypes.c_int, _dp, _dp, _dp, _ip, _ip]),
    "b200mpc_solve_batch_async": (ctypes.c_int, [_vp, ctypes.c_int, _dp, _dp, ctypes.c_int, _dp, _dp, _dp, _ip, _ip]),
    "b200mpc_wait": (ctypes.c_int, [_vp]),
    "b200mpc_solve_batch_device": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200mpc_solve_batch_multi": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int,
                                                 _dp, _dp, _dp, _ip, _ip]),
    "b200mpc_closed_loop_batch": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int, _dp, _dp, _ip]),
    "b200mpc_polyfit_batch": (ctypes.c_int, [_vp, ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_int, _dp]),
    "b200mpc_polyfit_batch_device": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    "b200mpc_polyeval_batch": (ctypes.c_int, [_vp, ctypes.c_int, _dp, ctypes.c_int, _dp, _dp]),
    "b200mpc_polyeval_batch_device": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp, _vp]),
    "b200mpc_rollout_batch": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_double, ctypes.c_double, _dp]),
    "b200mpc_rollout_batch_device": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_double,
                                                    ctypes.c_double, _vp, _vp]),
    "b200mpc_roadmap_reference_batch": (ctypes.c_int, [_vp, ctypes.c_int, _dp, _dp, ctypes.c_int, _dp, _dp]),
    "b200mpc_roadmap_reference_batch_device": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, _vp]),
    "b200mpc_read_roadmap_csv": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, _dp, _dp, ctypes.c_int, _ip]),
    "b200mpc_set_warm_start": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double]),
    "b200mpc_set_batch_split": (ctypes.c_int, [_vp, ctypes.c_int]),
    "b200mpc_set_compaction": (ctypes.c_int, [_vp, ctypes.c_double, ctypes.c_int]),
    "b200mpc_set_pipeline": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    "b200mpc_set_handover": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    "b200mpc_set_restoration": (ctypes.c_int, [_vp, ctypes.c_int]),
    "b200mpc_set_solver_mode": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "b200mpc_set_timing": (ctypes.c_int, [_vp, ctypes.c_int]),
    "b200mpc_kernel_time_ms": (ctypes.c_int, [_vp, _dp, _ip, ctypes.c_int]),
    "b200mpc_measure_fp64_peak": (ctypes.c_int, [_vp, _dp]),
    "b200mpc_selftest_division": (ctypes.c_int, [_vp, ctypes.c_int, _dp, _dp, _dp, _dp]),
    "b200mpc_launch_count": (ctypes.c_longlong, [_vp]),
}

_lib = None
_lib_lock = threading.Lock()


def lib_path():
    """In-tree library; B200MPC_LIB selects another build of the same library (kernel tuning experiments)."""
    return os.environ.get("B200MPC_LIB") or os.path.join(_PKG, "lib", "libb200mpc.so")


def load_library():
    """Loads libb200mpc.so (built in-tree by udacitympc_b200.build).  Raises if it is missing: no fallback."""
    global _lib
    with _lib_lock:
        if _lib is None:
            path = lib_path()
            if not os.path.exists(path):
                raise B200MPCError(
                    f"{path} not found: build it with `python -m udacitympc_b200.build` "
                    "(the CUDA library is the only compute path; there is no CPU fallback)")
            L = ctypes.CDLL(path)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise B200MPCError(f"b200mpc error {rc}: {load_library().b200mpc_last_error().decode()}")


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a):
    return a.ctypes.data_as(_dp)


def default_params(**kw):
    p = MPCParams()
    load_library().b200mpc_default_params(ctypes.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k}")
        setattr(p, k, v)
    return p


class MPC:
    """Drop-in for the reference's `class MPC` (mpc_to_line/src/MPC.h:7-17).

    `MPC().Solve(x0, coeffs)` returns `[x1, y1, psi1, v1, cte1, epsi1, delta0, a0]` exactly like
    solution/MPC.cpp:253-256.  Unlike the reference the configuration is not a set of edit-the-source globals
    (MPC.cpp:14-31) but keyword parameters with the same defaults, and batches are first class.
    Like the reference, Solve ignores the solver status (MPC.cpp:248-249); `last_status` exposes it."""

    def __init__(self, device=0, print_cost=False, full_polynomial=False, **params):
        self._lib = load_library()
        self.params = default_params(**params)
        self.device = device
        self.print_cost = print_cost
        # Solve(): the reference's FG_eval reads coeffs[0..1] only (MPC.cpp:117-118) whatever the vector's length, and
        # so does Solve() unless this extension is switched on (solve_batch always uses every column it is given)
        self.full_polynomial = full_polynomial
        h = _vp()
        _check(self._lib.b200mpc_create(ctypes.byref(self.params), device, ctypes.byref(h)))
        self._h = h
        self.last_status = None
        self.last_iters = None
        self.last_cost = None

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200mpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    @property
    def num_vars(self):
        return self._lib.b200mpc_num_vars(self._h)

    # -- the reference call
    def Solve(self, x0, coeffs):
        x0 = _f64(x0).ravel()
        coeffs = _f64(coeffs).ravel()
        if x0.size != 6:
            raise ValueError("state must have 6 entries (x, y, psi, v, cte, epsi)")
        if coeffs.size < 2:
            raise ValueError("coeffs must have at least 2 entries (MPC.cpp:117-118)")
        if not self.full_polynomial:
            coeffs = coeffs[:2]
        r = self.solve_batch(x0[None, :], coeffs[None, :], want_traj=False)
        self.last_status = int(r["status"][0])
        self.last_iters = int(r["iters"][0])
        self.last_cost = float(r["cost"][0])
        if self.print_cost:
            print(f"Cost {self.last_cost:g}")   # MPC.cpp:251-252
        return [float(v) for v in r["out8"][0]]

    # -- batched host-buffer call
    def solve_batch(self, states, coeffs, want_traj=False):
        """states (B,6), coeffs (B,ncoef) or (ncoef,) shared by all problems.  Returns dict(out8 (B,8), cost (B,),
        status (B,), iters (B,), traj (B,8N-2) if requested)."""
        states = _f64(states)
        if states.ndim != 2 or states.shape[1] != 6:
            raise ValueError("states must be (B, 6)")
        B = states.shape[0]
        coeffs = _f64(coeffs)
        if coeffs.ndim == 1:
            coeffs = np.ascontiguousarray(np.broadcast_to(coeffs, (B, coeffs.size)))
        if coeffs.shape[0] != B:
            raise ValueError("coeffs must be (B, ncoef)")
        ncoef = coeffs.shape[1]
        out8 = np.empty((B, 8))
        cost = np.empty(B)
        status = np.empty(B, dtype=np.int32)
        iters = np.empty(B, dtype=np.int32)
        traj = np.empty((B, self.num_vars)) if want_traj else None
        _check(self._lib.b200mpc_solve_batch(self._h, B, _ptr(states), _ptr(coeffs), ncoef, _ptr(out8),
                                             _ptr(traj) if want_traj else None, _ptr(cost),
                                             status.ctypes.data_as(_ip), iters.ctypes.data_as(_ip)))
        r = dict(out8=out8, cost=cost, status=status, iters=iters)
        if want_traj:
            r["traj"] = traj
        return r

    def solve_batch_device(self, B, d_state6, d_coeffs, ncoef, d_out8, d_traj=0, d_obj=0, d_status=0, d_iters=0,
                           stream=0):
        """Raw device-pointer call (integers), field-major buffers; asynchronous on `stream`."""
        _check(self._lib.b200mpc_solve_batch_device(self._h, B, d_state6, d_coeffs, ncoef, d_out8, d_traj or None,
                                                    d_obj or None, d_status or None, d_iters or None, stream or None))

    def wait(self):
        """Waits for everything queued on this handle's own stream (b200mpc_solve_batch_async calls)."""
        _check(self._lib.b200mpc_wait(self._h))

    def closed_loop(self, states, coeffs, steps):
        """solution/main.cpp:51-76 for B vehicles: returns dict(hist8 (steps,B,8), cost (steps,B), iters (steps,B))."""
        states = _f64(states)
        if states.ndim == 1:
            states = states[None, :]
        B = states.shape[0]
        coeffs = _f64(coeffs)
        if coeffs.ndim == 1:
            coeffs = np.ascontiguousarray(np.broadcast_to(coeffs, (B, coeffs.size)))
        hist = np.empty((steps, B, 8))
        cost = np.empty((steps, B))
        iters = np.empty((steps, B), dtype=np.int32)
        _check(self._lib.b200mpc_closed_loop_batch(self._h, B, steps, _ptr(states), _ptr(coeffs), coeffs.shape[1],
                                                   _ptr(hist), _ptr(cost), iters.ctypes.data_as(_ip)))
        return dict(hist8=hist, cost=cost, iters=iters)

    def set_solver_mode(self, mode=0, rounds=0, fused_below=-1):
        """mode 0 = per-pass kernels (default), 1 = fused kernel; see include/b200mpc.h."""
        _check(self._lib.b200mpc_set_solver_mode(self._h, mode, rounds, fused_below))

    def set_batch_split(self, parts):
        """Cut large batches into `parts` (1..4) sub-batches that run concurrently inside one call (default 4; use 1
        when the caller overlaps several calls itself)."""
        _check(self._lib.b200mpc_set_batch_split(self._h, int(parts)))

    def set_restoration(self, mode=2):
        """What follows a failed line search: 0 = status -2 at the iteration where the reference's Ipopt enters its
        restoration phase, 1 = the restoration step, 2 (the library default) = Ipopt's soft restoration phase first,
        then the restoration step.  See b200mpc_set_restoration."""
        _check(self._lib.b200mpc_set_restoration(self._h, int(mode)))

    def set_compaction(self, max_live_fraction=0.7, from_round=4):
        """Throughput path: after every round from `from_round` on, move the unfinished problems to consecutive
        workspace slots when they fill at most this fraction of the occupied slots (0 = off)."""
        _check(self._lib.b200mpc_set_compaction(self._h, float(max_live_fraction), int(from_round)))

    def set_handover(self, occupied_slots=1184, from_round=14):
        """Occupied workspace slots at which the cooperative kernel takes the rest of a batch over (default 1184: best for
        one call at a time; 64-256 for a caller that overlaps several calls).  See b200mpc_set_handover."""
        _check(self._lib.b200mpc_set_handover(self._h, int(occupied_slots), int(from_round)))

    def set_pipeline(self, depth=8, tail_slots=4096):
        """Pipelined solves: solve_batch_device calls issued on different streams overlap on this one handle -- the
        bulk of a batch runs in the full-size workspace, its last `tail_slots` unfinished problems finish in one of
        `depth` small tail contexts while the next batch's bulk runs (0 = off).  See b200mpc_set_pipeline."""
        _check(self._lib.b200mpc_set_pipeline(self._h, int(depth), int(tail_slots)))

    def set_warm_start(self, enable=True, mu_init=1e-4):
        """closed_loop() only: steps after the first start from the shifted previous solution (not reference behaviour)."""
        _check(self._lib.b200mpc_set_warm_start(self._h, int(enable), float(mu_init)))

    # -- measurement helpers
    def set_timing(self, enable=True):
        """Bracket every solve with CUDA events (read them with kernel_time_ms); off by default."""
        _check(self._lib.b200mpc_set_timing(self._h, int(enable)))

    def kernel_time_ms(self, reset=True):
        t = ctypes.c_double()
        n = ctypes.c_int()
        _check(self._lib.b200mpc_kernel_time_ms(self._h, ctypes.byref(t), ctypes.byref(n), int(reset)))
        return t.value, n.value

    def fp64_peak_tflops(self):
        t = ctypes.c_double()
        _check(self._lib.b200mpc_measure_fp64_peak(self._h, ctypes.byref(t)))
        return t.value

    def selftest_division(self, a, b):
        """(a / b, 1 / b) as the solver's sweeps compute them on the device (b200mpc_selftest_division)."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert a.shape == b.shape and a.ndim == 1
        q, r = np.empty_like(a), np.empty_like(a)
        _check(self._lib.b200mpc_selftest_division(self._h, len(a), a.ctypes.data_as(_dp), b.ctypes.data_as(_dp),
                                                   q.ctypes.data_as(_dp), r.ctypes.data_as(_dp)))
        return q, r

    def launch_count(self):
        return int(self._lib.b200mpc_launch_count(self._h))


def solve_batch_multi(mpcs, states, coeffs, want_traj=False):
    """One batch sharded by contiguous index ranges over several MPC handles (one per device)."""
    lib = load_library()
    states = _f64(states)
    B = states.shape[0]
    coeffs = _f64(coeffs)
    if coeffs.ndim == 1:
        coeffs = np.ascontiguousarray(np.broadcast_to(coeffs, (B, coeffs.size)))
    hs = (_vp * len(mpcs))(*[m.handle for m in mpcs])
    out8 = np.empty((B, 8)); cost = np.empty(B)
    status = np.empty(B, dtype=np.int32); iters = np.empty(B, dtype=np.int32)
    traj = np.empty((B, mpcs[0].num_vars)) if want_traj else None
    _check(lib.b200mpc_solve_batch_multi(hs, len(mpcs), B, _ptr(states), _ptr(coeffs), coeffs.shape[1], _ptr(out8),
                                         _ptr(traj) if want_traj else None, _ptr(cost), status.ctypes.data_as(_ip),
                                         iters.ctypes.data_as(_ip)))
    r = dict(out8=out8, cost=cost, status=status, iters=iters)
    if want_traj:
        r["traj"] = traj
    return r


# ---------------------------------------------------------------------------------------------
# helpers.h / globalKinematic.  They need a device handle; a process-wide default one is created lazily.
_default = None


def _default_mpc():
    global _default
    if _default is None:
        _default = MPC()
    return _default


def polyfit_batch(xs, ys, order, mpc=None):
    """B fits at once: xs, ys (B, m) -> coeffs (B, order+1), ascending powers."""
    m_ = mpc or _default_mpc()
    xs = _f64(xs); ys = _f64(ys)
    if xs.shape != ys.shape or xs.ndim != 2:
        raise ValueError("xs and ys must both be (B, m)")
    B, m = xs.shape
    out = np.empty((B, order + 1))
    _check(m_._lib.b200mpc_polyfit_batch(m_.handle, B, _ptr(xs), _ptr(ys), m, order, _ptr(out)))
    return out


def polyfit(xvals, yvals, order, mpc=None):
    """helpers.h:24-44.  Like the reference's assert (helpers.h:25-26) sizes must match and 1 <= order <= m-1."""
    xvals = _f64(xvals).ravel(); yvals = _f64(yvals).ravel()
    assert xvals.size == yvals.size
    assert 1 <= order <= xvals.size - 1
    return polyfit_batch(xvals[None, :], yvals[None, :], order, mpc)[0]


def polyeval_batch(coeffs, x, mpc=None):
    m_ = mpc or _default_mpc()
    coeffs = _f64(coeffs); x = _f64(x).ravel()
    B = x.size
    if coeffs.ndim == 1:
        coeffs = np.ascontiguousarray(np.broadcast_to(coeffs, (B, coeffs.size)))
    y = np.empty(B)
    _check(m_._lib.b200mpc_polyeval_batch(m_.handle, B, _ptr(coeffs), coeffs.shape[1], _ptr(x), _ptr(y)))
    return y


def polyeval(coeffs, x, mpc=None):
    """helpers.h:13-19."""
    return float(polyeval_batch(_f64(coeffs).ravel(), [x], mpc)[0])


def rollout_batch(states, actuators, dt, Lf=2.0, mpc=None):
    """states (B,4), actuators (B,H,2) -> (B,H,4): the state after each of H Euler steps."""
    m_ = mpc or _default_mpc()
    states = _f64(states); actuators = _f64(actuators)
    if actuators.ndim == 2:
        actuators = actuators[:, None, :]
    B, H = actuators.shape[0], actuators.shape[1]
    if states.shape != (B, 4) or actuators.shape[2] != 2:
        raise ValueError("states must be (B,4) and actuators (B,H,2)")
    actuators = np.ascontiguousarray(actuators)
    out = np.empty((B, H, 4))
    _check(m_._lib.b200mpc_rollout_batch(m_.handle, B, H, _ptr(states), _ptr(actuators), float(dt), float(Lf), _ptr(out)))
    return out


def roadmap_reference_batch(poses, centerline, mpc=None):
    """Roadmap front-end: poses (B,4) = (x, y, psi, v) in the road's global frame, centerline (n_wp,2) ->
    (state6 (B,6), coeffs (B,4)) in the vehicle frame, ready for MPC.solve_batch."""
    m_ = mpc or _default_mpc()
    poses = _f64(poses); centerline = _f64(centerline)
    if poses.ndim != 2 or poses.shape[1] != 4 or centerline.ndim != 2 or centerline.shape[1] != 2:
        raise ValueError("poses must be (B,4) and centerline (n_wp,2)")
    B = poses.shape[0]
    st = np.empty((B, 6)); cf = np.empty((B, 4))
    _check(m_._lib.b200mpc_roadmap_reference_batch(m_.handle, B, _ptr(poses), _ptr(centerline), centerline.shape[0], _ptr(st), _ptr(cf)))
    return st, cf


def read_roadmap_csv(path, float_fields=False):
    """The reference's roadmap file (7 numbers per line) -> (centerline (n_wp, 2), slope (n_wp,)), parsed by the library's
    host-side reader (b200mpc_read_roadmap_csv).  float_fields=True converts every field through single precision like
    the reference's std::stof (mpc_to_line/src/custom_MPC.h:35-44)."""
    lib = load_library()
    n = ctypes.c_int()
    _check(lib.b200mpc_read_roadmap_csv(os.fsencode(path), int(float_fields), None, None, 0, ctypes.byref(n)))
    cl = np.empty((n.value, 2)); sl = np.empty(n.value)
    _check(lib.b200mpc_read_roadmap_csv(os.fsencode(path), int(float_fields), _ptr(cl), _ptr(sl), n.value, ctypes.byref(n)))
    return cl, sl


def global_kinematic(state, actuators, dt, Lf=2.0, mpc=None):
    """global_kinematic_model/solution/main.cpp:36-62 (Lf = 2 there, :15)."""
    return rollout_batch(_f64(state).reshape(1, 4), _f64(actuators).reshape(1, 1, 2), dt, Lf, mpc)[0, 0]
