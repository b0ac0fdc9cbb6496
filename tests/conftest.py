import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session", autouse=True)
def _build_everything():
    """Builds the test-only oracle (+ oracle/_ref when /root/reference is present), the host build of the solver
    core used by the CPU algorithm tests, and the product library."""
    import __graft_entry__ as ge
    ge.build()
    yield


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


# ---- host build of the solver core (tests/hostsim): algorithm tests without a GPU
class _HostSim:
    def __init__(self, libname="libhostsim.so"):
        import oracle_bindings as ob
        self.ob = ob
        self.lib = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", libname))
        self.dp = ctypes.POINTER(ctypes.c_double)

    def solve(self, state6, coeffs, mode=0, **params):
        ob, dp = self.ob, self.dp
        p = ob.default_params(**params)
        N = p.N
        st = np.ascontiguousarray(state6, dtype=np.float64)
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        x = np.zeros(8 * N - 2); o8 = np.zeros(8); obj = ctypes.c_double(); it = ctypes.c_int()
        lam = np.zeros(6 * N); tr = np.zeros((400, 8)); nr = ctypes.c_int()
        rc = self.lib.hostsim_solve_mode(ctypes.byref(p), st.ctypes.data_as(dp), c.ctypes.data_as(dp), len(c),
                                         x.ctypes.data_as(dp), o8.ctypes.data_as(dp), ctypes.byref(obj), ctypes.byref(it),
                                         lam.ctypes.data_as(dp), tr.ctypes.data_as(dp), 400, ctypes.byref(nr), mode)
        return dict(status=rc, x=x, out8=o8, obj=obj.value, iters=it.value, lam=lam, trace=tr[:nr.value])

    def batch_interleaved(self, states, coeffs, compact=True, max_rounds=0, coop=False, **params):
        """B problems in the device memory layout (Solver<32>, two regions, per-pass execution, batch compaction)."""
        ob, dp = self.ob, self.dp
        p = ob.default_params(**params)
        st = np.ascontiguousarray(states, dtype=np.float64)
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        B = len(st)
        out8 = np.zeros((B, 8)); obj = np.zeros(B); it = np.zeros(B, dtype=np.int32); status = np.zeros(B, dtype=np.int32)
        ip = ctypes.POINTER(ctypes.c_int)
        rc = self.lib.hostsim_batch_interleaved(ctypes.byref(p), B, st.ctypes.data_as(dp), c.ctypes.data_as(dp), c.shape[1],
                                                1 if compact else 0, int(max_rounds), 1 if coop else 0, out8.ctypes.data_as(dp),
                                                obj.ctypes.data_as(dp),
                                                it.ctypes.data_as(ip), status.ctypes.data_as(ip))
        return dict(rc=rc, out8=out8, obj=obj, iters=it, status=status)


@pytest.fixture(scope="session")
def hostsim(_build_everything):
    return _HostSim()


@pytest.fixture(scope="session")
def hostsim_fuse(_build_everything):
    """The solver core compiled with MPC_FUSE_FACTOR=1 (the fused step + factor sweep, an experiment the product leaves out)."""
    return _HostSim("libhostsim_fuse.so")
