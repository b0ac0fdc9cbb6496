"""-m "not gpu": the C-ABI library loads, exports every symbol include/b200mpc.h declares, validates arguments,
and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import udacitympc_b200 as mp
from udacitympc_b200 import api
from conftest import ROOT, has_gpu


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200mpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mpc_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(mp.lib_path())
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_default_params_are_the_reference_values():
    p = api.default_params()
    assert (p.N, p.dt, p.Lf, p.ref_v) == (25, 0.05, 2.67, 40.0)           # MPC.cpp:14-15,27,31
    assert (p.delta_max, p.a_max) == (0.436332, 1.0)                      # MPC.cpp:194-203
    assert (p.w_cte, p.w_epsi, p.w_v, p.w_delta, p.w_a, p.w_ddelta, p.w_da) == (1.0,) * 7
    assert (p.tol, p.max_iter) == (1e-8, 3000)


def test_argument_errors_do_not_need_a_gpu():
    lib = mp.load_library()
    h = ctypes.c_void_p()
    p = api.default_params(N=1)
    assert lib.b200mpc_create(ctypes.byref(p), 0, ctypes.byref(h)) == -1
    assert b"N must be" in lib.b200mpc_last_error()
    assert lib.b200mpc_create(None, 0, ctypes.byref(h)) == -1
    assert lib.b200mpc_solve_batch(None, 1, None, None, 2, None, None, None, None, None) == -1
    assert lib.b200mpc_polyfit_batch(None, 1, None, None, 6, 3, None) == -1
    assert lib.b200mpc_num_vars(None) == 0
    for setter, args in (("b200mpc_set_restoration", (1,)), ("b200mpc_set_batch_split", (2,)), ("b200mpc_set_compaction", (0.7, 4)),
                         ("b200mpc_set_warm_start", (1, 1e-4)), ("b200mpc_set_pipeline", (4, 4096)), ("b200mpc_set_handover", (256, 14))):
        assert getattr(lib, setter)(None, *args) == -1 and b"null handle" in lib.b200mpc_last_error()


@pytest.mark.skipif(has_gpu(), reason="this container check only makes sense without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(mp.B200MPCError) as e:
        mp.MPC()
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(mp.B200MPCError):
        mp.polyfit([0.0, 1.0], [0.0, 1.0], 1)


def test_product_does_not_reference_oracle():
    """Nothing under udacitympc_b200/ or include/ may import, link or execute oracle/ or tests/hostsim."""
    bad = []
    for base in ("udacitympc_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f)).read()
                    if re.search(r"oracle_bindings|libmpc_oracle|mpc_oracle\.h|libhostsim|oracle/_ref|libmpc_ref", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
    out = os.popen(f"ldd {mp.lib_path()}").read()
    assert "oracle" not in out and "ipopt" not in out.lower()


def test_synth_rng_is_mt19937_64():
    from udacitympc_b200.synth import MT19937_64
    r = MT19937_64(5489)
    assert int(r.raw(10000)[-1]) == 9981545732273789042   # the C++ standard's check value for std::mt19937_64
    u = MT19937_64(20261018).uniform(5)
    assert np.all((u >= 0) & (u < 1))


def test_roadmap_csv_reader_needs_no_gpu(tmp_path):
    """b200mpc_read_roadmap_csv: host-only parse of the reference's 7-column roadmap file (custom_MPC.h:25-44)."""
    path = os.path.join(ROOT, "udacitympc_b200", "data", "roadmap.csv")
    cl, slope = mp.read_roadmap_csv(path)
    raw = np.loadtxt(path, delimiter=",")
    assert cl.shape == (246, 2) and np.array_equal(cl, raw[:, 4:6]) and np.array_equal(slope, raw[:, 6])
    from udacitympc_b200 import synth
    assert np.array_equal(cl, synth.roadmap_centerline())
    # the reference converts every field with std::stof: single precision
    clf, slf = mp.read_roadmap_csv(path, float_fields=True)
    assert np.array_equal(clf, raw[:, 4:6].astype(np.float32).astype(np.float64)) and not np.array_equal(clf, cl)
    bad = tmp_path / "bad.csv"
    bad.write_text("1,2,3,4,5,6,7\n1,2,3\n")
    with pytest.raises(mp.B200MPCError) as e:
        mp.read_roadmap_csv(str(bad))
    assert "bad.csv:2" in str(e.value) and "7" in str(e.value)
    bad.write_text("1,2,x,4,5,6,7\n")
    with pytest.raises(mp.B200MPCError):
        mp.read_roadmap_csv(str(bad))
    with pytest.raises(mp.B200MPCError):
        mp.read_roadmap_csv(str(tmp_path / "missing.csv"))
