"""Device side of the output path (lbm_b200/csrc/output.cuh) executed on the CPU: the 15-decimal rounding and the
base64 text of a field must equal what the host writer produces, which tests/test_vtk_writer.py pins byte for byte against the reference
binary's files.  The GPU run of the same code is checked end to end by tests/test_host_run_gpu.py (SHA-256 of the solution files)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def oh(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("oh") / "liboutput_harness.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-Wall", "-Wno-unknown-pragmas",
                           "-Wno-unused-function", "-I", os.path.join(HERE, "c", "fake_cuda"), os.path.join(HERE, "c", "output_harness.cpp"), "-o", so])
    L = C.CDLL(so)
    L.oh_check_round15.restype = C.c_int64
    L.oh_check_round15.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.oh_check_base64.argtypes = [C.c_void_p, C.c_int64]
    for fn in (L.oh_check_column_f64, L.oh_check_column_f32):
        fn.restype = C.c_int64
        fn.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int)]
    return L


def test_round15_equals_the_host_writer(oh):
    rng = np.random.default_rng(5)
    x = np.concatenate([np.clip(rng.standard_normal(400_000) * 10.0 ** rng.uniform(-30, 0.5, 400_000), -9.0, 9.0),   # velocities, densities, tiny values
                        1.0 + rng.standard_normal(200_000) * 1e-4,                               # densities around one
                        rng.integers(-10 ** 15, 10 ** 15, 100_000) / 1e15,                       # exact multiples of 1e-15
                        (rng.integers(-10 ** 15, 10 ** 15, 100_000) + 0.5) / 1e15,               # ties
                        np.array([0.0, -0.0, 1e-16, -1e-16, 4.9999999999999999e-16, 5e-16, -5e-16, 1e-300, 5e-324, 9.0, 9.007199254740991])])
    x = np.ascontiguousarray(x)
    n_slow = C.c_int64()
    assert oh.oh_check_round15(x.ctypes.data, len(x), C.byref(n_slow)) == 0
    assert n_slow.value == 0
    big = np.ascontiguousarray(np.array([9.1, -20.0, 1e20, np.inf, np.nan]))                     # outside the exact integer range: flagged
    assert oh.oh_check_round15(big.ctypes.data, len(big), C.byref(n_slow)) == 0 and n_slow.value == len(big)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 640, 1001, 65536])
def test_base64_field_equals_the_host_writer(oh, n):
    col = np.ascontiguousarray(np.random.default_rng(n).standard_normal(n))
    assert oh.oh_check_base64(col.ctypes.data, n) == 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_output_column_gathers_the_kept_cells_and_flags_what_it_cannot_round(oh, dtype):
    """k_output_column: the kept cells (a selection list in reference order pointing at permuted device cells) of one variable of the
    [nvar][stride] moment array, rounded like the host writer; one value outside the exact range raises the flag for the whole call"""
    rng = np.random.default_rng(11)
    stride, nvar, n = 4096, 4, 1500
    src = np.ascontiguousarray((rng.standard_normal((nvar, stride)) * 0.1).astype(dtype))
    sel = np.ascontiguousarray(rng.permutation(stride)[:n].astype(np.int32))
    fn = oh.oh_check_column_f64 if dtype == np.float64 else oh.oh_check_column_f32
    slow = C.c_int(7)
    for var in range(nvar):
        assert fn(src.ctypes.data, stride, var, sel.ctypes.data, n, C.byref(slow)) == 0 and slow.value == 0
    src[2, sel[17]] = np.nan
    assert fn(src.ctypes.data, stride, 2, sel.ctypes.data, n, C.byref(slow)) == 0 and slow.value == 1
    assert fn(src.ctypes.data, stride, 1, sel.ctypes.data, n, C.byref(slow)) == 0 and slow.value == 0
