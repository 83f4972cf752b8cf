"""The Poisson GPU pipeline (lbm_b200/csrc/poisson.cuh) executed on the CPU: tests/c/poisson_harness.cpp compiles the SAME kernel bodies
and the SAME host set-up with g++ (thread indices walked by a loop, IEEE intrinsics as plain operators under -ffp-contract=off) and runs
them in PoissonSolver::step's launch order.  The result must equal the reference's dumps bit for bit on the five Poisson cases of its
test/run.sh -- so what tests/test_zzz_poisson_gpu.py still has to prove on hardware is only the CUDA plumbing around these bodies."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from casebuilder import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["poisson1D", "poisson1D_reaction", "poisson2D", "poisson2D_helmholtz", "poissonD2Q9", "poisson2D_reaction", "step_poisson"]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ph") / "libpoisson_harness.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
                           "-Wno-unused-variable", os.path.join(HERE, "c", "poisson_harness.cpp"), "-o", so])
    L = C.CDLL(so)
    L.ph_create.restype = C.c_void_p
    L.ph_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
    L.ph_destroy.argtypes = [C.c_void_p]
    L.ph_add_bc.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_double]
    L.ph_init.argtypes = [C.c_void_p]
    L.ph_step.argtypes = [C.c_void_p, C.c_int64]
    L.ph_potential.argtypes = [C.c_void_p, C.c_void_p]
    L.ph_array.restype = C.POINTER(C.c_double)
    L.ph_array.argtypes = [C.c_void_p, C.c_int]
    L.ph_error.restype = C.c_char_p
    L.ph_error.argtypes = [C.c_void_p]
    return L


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make(L, spec):
    nghbr = np.ascontiguousarray(spec.nghbr, dtype=np.int64)
    h = L.ph_create(spec.ndim, spec.ndist, spec.n, nghbr.ctypes.data, nghbr.shape[1], spec.omega, spec.poisson["dt"], spec.poisson["rate"])
    assert h
    for bc in spec.bcs:
        cells = np.ascontiguousarray(bc["cells"], dtype=np.int64)
        values = np.ascontiguousarray(bc["values"], dtype=np.float64)
        L.ph_add_bc(h, int(bc["neumann"]), cells.ctypes.data, len(cells), values.ctypes.data, bc["grad"])
    return h


@pytest.mark.parametrize("name", CASES)
def test_kernel_bodies_reproduce_the_reference_dump(name, harness, oracle_mod):
    L = harness
    spec = load_golden(name)
    h = make(L, spec)
    assert L.ph_init(h) == 0, L.ph_error(h)
    q = spec.ndist
    view = lambda which, width: np.ctypeslib.as_array(L.ph_array(h, which), shape=(spec.n, width))
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, q, spec.nghbr, spec.omega))
    o.init()
    assert np.array_equal(view(0, q), o.f) and np.array_equal(view(1, q), o.fold) and np.array_equal(view(2, 1), o.vars)
    done = 0
    for s in spec.golden["steps"]:
        L.ph_step(h, int(s) - done)
        done = int(s)
        for which, arr, width in ((1, "fold", q), (0, "f", q), (2, "vars", 1), (3, "varsold", 1)):
            assert sha(view(which, width)) == spec.digests[f"{arr}_{s}"], f"{name} step {s}: {arr} differs from the reference dump"
    pot = np.empty(spec.n)
    L.ph_potential(h, pot.ctypes.data)
    o.step(done)
    o.update_moments()
    assert np.array_equal(pot, o.vars[:, 0])
    L.ph_destroy(h)


def test_order_hazards_are_refused(harness):
    """an extrapolation cell that is itself an entry of the same surface: the reference's serial loop would see the rewritten cell"""
    L = harness
    nghbr = np.array([[-1, 1], [0, 2], [1, 3], [2, -1]], dtype=np.int64)
    h = L.ph_create(1, 3, 4, nghbr.ctypes.data, 2, 1.0, 0.25, 1.0)
    cells = np.array([0, 1], dtype=np.int64)  # cell 1 has both neighbours: no extrapolation direction
    vals = np.array([1.0, 1.0])
    L.ph_add_bc(h, 0, cells.ctypes.data, 2, vals.ctypes.data, 0.0)
    assert L.ph_init(h) == -1 and b"No valid extrapolation" in L.ph_error(h)
    L.ph_destroy(h)
    h = L.ph_create(1, 3, 4, nghbr.ctypes.data, 2, 1.0, 0.25, 1.0)
    cells = np.array([0, 0], dtype=np.int64)
    L.ph_add_bc(h, 0, cells.ctypes.data, 2, vals.ctypes.data, 0.0)
    assert L.ph_init(h) == -5 and b"listed twice" in L.ph_error(h)
    L.ph_destroy(h)
