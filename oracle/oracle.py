"""ctypes binding of the CPU oracle (oracle/lbm_oracle.c).

TEST INFRASTRUCTURE.  Import only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product (lbm_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BGK, TRT, MRT = 0, 1, 2


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liblbm_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liblbm_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        i64, dbl, vp, i32 = C.c_int64, C.c_double, C.c_void_p, C.c_int
        pi64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        pdbl = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i32, i32, i64, pi64, i32, dbl]
        L.orc_set_geometry.argtypes = [vp, pdbl, pdbl, pdbl, dbl]
        L.orc_set_collision.argtypes = [vp, i32, dbl, C.c_void_p]
        L.orc_set_omp_collide.argtypes = [vp, i32]
        L.orc_add_bc_wall_bb.argtypes = [vp, pi64, pdbl, i64, dbl]
        L.orc_add_bc_dirichlet_bb.argtypes = [vp, pi64, pdbl, i64, pdbl]
        L.orc_add_bc_pressure.argtypes = [vp, pi64, pdbl, i64, dbl]
        L.orc_add_bc_periodic.argtypes = [vp, pi64, pdbl, i64, pi64, i64, dbl]
        L.orc_set_forcing.argtypes = [vp, pi64, i64, pi64, i64, dbl]
        L.orc_add_bc_wall_wetnode.argtypes = [vp, i32, pi64, pdbl, i64, i32, pdbl]
        L.orc_set_poisson.argtypes = [vp, dbl, dbl]
        L.orc_add_bc_poisson_neem.argtypes = [vp, i32, pi64, pdbl, i64, pdbl, dbl]
        L.orc_destroy.argtypes = [vp]
        L.orc_init.argtypes = [vp]
        L.orc_step.argtypes = [vp, i64]
        L.orc_update_moments.argtypes = [vp]
        L.orc_step_collide.argtypes = [vp]
        L.orc_step_stream.argtypes = [vp]
        L.orc_residual.argtypes = [vp, pdbl]
        L.orc_residual.restype = i32
        for name in ("orc_f", "orc_fold", "orc_feq", "orc_vars", "orc_varsold"):
            getattr(L, name).restype = C.POINTER(C.c_double)
            getattr(L, name).argtypes = [vp]
        L.orc_threads.restype = i32
        _LIB = L
    return _LIB


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """One reference-semantics LBM solver on the CPU.  Method names mirror the C-ABI of the product."""

    def __init__(self, ndim, ndist, nghbr, omega):
        nghbr = _i64(nghbr)
        self.ndim, self.ndist, self.nvar = ndim, ndist, ndim + 1
        self.n = nghbr.shape[0]
        self._h = lib().orc_create(ndim, ndist, self.n, nghbr, nghbr.shape[1], float(omega))
        if not self._h:
            raise ValueError(f"unsupported lattice D{ndim}Q{ndist}")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_geometry(self, center, bbmin, bbmax, cell_length):
        lib().orc_set_geometry(self._h, _f64(center), _f64(bbmin), _f64(bbmax), float(cell_length))

    def set_collision(self, model, omega_minus=0.0, rates=None):
        r = None if rates is None else _f64(rates).ctypes.data_as(C.c_void_p)
        self._rates = rates
        lib().orc_set_collision(self._h, model, float(omega_minus), r)

    def set_omp_collide(self, on):
        lib().orc_set_omp_collide(self._h, int(on))

    def add_wall_bb(self, cells, normals, tangential=0.0):
        rc = lib().orc_add_bc_wall_bb(self._h, _i64(cells), _f64(normals), len(cells), float(tangential))
        if rc != 0:
            raise ValueError("tangential wall velocity is 2D only (bnd_wall.h:52-54)")

    def add_dirichlet_bb(self, cells, normals, value):
        lib().orc_add_bc_dirichlet_bb(self._h, _i64(cells), _f64(normals), len(cells), _f64(value))

    def add_pressure(self, cells, normals, pressure):
        lib().orc_add_bc_pressure(self._h, _i64(cells), _f64(normals), len(cells), float(pressure))

    def add_periodic(self, cells, normals, connected, pressure=float("nan")):
        rc = lib().orc_add_bc_periodic(self._h, _i64(cells), _f64(normals), len(cells), _i64(connected), len(connected),
                                       float(pressure))
        if rc != 0:
            raise ValueError("periodic BC needs set_geometry first")

    def add_wall_wetnode(self, model, cells, normals, velocity=None):
        kind = {"equilibrium": 6, "neem": 7, "nebb": 8}[model]
        v = _f64(np.zeros(self.ndim) if velocity is None else velocity)
        rc = lib().orc_add_bc_wall_wetnode(self._h, kind, _i64(cells), _f64(normals), len(cells), int(velocity is not None), v)
        if rc == -1:
            raise ValueError("Not implemented for this distribution!")
        if rc == -2:
            raise ValueError("No valid extrapolation cellId")

    def set_poisson(self, dt, rate):
        """switch to the Poisson equation (one variable: the potential); call before adding boundary conditions"""
        lib().orc_set_poisson(self._h, float(dt), float(rate))
        self.nvar = 1

    def add_poisson_neem(self, kind, cells, normals, values, grad=0.0):
        rc = lib().orc_add_bc_poisson_neem(self._h, int(kind == "neumann"), _i64(cells), _f64(normals), len(cells), _f64(values), float(grad))
        if rc != 0:
            raise ValueError("No valid extrapolation cellId")

    def set_forcing(self, inlet, outlet, gradient):
        rc = lib().orc_set_forcing(self._h, _i64(inlet), len(inlet), _i64(outlet), len(outlet), float(gradient))
        if rc != 0:
            raise ValueError("forcing needs set_geometry first")

    def init(self):
        lib().orc_init(self._h)

    def step(self, n=1):
        lib().orc_step(self._h, int(n))

    def step_collide(self):
        lib().orc_step_collide(self._h)

    def step_stream(self):
        lib().orc_step_stream(self._h)

    def update_moments(self):
        lib().orc_update_moments(self._h)

    def residual(self):
        out = np.zeros(self.nvar)
        bad = lib().orc_residual(self._h, out)
        return out, bool(bad)

    def _view(self, fn, width):
        p = getattr(lib(), fn)(self._h)
        return np.ctypeslib.as_array(p, shape=(self.n, width))

    @property
    def f(self):
        return self._view("orc_f", self.ndist)

    @property
    def fold(self):
        return self._view("orc_fold", self.ndist)

    @property
    def feq(self):
        return self._view("orc_feq", self.ndist)

    @property
    def vars(self):
        return self._view("orc_vars", self.nvar)

    @property
    def varsold(self):
        return self._view("orc_varsold", self.nvar)


def threads():
    return lib().orc_threads()
