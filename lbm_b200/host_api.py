"""ctypes binding of the host-side mirror (lbm_b200/host: grid pipeline + LBMSolver), used by the tests."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "host")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liblbm_host.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp = C.c_void_p
        L.lbmhost_grid_build.restype = vp
        L.lbmhost_grid_build.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.lbmhost_grid_free.argtypes = [vp]
        L.lbmhost_grid_ncells.restype = C.c_int64
        L.lbmhost_grid_ncells.argtypes = [vp]
        L.lbmhost_grid_ndim.argtypes = [vp]
        L.lbmhost_grid_stride.argtypes = [vp]
        L.lbmhost_grid_cell_length.restype = C.c_double
        L.lbmhost_grid_cell_length.argtypes = [vp]
        L.lbmhost_grid_copy.argtypes = [vp, vp, vp, vp]
        L.lbmhost_grid_nsurfaces.argtypes = [vp]
        L.lbmhost_grid_surface_name.restype = C.c_char_p
        L.lbmhost_grid_surface_name.argtypes = [vp, C.c_int]
        L.lbmhost_grid_surface_size.restype = C.c_int64
        L.lbmhost_grid_surface_size.argtypes = [vp, C.c_int]
        L.lbmhost_grid_surface_copy.argtypes = [vp, C.c_int, vp, vp]
        L.lbmhost_postprocess_line.restype = C.c_int64
        L.lbmhost_postprocess_line.argtypes = [vp, vp, C.c_char_p, C.c_char_p, C.c_int]
        L.lbmhost_run.argtypes = [C.c_char_p, vp, vp, C.c_int64, C.c_char_p, C.c_int]
        L.lbmhost_ugrid_build.restype = vp
        L.lbmhost_ugrid_build.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.lbmhost_ugrid_free.argtypes = [vp]
        L.lbmhost_ugrid_ncells.restype = C.c_int64
        L.lbmhost_ugrid_ncells.argtypes = [vp]
        for fn in ("ndim", "stride", "level"):
            getattr(L, f"lbmhost_ugrid_{fn}").argtypes = [vp]
        L.lbmhost_ugrid_cell_length.restype = C.c_double
        L.lbmhost_ugrid_cell_length.argtypes = [vp]
        L.lbmhost_ugrid_bbox.argtypes = [vp, vp, vp]
        L.lbmhost_ugrid_rows.argtypes = [vp, vp, C.c_int64, vp, C.c_int, vp, C.c_char_p, C.c_int]
        L.lbmhost_ugrid_sources.argtypes = [vp, vp, C.c_int64, vp, C.c_int, C.c_char_p, C.c_int]
        L.lbmhost_ugrid_nsurfaces.argtypes = [vp, C.c_char_p, C.c_int]
        L.lbmhost_ugrid_surface_name.restype = C.c_char_p
        L.lbmhost_ugrid_surface_name.argtypes = [vp, C.c_int]
        L.lbmhost_ugrid_surface_size.restype = C.c_int64
        L.lbmhost_ugrid_surface_size.argtypes = [vp, C.c_int]
        L.lbmhost_ugrid_surface_copy.argtypes = [vp, C.c_int, vp, vp]
        L.lbmhost_eval_expression.argtypes = [C.c_char_p, vp, C.c_int64, C.c_int, vp, C.c_char_p, C.c_int]
        _LIB = L
    return _LIB


def build_grid(config_path):
    """Grid generator run + transferGrid of the host mirror -> dict of tables in the reference's format."""
    L = lib()
    err = C.create_string_buffer(1024)
    h = L.lbmhost_grid_build(config_path.encode(), err, 1024)
    if not h:
        raise RuntimeError(err.value.decode())
    try:
        n, ndim, stride = L.lbmhost_grid_ncells(h), L.lbmhost_grid_ndim(h), L.lbmhost_grid_stride(h)
        nghbr = np.empty((n, stride), dtype=np.int64)
        center = np.empty((n, ndim))
        props = np.empty(n, dtype=np.uint16)
        L.lbmhost_grid_copy(h, nghbr.ctypes.data, center.ctypes.data, props.ctypes.data)
        surfaces = []
        for k in range(L.lbmhost_grid_nsurfaces(h)):
            m = L.lbmhost_grid_surface_size(h, k)
            cells = np.empty(m, dtype=np.int64)
            normals = np.empty((m, ndim))
            if m:
                L.lbmhost_grid_surface_copy(h, k, cells.ctypes.data, normals.ctypes.data)
            surfaces.append((L.lbmhost_grid_surface_name(h, k).decode(), cells, normals))
        return dict(n=n, ndim=ndim, nghbr=nghbr, center=center, props=props, surfaces=surfaces,
                    cell_length=L.lbmhost_grid_cell_length(h))
    finally:
        L.lbmhost_grid_free(h)


def postprocess_line(config_path, vars_, out_path):
    """What the reference's postprocessing function "line" (hook atEnd) writes to ./line.csv, from caller-supplied m_vars
    ([ncells][ndim+1]); returns the number of cells on the line.  Host-side only (grid pipeline + formatting), no GPU."""
    L = lib()
    err = C.create_string_buffer(1024)
    h = L.lbmhost_grid_build(config_path.encode(), err, 1024)
    if not h:
        raise RuntimeError(err.value.decode())
    try:
        v = np.ascontiguousarray(vars_, dtype=np.float64)
        n = L.lbmhost_grid_ncells(h)
        assert v.size in (n, n * (L.lbmhost_grid_ndim(h) + 1))  # one variable per cell for the Poisson equation types
        n = L.lbmhost_postprocess_line(h, v.ctypes.data, out_path.encode(), err, 1024)
        if n < 0:
            raise RuntimeError(err.value.decode())
        return int(n)
    finally:
        L.lbmhost_grid_free(h)


def run(config_path, nvars=0):
    """Whole pipeline like the `lbm` executable. Returns (rc, message, dict of results, vars or None)."""
    L = lib()
    err = C.create_string_buffer(2048)
    out = np.zeros(6)
    vars_ = np.zeros(nvars) if nvars else None
    rc = L.lbmhost_run(config_path.encode(), out.ctypes.data, vars_.ctypes.data if nvars else None, nvars, err, 2048)
    keys = ["max_error", "l2_error", "gre", "steps", "converged", "residual"]
    return rc, err.value.decode(), dict(zip(keys, out.tolist())), vars_


def partitioned_plan(config_path, rank, world, ndim, ndist):
    """The device plan of rank `rank` of a `world`-rank run as the C++ host sets it up (LBMSolver::setupGpuPartitioned), on an
    inspection-only handle -- for comparison with the Python path (lbm_b200.partition + lbm_b200.cases)."""
    from . import capi
    L = lib()
    L.lbmhost_partitioned_handle.restype = C.c_void_p
    L.lbmhost_partitioned_handle.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    err = C.create_string_buffer(1024)
    h = L.lbmhost_partitioned_handle(str(config_path).encode(), int(rank), int(world), err, 1024)
    if not h:
        raise RuntimeError(err.value.decode())
    s = capi.Solver.__new__(capi.Solver)
    s._lib, s._h = capi.load_library(), C.c_void_p(h)
    s.ndim, s.ndist, s.nvar, s.n = ndim, ndist, ndim + 1, 0
    try:
        return s.debug_plan()
    finally:
        s.close()


def eval_expression(text, points):
    """A boundary-value expression of a configuration ("value": "cos(pi*x)") at points [n, ndim] (lbm_b200/host/expr.hpp)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    out = np.empty(len(pts))
    err = C.create_string_buffer(512)
    if lib().lbmhost_eval_expression(text.encode(), pts.ctypes.data, len(pts), pts.shape[1], out.ctypes.data, err, 512) != 0:
        raise ValueError(err.value.decode())
    return out


class UniformGrid:
    """Single-level grid of a configuration with table rows ON DEMAND (lbm_b200/host/uniform_grid.hpp): what one rank of a partitioned
    run asks the grid pipeline.  Same cell order, neighbour rows, centres and boundary surfaces as build_grid() on the same
    configuration (tests/test_uniform_grid.py), without ever holding a table over the whole domain."""

    def __init__(self, config_path):
        self._L = lib()
        err = C.create_string_buffer(1024)
        self._h = self._L.lbmhost_ugrid_build(str(config_path).encode(), err, 1024)
        if not self._h:
            raise RuntimeError(err.value.decode())
        L, h = self._L, self._h
        self.n, self.ndim, self.stride, self.level = L.lbmhost_ugrid_ncells(h), L.lbmhost_ugrid_ndim(h), L.lbmhost_ugrid_stride(h), L.lbmhost_ugrid_level(h)
        self.cell_length = L.lbmhost_ugrid_cell_length(h)
        self.bbmin, self.bbmax = np.zeros(self.ndim), np.zeros(self.ndim)
        L.lbmhost_ugrid_bbox(h, self.bbmin.ctypes.data, self.bbmax.ctypes.data)
        self._surfaces = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.lbmhost_ugrid_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _call(self, fn, ids, want_center=False):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        out = np.empty((len(ids), self.stride), dtype=np.int64)
        err = C.create_string_buffer(512)
        if fn == "rows":
            center = np.empty((len(ids), self.ndim)) if want_center else None
            rc = self._L.lbmhost_ugrid_rows(self._h, ids.ctypes.data, len(ids), out.ctypes.data, self.stride,
                                            center.ctypes.data if want_center else None, err, 512)
        else:
            center = None
            rc = self._L.lbmhost_ugrid_sources(self._h, ids.ctypes.data, len(ids), out.ctypes.data, self.stride, err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return out, center

    def rows(self, ids, want_center=False):
        """push-table rows [len(ids), stride] (global ids, -1 = none) and optionally the cell centres"""
        return self._call("rows", ids, want_center)

    def sources(self, ids):
        """pull sources: out[r, s] = the cell whose push in direction s lands in ids[r]"""
        return self._call("sources", ids)[0]

    def surfaces(self):
        """[(name, cells, normals)] in creation order, like build_grid()["surfaces"]"""
        if self._surfaces is None:
            err = C.create_string_buffer(512)
            ns = self._L.lbmhost_ugrid_nsurfaces(self._h, err, 512)
            if ns < 0:
                raise RuntimeError(err.value.decode())
            out = []
            for k in range(ns):
                m = self._L.lbmhost_ugrid_surface_size(self._h, k)
                cells = np.empty(m, dtype=np.int64)
                normals = np.empty((m, self.ndim))
                if m:
                    self._L.lbmhost_ugrid_surface_copy(self._h, k, cells.ctypes.data, normals.ctypes.data)
                out.append((self._L.lbmhost_ugrid_surface_name(self._h, k).decode(), cells, normals))
            self._surfaces = out
        return self._surfaces
