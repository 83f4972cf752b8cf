"""ctypes binding of the C ABI (include/lbm_b200.h).

`Solver` mirrors the call sequence the reference's LBMSolver::run performs (src/lbm/solver.cpp:176-214):
create -> set_topology -> (set_geometry) -> add_* boundary conditions in application order -> (set_forcing) ->
init -> step / residual / read-back.  Errors surface as LbmB200Error carrying the library's message, the way the
reference surfaces them through TERMM (src/common/term.h:37).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None

FP64, FP32 = 0, 1
BGK, TRT, MRT = 0, 1, 2
STRICT, FAST = 0, 1
ABI_VERSION = 1


class LbmB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lbm_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("ndim", C.c_int32), ("ndist", C.c_int32), ("precision", C.c_int32),
                ("collision", C.c_int32), ("arithmetic", C.c_int32), ("device", C.c_int32), ("track_vars", C.c_int32),
                ("omega", C.c_double), ("omega_minus", C.c_double), ("mrt_rates", C.c_double * 27)]


class Stats(C.Structure):
    _fields_ = [("ncells", C.c_int64), ("cells_fast", C.c_int64), ("cells_generic", C.c_int64), ("chunk_cells", C.c_int64),
                ("slots_bc", C.c_int64), ("slots_stale", C.c_int64), ("device_bytes", C.c_int64), ("launches", C.c_int64),
                ("launches_main", C.c_int64), ("bytes_per_cell_alg", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("cells_ghost", C.c_int64), ("halo_bytes", C.c_int64)]


class PlanView(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n", "n_owned", "npad", "chunk", "nsel", "n_fast_chunks", "n_fast_outer", "gen_begin", "n_gen",
                                         "n_gen_outer", "gen_stride", "ghost_begin", "n_ghost_blocks", "n_values_static")] + [
        ("ref2dev", C.POINTER(C.c_int32)), ("tmpl", C.POINTER(C.c_uint16)), ("chunk_nb", C.POINTER(C.c_int32)),
        ("codes", C.POINTER(C.c_int32)), ("copytab", C.POINTER(C.c_int32)), ("n_copy", C.c_int64),
        ("addtab", C.POINTER(C.c_double)), ("n_add", C.c_int64), ("wall_desc", C.POINTER(C.c_double)), ("n_wall", C.c_int64),
        ("abb_p", C.POINTER(C.c_double)), ("abb_cells", C.POINTER(C.c_int32)), ("n_abb", C.c_int64),
        ("values", C.POINTER(C.c_double)), ("n_values", C.c_int64), ("stale_ref", C.POINTER(C.c_int64)), ("n_stale", C.c_int64),
        ("send_index", C.POINTER(C.c_int64)), ("n_send", C.c_int64), ("recv_index", C.POINTER(C.c_int64)), ("n_recv", C.c_int64),
        ("vsend_cells", C.POINTER(C.c_int32)), ("n_vsend", C.c_int64), ("n_vrecv", C.c_int64),
        ("chunk_abb_base", C.POINTER(C.c_int32)), ("chunk_abb", C.POINTER(C.c_int32)), ("n_chunk_abb_rows", C.c_int64),
        ("perm_end", C.c_int64), ("gb_begin", C.c_int64), ("gb_end", C.c_int64), ("layout", C.POINTER(C.c_int32))]


def mrt_moment_kinds(ndim, ndist):
    """kind of every moment of the MRT basis: 0 conserved, 1 shear, 2 bulk, 3 ghost (lbm_b200_mrt_moments)"""
    lib = load_library()
    kinds = (C.c_int32 * 27)()
    rc = lib.lbm_b200_mrt_moments(int(ndim), int(ndist), kinds)
    if rc != 0:
        raise LbmB200Error(rc, lib.lbm_b200_last_error().decode())
    return np.array(kinds[:ndist], dtype=np.int32)


def mrt_rates(ndim, ndist, shear, bulk=None, ghost=None):
    """mrt_rates array for Solver(collision=MRT): `shear` sets the viscosity (= the BGK omega of the same viscosity); bulk / ghost
    default to the shear rate (then the operator is BGK); ghost may be a sequence, one rate per ghost moment in basis order"""
    kinds = mrt_moment_kinds(ndim, ndist)
    rates = np.full(27, float(shear))
    rates[:ndist][kinds == 2] = float(shear if bulk is None else bulk)
    if ghost is not None:
        g = np.atleast_1d(np.asarray(ghost, dtype=np.float64))
        idx = np.nonzero(kinds == 3)[0]
        rates[idx] = g[np.arange(len(idx)) % len(g)]
    return rates


def pop_slot(plan, j, cells):
    """Position of device cells `cells` inside the population array of direction j (per-direction in-chunk layouts,
    include/lbm_b200.h: lbm_b200_plan_view.layout).  `plan` is the dict Solver.debug_plan() returns."""
    cells = np.asarray(cells, dtype=np.int64)
    lay = int(plan["layout"][j])
    if lay == 0:
        return cells.copy()
    inblock = (cells < plan["perm_end"]) | ((cells >= plan["gb_begin"]) & (cells < plan["gb_end"]))
    o = cells & 511
    perm = ((o >> 3) | ((o & 7) << 6)) if lay == 1 else ((o >> 6) | ((o & 63) << 3))
    return np.where(inblock, (cells & ~np.int64(511)) | perm, cells)


def pop_gather(plan, arr, cells):
    """arr: population array [Q, npad] in device layout -> [len(cells), Q] values of the device cells `cells`"""
    q = arr.shape[0]
    return np.stack([arr[j, pop_slot(plan, j, cells)] for j in range(q)], axis=1)


def pop_scatter(plan, arr, cells, values):
    """inverse of pop_gather: values [len(cells), Q] -> arr [Q, npad]"""
    for j in range(arr.shape[0]):
        arr[j, pop_slot(plan, j, cells)] = values[:, j]


class PartitionView(C.Structure):
    _fields_ = [("lo", C.c_int64), ("hi", C.c_int64), ("n_owned", C.c_int64), ("n_ghost", C.c_int64),
                ("ghosts", C.POINTER(C.c_int64)), ("nghbr", C.POINTER(C.c_int64)), ("stride", C.c_int32), ("npeers", C.c_int32),
                ("peers", C.POINTER(C.c_int32)),
                ("send_count", C.POINTER(C.c_int64)), ("recv_count", C.POINTER(C.c_int64)), ("send_cell", C.POINTER(C.c_int64)),
                ("recv_cell", C.POINTER(C.c_int64)), ("send_dir", C.POINTER(C.c_int32)), ("recv_dir", C.POINTER(C.c_int32)),
                ("vsend_count", C.POINTER(C.c_int64)), ("vrecv_count", C.POINTER(C.c_int64)), ("vsend_cell", C.POINTER(C.c_int64)),
                ("vrecv_cell", C.POINTER(C.c_int64))]


ROWS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64))


def library_path():
    # LBM_B200_LIB selects a tuning build of the same library (kernel experiments); never a different backend
    return os.environ.get("LBM_B200_LIB") or os.path.join(_HERE, "liblbm_b200.so")


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force and os.path.exists(library_path()):
        os.remove(library_path())
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")])
    return library_path()


def abi_symbols():
    """Every function include/lbm_b200.h declares (used by the CPU-side ABI test)."""
    text = open(os.path.join(_ROOT, "include", "lbm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_b200_[a-z_0-9]+)\s*\(", text)))


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LbmB200Error(-3, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    pi64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    pdbl = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    L.lbm_b200_default_config.argtypes = [C.POINTER(Config)]
    L.lbm_b200_default_config.restype = None
    L.lbm_b200_create.argtypes = [C.POINTER(Config), i64, C.POINTER(vp)]
    L.lbm_b200_mrt_moments.argtypes = [i32, i32, C.POINTER(C.c_int32)]
    L.lbm_b200_destroy.argtypes = [vp]
    L.lbm_b200_destroy.restype = None
    L.lbm_b200_set_topology.argtypes = [vp, pi64, i32]
    L.lbm_b200_set_geometry.argtypes = [vp, pdbl, pdbl, pdbl, dbl]
    L.lbm_b200_add_wall_bb.argtypes = [vp, pi64, pdbl, i64, dbl]
    L.lbm_b200_add_dirichlet_bb.argtypes = [vp, pi64, pdbl, i64, pdbl]
    L.lbm_b200_add_pressure.argtypes = [vp, pi64, pdbl, i64, dbl]
    L.lbm_b200_set_poisson.argtypes = [vp, dbl, dbl]
    L.lbm_b200_add_poisson_neem.argtypes = [vp, i32, pi64, pdbl, i64, pdbl, dbl]
    L.lbm_b200_add_wall_wetnode.argtypes = [vp, i32, pi64, pdbl, i64, i32, pdbl]
    L.lbm_b200_add_periodic.argtypes = [vp, pi64, pdbl, i64, pi64, i64, dbl]
    L.lbm_b200_set_forcing.argtypes = [vp, pi64, i64, pi64, i64, dbl]
    L.lbm_b200_set_stream.argtypes = [vp, vp]
    L.lbm_b200_init.argtypes = [vp]
    L.lbm_b200_p2p_export.argtypes = [vp, vp]
    L.lbm_b200_p2p_import.argtypes = [vp, i32, vp]
    L.lbm_b200_step.argtypes = [vp, i64]
    L.lbm_b200_step_timed.argtypes = [vp, i64, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.lbm_b200_synchronize.argtypes = [vp]
    L.lbm_b200_residual.argtypes = [vp, pdbl, C.POINTER(i32)]
    L.lbm_b200_get_populations.argtypes = [vp, vp, vp]
    L.lbm_b200_set_populations.argtypes = [vp, vp, pdbl]
    L.lbm_b200_get_vars.argtypes = [vp, vp, vp]
    L.lbm_b200_get_moments.argtypes = [vp, pdbl]
    L.lbm_b200_output_chars.argtypes = [i64]
    L.lbm_b200_output_chars.restype = i64
    L.lbm_b200_encode_output.argtypes = [vp, vp, vp, i64, C.POINTER(i64)]
    L.lbm_b200_host_alloc.argtypes = [C.POINTER(vp), i64]
    L.lbm_b200_host_free.argtypes = [vp]
    L.lbm_b200_steps_done.argtypes = [vp]
    L.lbm_b200_steps_done.restype = i64
    L.lbm_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.lbm_b200_box_ncells.argtypes = [i32, pi64]
    L.lbm_b200_box_ncells.restype = i64
    L.lbm_b200_box_topology.argtypes = [i32, pi64, np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"), pi64, i32, vp, vp]
    pi32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    L.lbm_b200_set_ghosts.argtypes = [vp, i64]
    L.lbm_b200_set_halo.argtypes = [vp, i32, pi32, pi64, pi64, pi32, pi64, pi64, pi32]
    L.lbm_b200_set_vars_halo.argtypes = [vp, pi64, pi64, pi64, pi64]
    L.lbm_b200_partition_create.argtypes = [i64, i32, i32, i32, i32, i32, ROWS_FN, vp, i32, C.POINTER(vp), C.POINTER(vp), pi64, C.POINTER(vp)]
    L.lbm_b200_partition_get.argtypes = [vp, C.POINTER(PartitionView)]
    L.lbm_b200_partition_restrict.argtypes = [vp, pi64, i64, pi64, pi64]
    L.lbm_b200_partition_restrict.restype = i64
    L.lbm_b200_partition_apply.argtypes = [vp, vp]
    L.lbm_b200_partition_destroy.argtypes = [vp]
    L.lbm_b200_partition_destroy.restype = None
    L.lbm_b200_comm_unique_id.argtypes = [C.c_char_p]
    L.lbm_b200_comm_init.argtypes = [vp, C.c_char_p, i32, i32]
    L.lbm_b200_box_rows.argtypes = [i32, pi64, pi32, pi64, i64, pi64, i32, vp]
    L.lbm_b200_sfc_index.argtypes = [i32, pdbl, i32]
    L.lbm_b200_sfc_index.restype = i64
    L.lbm_b200_debug_plan.argtypes = [vp, C.POINTER(PlanView)]
    L.lbm_b200_last_error.restype = C.c_char_p
    L.lbm_b200_abi_version.restype = C.c_int
    _LIB = L
    return L


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Solver:
    """One LBM solver instance on one CUDA device (reference: LBMSolver, src/lbm/solver.h:14)."""

    def __init__(self, ndim, ndist, nghbr, omega, *, precision=FP64, collision=BGK, arithmetic=STRICT, device=0,
                 track_vars=1, omega_minus=None, mrt_rates=None, stream=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        nghbr = _i64(nghbr)
        if nghbr.ndim != 2:
            raise ValueError("nghbr must be [ncells, stride]")
        self.ndim, self.ndist, self.nvar = int(ndim), int(ndist), int(ndim) + 1
        self.n = nghbr.shape[0]
        cfg = Config()
        self._lib.lbm_b200_default_config(C.byref(cfg))
        cfg.ndim, cfg.ndist = self.ndim, self.ndist
        cfg.precision, cfg.collision, cfg.arithmetic = precision, collision, arithmetic
        cfg.device, cfg.track_vars = device, track_vars
        cfg.omega = float(omega)
        cfg.omega_minus = float(omega if omega_minus is None else omega_minus)
        rates = np.full(27, float(omega)) if mrt_rates is None else _f64(mrt_rates)
        for i in range(min(27, len(rates))):
            cfg.mrt_rates[i] = float(rates[i])
        self._check(self._lib.lbm_b200_create(C.byref(cfg), self.n, C.byref(self._h)))
        self._check(self._lib.lbm_b200_set_topology(self._h, nghbr, nghbr.shape[1]))
        if stream is not None:
            self.set_stream(stream)

    def _check(self, rc):
        if rc != 0:
            raise LbmB200Error(rc, self._lib.lbm_b200_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lbm_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up (same names as oracle.Oracle so tests drive both with one spec)
    def set_geometry(self, center, bbmin, bbmax, cell_length):
        self._check(self._lib.lbm_b200_set_geometry(self._h, _f64(center), _f64(bbmin), _f64(bbmax), float(cell_length)))

    def add_wall_bb(self, cells, normals, tangential=0.0):
        self._check(self._lib.lbm_b200_add_wall_bb(self._h, _i64(cells), _f64(normals), len(cells), float(tangential)))

    def add_dirichlet_bb(self, cells, normals, value):
        self._check(self._lib.lbm_b200_add_dirichlet_bb(self._h, _i64(cells), _f64(normals), len(cells), _f64(value)))

    def add_wall_wetnode(self, model, cells, normals, velocity=None):
        kind = {"equilibrium": 0, "neem": 1, "nebb": 2}[model]
        v = _f64(np.zeros(self.ndim) if velocity is None else velocity)
        self._check(self._lib.lbm_b200_add_wall_wetnode(self._h, kind, _i64(cells), _f64(normals), len(cells),
                                                        int(velocity is not None), v))

    def set_poisson(self, dt, rate):
        """Poisson equation types: one variable per cell (the potential)."""
        self._check(self._lib.lbm_b200_set_poisson(self._h, float(dt), float(rate)))
        self.nvar = 1

    def add_poisson_neem(self, kind, cells, normals, values, grad=0.0):
        self._check(self._lib.lbm_b200_add_poisson_neem(self._h, int(kind == "neumann"), _i64(cells), _f64(normals), len(cells),
                                                        _f64(values), float(grad)))

    def add_pressure(self, cells, normals, pressure):
        self._check(self._lib.lbm_b200_add_pressure(self._h, _i64(cells), _f64(normals), len(cells), float(pressure)))

    def add_periodic(self, cells, normals, connected, pressure=float("nan")):
        self._check(self._lib.lbm_b200_add_periodic(self._h, _i64(cells), _f64(normals), len(cells), _i64(connected),
                                                    len(connected), float(pressure)))

    def set_forcing(self, inlet, outlet, gradient):
        self._check(self._lib.lbm_b200_set_forcing(self._h, _i64(inlet), len(inlet), _i64(outlet), len(outlet),
                                                   float(gradient)))

    # ---- multi-GPU
    def set_ghosts(self, nghost):
        self._check(self._lib.lbm_b200_set_ghosts(self._h, int(nghost)))

    def set_halo(self, peers, send_count, send_cell, send_dir, recv_count, recv_cell, recv_dir):
        i32a = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self._check(self._lib.lbm_b200_set_halo(self._h, len(peers), i32a(peers), _i64(send_count), _i64(send_cell), i32a(send_dir),
                                                _i64(recv_count), _i64(recv_cell), i32a(recv_dir)))

    def set_vars_halo(self, send_count, send_cell, recv_count, recv_cell):
        self._check(self._lib.lbm_b200_set_vars_halo(self._h, _i64(send_count), _i64(send_cell), _i64(recv_count), _i64(recv_cell)))

    def comm_init(self, unique_id, rank, nranks):
        self._check(self._lib.lbm_b200_comm_init(self._h, bytes(unique_id), int(rank), int(nranks)))

    def set_stream(self, cuda_stream):
        self._check(self._lib.lbm_b200_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def debug_plan(self):
        """The device plan as numpy arrays (host-side layout planning only; works on a handle created with device=-1)."""
        v = PlanView()
        self._check(self._lib.lbm_b200_debug_plan(self._h, C.byref(v)))
        qm = self.ndist - 1

        def arr(ptr, count, shape=None):
            a = np.ctypeslib.as_array(ptr, shape=(int(count),)).copy() if count else np.zeros(0)
            return a.reshape(shape) if shape is not None and count else a
        out = {k: int(getattr(v, k)) for k, t in PlanView._fields_ if t is C.c_int64}
        out["ref2dev"] = arr(v.ref2dev, v.n)
        out["tmpl"] = arr(v.tmpl, qm * v.chunk, (qm, int(v.chunk)))
        out["chunk_nb"] = arr(v.chunk_nb, v.n_fast_chunks * (v.nsel + 1), (int(v.n_fast_chunks), int(v.nsel) + 1))
        out["codes"] = arr(v.codes, qm * v.gen_stride, (qm, int(v.gen_stride)))
        out["copytab"] = arr(v.copytab, v.n_copy * 2, (int(v.n_copy), 2))
        out["addtab"] = arr(v.addtab, v.n_add * 4, (int(v.n_add), 4))
        nsel = int(v.nsel)
        out["wall_desc"] = arr(v.wall_desc, v.n_wall * 4, (int(v.n_wall) // (qm * nsel) if v.n_wall else 0, nsel, qm, 4))
        out["abb_p"] = arr(v.abb_p, v.n_abb)
        out["abb_cells"] = arr(v.abb_cells, v.n_abb * 3, (int(v.n_abb), 3))
        out["values"] = arr(v.values, v.n_values)
        out["stale_ref"] = arr(v.stale_ref, v.n_stale)
        out["send_index"] = arr(v.send_index, v.n_send)
        out["recv_index"] = arr(v.recv_index, v.n_recv)
        out["vsend_cells"] = arr(v.vsend_cells, v.n_vsend)
        out["chunk_abb_base"] = arr(v.chunk_abb_base, v.n_fast_chunks)
        out["chunk_abb"] = arr(v.chunk_abb, v.n_chunk_abb_rows * v.chunk, (int(v.n_chunk_abb_rows), int(v.chunk)))
        out["layout"] = arr(v.layout, self.ndist)
        return out

    # ---- run
    def init(self):
        self._check(self._lib.lbm_b200_init(self._h))

    P2P_BLOB = 1024

    def p2p_export(self):
        """this rank's mailbox description for the peer-to-peer halo (bytes; all-gather them in rank order and call p2p_import)"""
        buf = C.create_string_buffer(self.P2P_BLOB)
        self._check(self._lib.lbm_b200_p2p_export(self._h, buf))
        return buf.raw

    def p2p_import(self, blobs):
        """blobs: the exported descriptions of ALL ranks, in rank order"""
        raw = b"".join(blobs)
        assert len(raw) == self.P2P_BLOB * len(blobs)
        self._check(self._lib.lbm_b200_p2p_import(self._h, len(blobs), raw))

    def p2p_connect(self, dist, lp=None):
        """export -> torch.distributed all-gather -> import; call on every rank after init().  Returns False (and leaves the exchange on
        NCCL, on ALL ranks) when some rank cannot take part: a velocity halo of a pressure boundary across a cut (lp.vsend_count /
        lp.vrecv_count), or no halo lists at all."""
        able = True
        if lp is not None:
            able = not (sum(lp.vsend_count or [0]) or sum(lp.vrecv_count or [0])) and len(lp.peers) > 0
        flags = [None] * dist.get_world_size()
        dist.all_gather_object(flags, bool(able))
        if not all(flags):
            return False
        mine = self.p2p_export()
        blobs = [None] * dist.get_world_size()
        dist.all_gather_object(blobs, mine)
        self.p2p_import(blobs)
        return True

    def step(self, n=1):
        self._check(self._lib.lbm_b200_step(self._h, int(n)))

    def step_timed(self, n=1):
        """Returns (ms over all kernels of the n steps, ms inside the fused stream+collide kernel)."""
        a, b = C.c_float(), C.c_float()
        self._check(self._lib.lbm_b200_step_timed(self._h, int(n), C.byref(a), C.byref(b)))
        return a.value, b.value

    def synchronize(self):
        self._check(self._lib.lbm_b200_synchronize(self._h))

    def residual(self):
        out = np.zeros(self.nvar)
        bad = C.c_int32()
        self._check(self._lib.lbm_b200_residual(self._h, out, C.byref(bad)))
        return out, bool(bad.value)

    # ---- read-back (reference layout)
    def _get2(self, fn, width, want_a, want_b):
        a = np.empty((self.n, width)) if want_a else None
        b = np.empty((self.n, width)) if want_b else None
        pa = a.ctypes.data_as(C.c_void_p) if want_a else None
        pb = b.ctypes.data_as(C.c_void_p) if want_b else None
        self._check(fn(self._h, pa, pb))
        return a, b

    @property
    def f(self):
        return self._get2(self._lib.lbm_b200_get_populations, self.ndist, True, False)[0]

    @property
    def fold(self):
        return self._get2(self._lib.lbm_b200_get_populations, self.ndist, False, True)[1]

    @property
    def vars(self):
        return self._get2(self._lib.lbm_b200_get_vars, self.nvar, True, False)[0]

    @property
    def varsold(self):
        return self._get2(self._lib.lbm_b200_get_vars, self.nvar, False, True)[1]

    def moments(self):
        out = np.empty((self.n, self.nvar))
        self._check(self._lib.lbm_b200_get_moments(self._h, out))
        return out

    def encode_output(self, keep=None, out=None, raw=False):
        """Device side of LBMSolver::output: list of NVAR byte strings, the base64 payload of every field of the kept cells as the
        reference's binary VTK file stores it (lbm_b200_encode_output); keep: bool / uint8 per owned cell or None.
        out: uint8 array to receive the text (e.g. pinned memory); raw: return views of it instead of bytes objects"""
        k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        nkeep = self.n if k is None else int(k.sum())
        total = int(self._lib.lbm_b200_output_chars(nkeep)) * self.nvar
        buf = np.empty(max(1, total), dtype=np.uint8) if out is None else out
        if buf.dtype != np.uint8 or buf.size < total or not buf.flags.c_contiguous:
            raise ValueError("encode_output: `out` must be a contiguous uint8 array of at least nvar * output_chars(kept) elements")
        off = (C.c_int64 * (self.nvar + 1))()
        self._check(self._lib.lbm_b200_encode_output(self._h, None if k is None else k.ctypes.data, buf.ctypes.data, total, off))
        views = [buf[off[v]:off[v + 1]] for v in range(self.nvar)]
        return views if raw else [v.tobytes() for v in views]

    def set_populations(self, f, fold):
        """m_fold (and m_f unless None: it is not an input of the next step) in the reference's layout."""
        f = None if f is None else _f64(f)
        self._check(self._lib.lbm_b200_set_populations(self._h, None if f is None else f.ctypes.data, _f64(fold)))

    @property
    def steps_done(self):
        return int(self._lib.lbm_b200_steps_done(self._h))

    def stats(self):
        st = Stats()
        self._check(self._lib.lbm_b200_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in Stats._fields_}


class HostBuffer:
    """Page-locked host memory (lbm_b200_host_alloc) as a uint8 numpy array `.array`; `.view(dtype, shape)` reinterprets it.  The memory
    lives as long as this object: keep it while any view is in use"""

    def __init__(self, nbytes):
        self._lib = load_library()
        self._p = C.c_void_p()
        rc = self._lib.lbm_b200_host_alloc(C.byref(self._p), int(nbytes))
        if rc != 0:
            raise LbmB200Error(rc, self._lib.lbm_b200_last_error().decode())
        self.array = np.ctypeslib.as_array((C.c_uint8 * int(nbytes)).from_address(self._p.value))

    def view(self, dtype, shape):
        return self.array[:int(np.prod(shape)) * np.dtype(dtype).itemsize].view(dtype).reshape(shape)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            self._lib.lbm_b200_host_free(self._p)
            self._p = None

    def __del__(self):
        self.close()


def box_topology(shape, periodic, want_center=True, want_coords=False):
    """Synthetic benchmark box in the reference's table format (include/lbm_b200.h: lbm_b200_box_topology)."""
    L = load_library()
    shape = _i64(shape)
    ndim = len(shape)
    per = np.ascontiguousarray(periodic, dtype=np.int32)
    n = int(L.lbm_b200_box_ncells(ndim, shape))
    if n <= 0:
        raise ValueError("bad shape")
    stride = 8 if ndim == 2 else 26
    nghbr = np.empty((n, stride), dtype=np.int64)
    center = np.empty((n, ndim)) if want_center else None
    coords = np.empty((n, ndim), dtype=np.int64) if want_coords else None
    rc = L.lbm_b200_box_topology(ndim, shape, per, nghbr.reshape(-1), stride,
                                 center.ctypes.data_as(C.c_void_p) if want_center else None,
                                 coords.ctypes.data_as(C.c_void_p) if want_coords else None)
    if rc != 0:
        raise LbmB200Error(rc, L.lbm_b200_last_error().decode())
    return nghbr, center, coords


def sfc_index(x, level):
    """hilbert::index of the reference for unit-cube coordinates x (host code of the library)."""
    x = _f64(x)
    return int(load_library().lbm_b200_sfc_index(len(x), x, int(level)))


def comm_unique_id():
    """128-byte NCCL id (rank 0 creates it, the caller broadcasts it)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.lbm_b200_comm_unique_id(buf)
    if rc != 0:
        raise LbmB200Error(rc, L.lbm_b200_last_error().decode())
    return buf.raw


def box_rows(shape, periodic, cells, want_center=False):
    """Rows of the synthetic box table for the given global cell ids."""
    L = load_library()
    shape = _i64(shape)
    ndim = len(shape)
    per = np.ascontiguousarray(periodic, dtype=np.int32)
    cells = _i64(cells)
    stride = 8 if ndim == 2 else 26
    nghbr = np.empty((len(cells), stride), dtype=np.int64)
    center = np.empty((len(cells), ndim)) if want_center else None
    rc = L.lbm_b200_box_rows(ndim, shape, per, cells, len(cells), nghbr.reshape(-1), stride,
                             center.ctypes.data_as(C.c_void_p) if want_center else None)
    if rc != 0:
        raise LbmB200Error(rc, L.lbm_b200_last_error().decode())
    return nghbr, center


class NativePartition:
    """The local problem of one rank, built by the library's own partition code (lbm_b200/csrc/partition.hpp) from a row provider
    (anything with .n, .qm, .rows(ids), .sources(ids): lbm_b200.partition.TableRows / BoxRows / GridRows).  Same arrays as
    lbm_b200.partition.plan_rank (the numpy twin, tests/test_partition_native.py)."""

    def __init__(self, provider, ndim, ndist, stride, rank, world, pressure=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        qm = ndist - 1

        def rows_cb(_user, ids_p, n, rows_p, src_p):
            try:
                ids = np.ctypeslib.as_array(ids_p, shape=(n,)).copy()
                if rows_p:
                    out = np.ctypeslib.as_array(rows_p, shape=(n, stride))
                    out[:, :qm] = provider.rows(ids)[:, :qm]
                    out[:, qm:] = -1
                if src_p:
                    out = np.ctypeslib.as_array(src_p, shape=(n, stride))
                    out[:, :qm] = provider.sources(ids)[:, :qm]
                    out[:, qm:] = -1
                return 0
            except Exception:  # noqa: BLE001 -- reported through the library's error path
                return 1
        self._cb = ROWS_FN(rows_cb)
        pressure = pressure or []
        keep = [(_i64(c), _f64(nrm)) for c, nrm in pressure]
        cells = (C.c_void_p * max(1, len(keep)))(*[c.ctypes.data for c, _ in keep])
        normals = (C.c_void_p * max(1, len(keep)))(*[nrm.ctypes.data for _, nrm in keep])
        counts = _i64([len(c) for c, _ in keep] or [0])
        rc = self._lib.lbm_b200_partition_create(int(provider.n), int(ndim), int(ndist), int(stride), int(rank), int(world), self._cb, None,
                                                 len(keep), cells, normals, counts, C.byref(self._h))
        if rc != 0:
            raise LbmB200Error(rc, self._lib.lbm_b200_last_error().decode())
        v = PartitionView()
        self._lib.lbm_b200_partition_get(self._h, C.byref(v))

        def arr(ptr, count, dtype):
            return np.ctypeslib.as_array(ptr, shape=(int(count),)).astype(dtype) if count else np.zeros(0, dtype)
        self.rank, self.world = rank, world
        self.lo, self.hi, self.n_owned, self.n_ghost = int(v.lo), int(v.hi), int(v.n_owned), int(v.n_ghost)
        self.ghosts = arr(v.ghosts, v.n_ghost, np.int64)
        self.nghbr = arr(v.nghbr, (v.n_owned + v.n_ghost) * v.stride, np.int64).reshape(-1, int(v.stride))
        np_ = int(v.npeers)
        self.peers = arr(v.peers, np_, np.int32).tolist()
        self.send_count = arr(v.send_count, np_, np.int64).tolist()
        self.recv_count = arr(v.recv_count, np_, np.int64).tolist()
        self.vsend_count = arr(v.vsend_count, np_, np.int64).tolist()
        self.vrecv_count = arr(v.vrecv_count, np_, np.int64).tolist()
        self.send_cell = arr(v.send_cell, sum(self.send_count), np.int64)
        self.send_dir = arr(v.send_dir, sum(self.send_count), np.int32)
        self.recv_cell = arr(v.recv_cell, sum(self.recv_count), np.int64)
        self.recv_dir = arr(v.recv_dir, sum(self.recv_count), np.int32)
        self.vsend_cell = arr(v.vsend_cell, sum(self.vsend_count), np.int64)
        self.vrecv_cell = arr(v.vrecv_cell, sum(self.vrecv_count), np.int64)

    def restrict(self, cells, normals):
        """Boundary-condition entries of the cells this rank owns, order kept."""
        cells = _i64(cells)
        loc, idx = np.empty(len(cells), np.int64), np.empty(len(cells), np.int64)
        m = int(self._lib.lbm_b200_partition_restrict(self._h, cells, len(cells), loc, idx))
        return loc[:m], np.asarray(normals)[idx[:m]]

    def apply_halo(self, solver):
        rc = self._lib.lbm_b200_partition_apply(self._h, solver._h)
        if rc != 0:
            raise LbmB200Error(rc, self._lib.lbm_b200_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lbm_b200_partition_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
