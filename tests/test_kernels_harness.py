"""The device code of the fused path (lbm_b200/csrc/kernels.cuh, unmodified) executed on the CPU: tests/c/kernels_harness.cpp compiles it
with g++ against a stand-in <cuda_runtime.h> and drives, thread index by thread index, the gather of the fused kernel (link codes, chunk
templates, wall descriptors, pressure-face chunks, ghost blocks), the per-cell update, the halo pack / unpack kernels, and the velocity
halo of the pressure boundary condition (k_velocity_pack, k_pressure_extrapolate) -- in the order Solver::one_step launches them.
Partitioned runs emulated this way must equal the single-domain oracle BIT FOR BIT.  This is how the multi-GPU features that had no
hardware run in round 1 are checked beyond the plan level; NCCL itself is replaced by copying the wire buffers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lbm_b200
from lbm_b200 import partition
from lbm_b200.capi import pop_gather, pop_scatter

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def kh(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("kh") / "libkernels_harness.so")
    pkg = os.path.join(ROOT, "lbm_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
                           "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-unknown-pragmas", "-I", os.path.join(HERE, "c", "fake_cuda"),
                           "-DLBM_THREADS=32", "-DLBM_FAST_THREADS=32", "-pthread",  # 32 threads per block: enough for the NSEL + 1 loaders of a chunk, cheap to start
                           os.path.join(HERE, "c", "kernels_harness.cpp"), "-o", so, "-L", pkg, "-llbm_b200", f"-Wl,-rpath,{pkg}"])
    L = C.CDLL(so)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    L.kh_create.restype = vp
    L.kh_create.argtypes = [vp, C.c_int, C.c_int, C.c_double]
    L.kh_destroy.argtypes = [vp]
    L.kh_npad.restype = C.c_int64
    L.kh_npad.argtypes = [vp]
    for fn in ("kh_A", "kh_B", "kh_vars", "kh_values"):
        getattr(L, fn).restype = dp
        getattr(L, fn).argtypes = [vp]
    L.kh_uext.restype = dp
    L.kh_uext.argtypes = [vp, C.c_int]
    L.kh_set_first.argtypes = [vp, C.c_int]
    L.kh_set_vrecv.argtypes = [vp, vp, C.c_int64]
    L.kh_gather_all.argtypes = [vp, vp, vp]
    L.kh_update.argtypes = [vp]
    L.kh_step_kernel.argtypes = [vp, C.c_int]
    L.kh_step_kernel_split.argtypes = [vp, C.c_int]
    L.kh_set_collision.argtypes = [vp, C.c_int, C.c_double, vp]
    L.kh_velocity_pack.argtypes = [vp, vp, C.c_int, vp]
    L.kh_pressure_extrapolate.argtypes = [vp]
    L.kh_halo_pack.argtypes = [vp, vp, C.c_int64, vp]
    L.kh_halo_unpack.argtypes = [vp, vp, C.c_int64, vp]
    L.kh_swap.argtypes = [vp, C.c_int]
    L.kh_check_div_const.restype = C.c_int64
    L.kh_check_div_const.argtypes = [vp, C.c_int64]
    return L


class Rank:
    """one rank of the emulated run: inspection-only solver + harness context + its view of the device arrays"""

    def __init__(self, L, spec, lp, init_f, init_fold):
        from cases3d import add_restricted
        self.L, self.lp, self.q, self.ndim = L, lp, spec.ndist, spec.ndim
        self.solver = add_restricted(lbm_b200.Solver(spec.ndim, spec.ndist, lp.nghbr, spec.omega, device=-1), spec, lp)
        lp.apply_halo(self.solver)
        self.plan = self.solver.debug_plan()
        self.h = L.kh_create(self.solver._h, spec.ndim, spec.ndist, spec.omega)
        assert self.h
        self.npad = L.kh_npad(self.h)
        self.glob = np.concatenate([np.arange(lp.lo, lp.hi), lp.ghosts])     # local cell -> global id
        self.r2d = self.plan["ref2dev"].astype(np.int64)
        A = self.arr("kh_A", self.q)
        pop_scatter(self.plan, A, self.r2d[:lp.n_owned], init_f[lp.lo:lp.hi])   # lbm_b200_init: f = feq of the initial condition
        vals = np.ctypeslib.as_array(L.kh_values(self.h), shape=(max(1, self.plan["n_values"]),))
        dev2loc = np.full(self.npad, -1)
        dev2loc[self.r2d] = np.arange(len(self.r2d))
        for k, sr in enumerate(self.plan["stale_ref"].astype(np.int64)):      # slots nothing writes keep their initial value
            vals[k + 1] = init_fold[self.glob[dev2loc[sr // self.q]], sr % self.q]

    def arr(self, fn, width):
        return np.ctypeslib.as_array(getattr(self.L, fn)(self.h), shape=(width, self.npad))

    def owned(self, fn, width):
        """populations of the owned cells, [n_owned, Q], read through the per-direction in-chunk layouts"""
        return pop_gather(self.plan, self.arr(fn, width), self.r2d[:self.lp.n_owned])

    def close(self):
        self.L.kh_destroy(self.h)
        self.solver.close()


def emulate(L, spec, world, steps, oracle_mod, kernel=False, collision=0):
    from cases3d import mrt_rates, pressure_surfaces
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    rates = np.ascontiguousarray(mrt_rates(spec.ndist, spec.omega))
    om_minus = 1.0 / 0.8
    o.set_collision(collision, om_minus, rates)
    o.init()
    init_f, init_fold = o.f.copy(), o.fold.copy()
    provider = partition.TableRows(spec.nghbr, spec.ndist)
    ranks = [Rank(L, spec, partition.plan_rank(provider, r, world, spec.nghbr.shape[1], pressure_surfaces(spec)), init_f, init_fold)
             for r in range(world)]
    for rk in ranks:
        L.kh_set_collision(rk.h, collision, om_minus, rates.ctypes.data)
    stats = dict(fast=sum(rk.plan["n_fast_chunks"] for rk in ranks), abb_chunks=sum(rk.plan["n_chunk_abb_rows"] for rk in ranks),
                 ghost_blocks=sum(rk.plan["n_ghost_blocks"] for rk in ranks), vrecv=sum(rk.plan["n_vrecv"] for rk in ranks))
    try:
        for step in range(1, steps + 1):
            o.step(1)
            for rk in ranks:                                    # main kernel: every owned cell, A -> B
                if kernel == "split":
                    L.kh_step_kernel_split(rk.h, 3)             # outer launch, then inner launch (overlapped path of one_step)
                elif kernel:
                    L.kh_step_kernel(rk.h, 3)                   # the kernels themselves: k_step_generic + 3 persistent k_step_fast CTAs
                else:
                    L.kh_update(rk.h)
            wires = {}
            for r, rk in enumerate(ranks):                      # k_halo_pack + k_velocity_pack, per peer in list order
                so = vo = 0
                for k, peer in enumerate(rk.lp.peers):
                    ns, nv = rk.lp.send_count[k], rk.lp.vsend_count[k] if rk.lp.vsend_count else 0
                    idx = np.ascontiguousarray(rk.plan["send_index"][so:so + ns].astype(np.int64))
                    buf = np.empty(ns)
                    L.kh_halo_pack(rk.h, idx.ctypes.data, ns, buf.ctypes.data)
                    cells = np.ascontiguousarray(rk.plan["vsend_cells"][vo:vo + nv].astype(np.int32))
                    vbuf = np.empty(3 * nv)
                    if nv:
                        L.kh_velocity_pack(rk.h, cells.ctypes.data, nv, vbuf.ctypes.data)
                    wires[(r, peer)] = (buf, vbuf)
                    so += ns
                    vo += nv
            for r, rk in enumerate(ranks):                      # "NCCL": the peer's buffers arrive; k_halo_unpack; velocity receive buffer
                ro = 0
                vparts = []
                for k, peer in enumerate(rk.lp.peers):
                    buf, vbuf = wires[(peer, r)]
                    nr = rk.lp.recv_count[k]
                    assert len(buf) == nr and len(vbuf) == 3 * (rk.lp.vrecv_count[k] if rk.lp.vrecv_count else 0)
                    idx = np.ascontiguousarray(rk.plan["recv_index"][ro:ro + nr].astype(np.int64))
                    L.kh_halo_unpack(rk.h, idx.ctypes.data, nr, np.ascontiguousarray(buf).ctypes.data)
                    vparts.append(vbuf)
                    ro += nr
                v = np.ascontiguousarray(np.concatenate(vparts)) if vparts else np.zeros(0)
                L.kh_set_vrecv(rk.h, v.ctypes.data, len(v))
            for rk in ranks:                                    # pressure extrapolation of this step, then the buffers flip
                nabb = rk.plan["n_abb"]
                if nabb:
                    L.kh_pressure_extrapolate(rk.h)
                L.kh_swap(rk.h, int(nabb > 0))
            for r, rk in enumerate(ranks):
                mine = rk.owned("kh_A", rk.q)
                assert np.array_equal(mine, o.f[rk.lp.lo:rk.lp.hi]), f"step {step}, rank {r}: m_f differs"
        for r, rk in enumerate(ranks):                          # k_gather_all: m_fold and its moments of the final state
            fold = np.zeros((rk.q, rk.npad))
            mom = np.zeros((rk.ndim + 1, rk.npad))
            L.kh_gather_all(rk.h, fold.ctypes.data, mom.ctypes.data)
            sel = rk.r2d[:rk.lp.n_owned]
            assert np.array_equal(fold[:, sel].T, o.fold[rk.lp.lo:rk.lp.hi]), f"rank {r}: m_fold differs"
        o.update_moments()
        for r, rk in enumerate(ranks):
            fold = np.zeros((rk.q, rk.npad))
            mom = np.zeros((rk.ndim + 1, rk.npad))
            L.kh_gather_all(rk.h, fold.ctypes.data, mom.ctypes.data)
            assert np.array_equal(mom[:, rk.r2d[:rk.lp.n_owned]].T, o.vars[rk.lp.lo:rk.lp.hi]), f"rank {r}: moments differ"
    finally:
        for rk in ranks:
            rk.close()
    return stats


def pressure_box(shape, ndist):
    from test_plan_cpu import _pressure_box
    return _pressure_box(shape, ndist)


@pytest.mark.parametrize("world,shape,ndist", [(1, (24, 24, 24), 19), (2, (18, 16, 16), 19), (3, (10, 10, 10), 27), (2, (34, 64), 9), (2, (26, 24, 24), 19)])
def test_partitioned_pressure_boxes_on_the_device_code(world, shape, ndist, kh, oracle_mod):
    stats = emulate(kh, pressure_box(shape, ndist), world, 6, oracle_mod)
    if shape in ((18, 16, 16), (10, 10, 10), (34, 64)):
        assert stats["vrecv"] > 0, "no pressure stencil crosses a cut: the velocity halo was not exercised"
    if shape in ((24, 24, 24), (26, 24, 24)):
        assert stats["abb_chunks"] > 0, "no pressure-face chunk on the index-free path"


@pytest.mark.parametrize("name,world", [("sphere3d", 1), ("sphere3d", 4), ("step3d", 3)])
def test_baseline_configs_on_the_device_code(name, world, kh, oracle_mod):
    """BASELINE.json configs[3] / [4] with their collision operators: sphere = MRT, step = TRT (oracle/lbm_oracle.c: literature forms)"""
    from cases3d import build_case
    stats = emulate(kh, build_case(name, 5), world, 4, oracle_mod, collision={"sphere3d": 2, "step3d": 1}[name])
    if world > 1:
        assert stats["ghost_blocks"] >= 0


@pytest.mark.parametrize("world,shape,ndist", [(1, (24, 24, 24), 19), (2, (26, 24, 24), 19), (1, (16, 16, 16), 27), (2, (96, 64), 9)])
def test_the_fused_kernel_itself_on_the_cpu(world, shape, ndist, kh, oracle_mod):
    """k_step_generic and k_step_fast -- persistent chunk CTAs with the ticket counter, the cp.async tile pipeline through swizzled shared memory and its
    barrier -- run as 32 OS threads per block: interior chunks, wall / edge / corner chunks, pressure-face chunks, ghost blocks."""
    stats = emulate(kh, pressure_box(shape, ndist), world, 3, oracle_mod, kernel=True)
    assert stats["fast"] > 0


@pytest.mark.parametrize("name,collision", [("step3d", 1), ("sphere3d", 2)])
def test_the_fused_kernel_with_trt_and_mrt(name, collision, kh, oracle_mod):
    from cases3d import build_case
    emulate(kh, build_case(name, 5), 2, 2, oracle_mod, kernel=True, collision=collision)


@pytest.mark.parametrize("world,shape,ndist,periodic", [(2, (32, 16, 16), 19, True), (4, (32, 32, 16), 19, True), (2, (16, 16, 24), 27, True)])
def test_outer_and_inner_launch_of_the_overlapped_path(world, shape, ndist, periodic, kh, oracle_mod):
    """The benchmark box cut into SFC ranges, stepped the way the overlapped path does it: outer cells first (their own ticket
    counter), then the inner cells; wall / edge chunks next to ghost blocks.  Bit-identical to the single-domain oracle."""
    from test_plan_cpu import box_spec
    spec, _ = box_spec(shape, ndist, (True, False, False), ("+z", (0.05, 0.0, 0.0)))
    stats = emulate(kh, spec, world, 3, oracle_mod, kernel="split")
    assert stats["fast"] > 0 and stats["ghost_blocks"] > 0


def test_division_by_the_lattice_constants_is_correctly_rounded(kh):
    """STRICT arithmetic divides by cs^2, 2 cs^4 and 2 cs^2 like the reference (equilibrium_func.h:53), but through two fused multiply-adds
    (Ar::div_const).  The result must equal true IEEE division bit for bit: 3 x 3 M operands over 30 decades, both signs, plus edge values."""
    rng = np.random.default_rng(11)
    a = np.concatenate([rng.standard_normal(1_500_000) * 10.0 ** rng.uniform(-25, 5, 1_500_000), rng.random(1_500_000) * 0.2,
                        np.array([0.0, -0.0, 1.0, 3.0, 1.0 / 3.0, 2.0 / 9.0, 2.0 / 3.0, 1e-200, -1e-200, 1e200])])
    a = np.ascontiguousarray(a)
    assert kh.kh_check_div_const(a.ctypes.data, len(a)) == 0
