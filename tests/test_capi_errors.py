"""Error behaviour and edge cases of the C ABI (include/lbm_b200.h) that need no GPU: argument checks, call order, and the
configurations the device plan refuses because the reference's result depends on the order of its serial boundary loops.  Uses
inspection-only handles (config.device = -1): set-up calls and the host-side planner work, computing is refused.  Messages reuse the
reference's wording where one exists (src/lbm/solverExe.h:90, src/lbm/constants.h:75, src/lbm/bnd/bnd_periodic.h:180)."""
import ctypes as C

import numpy as np
import pytest

import lbm_b200
from gridgen import box_grid
from lbm_b200.capi import Config, load_library


def make(ndim=2, ndist=9, shape=(8, 8), periodic=(False, False), **kw):
    g = box_grid(shape, periodic)
    return lbm_b200.Solver(ndim, ndist, g["nghbr"], kw.pop("omega", 1.2), device=-1, **kw), g


def code_of(fn, *a, **k):
    with pytest.raises(lbm_b200.LbmB200Error) as e:
        fn(*a, **k)
    return e.value.code, str(e.value)


def test_create_rejects_bad_configurations():
    lib = load_library()
    nghbr = np.full((4, 8), -1, dtype=np.int64)
    for kw, text in ((dict(ndim=2, ndist=7), "Unsupported model"), (dict(ndim=4, ndist=9), "Unsupported model")):
        c, msg = code_of(lbm_b200.Solver, kw["ndim"], kw["ndist"], nghbr, 1.0, device=-1)
        assert c == -1 and text in msg
    for omega in (0.0, 2.0, -1.0, float("nan")):
        c, msg = code_of(lbm_b200.Solver, 2, 9, nghbr, omega, device=-1)
        assert c == -1 and "omega" in msg
    c, msg = code_of(lbm_b200.Solver, 2, 9, nghbr, 1.0, device=-1, collision=7)
    assert c == -1 and "Invalid equation configuration!" in msg
    c, msg = code_of(lbm_b200.Solver, 2, 9, nghbr, 1.0, device=-1, precision=5)
    assert c == -1
    cfg = Config()
    lib.lbm_b200_default_config(C.byref(cfg))
    cfg.abi_version = 99
    h = C.c_void_p()
    assert lib.lbm_b200_create(C.byref(cfg), 4, C.byref(h)) == -1 and b"ABI version" in lib.lbm_b200_last_error()
    lib.lbm_b200_default_config(C.byref(cfg))
    cfg.device = -1
    assert lib.lbm_b200_create(C.byref(cfg), 0, C.byref(h)) == -1  # empty grid
    assert lib.lbm_b200_create(None, 4, C.byref(h)) == -1


def test_topology_and_cell_lists_are_range_checked():
    g = box_grid((4, 4), (False, False))
    bad = g["nghbr"].copy()
    bad[3, 1] = 16  # one past the end
    c, msg = code_of(lbm_b200.Solver, 2, 9, bad, 1.0, device=-1)
    assert c == -1 and "out of range" in msg
    bad[3, 1] = -2
    c, msg = code_of(lbm_b200.Solver, 2, 9, bad, 1.0, device=-1)
    assert c == -1
    narrow = np.full((16, 7), -1, dtype=np.int64)  # stride smaller than Q-1
    c, msg = code_of(lbm_b200.Solver, 2, 9, narrow, 1.0, device=-1)
    assert c == -1 and "neighbour table" in msg
    s, g = make()
    n = np.tile([0.0, -1.0], (1, 1))
    c, msg = code_of(s.add_wall_bb, [64], n)
    assert c == -1 and "out of range" in msg
    c, msg = code_of(s.add_pressure, [-1], n, 1.0)
    assert c == -1
    c, msg = code_of(s.add_periodic, [0], n, [], float("nan"))
    assert c == -1 and "Invalid connected surface" in msg
    c, msg = code_of(s.set_ghosts, 64)
    assert c == -1
    c, msg = code_of(s.set_vars_halo, [0], [], [0], [])
    assert c == -2 and "set_halo first" in msg


def test_empty_boundary_lists_are_accepted_and_change_nothing():
    s, g = make(shape=(16, 16))
    s.add_wall_bb(np.zeros(0, np.int64), np.zeros((0, 2)))
    s.add_pressure(np.zeros(0, np.int64), np.zeros((0, 2)), 1.0)
    plan = s.debug_plan()
    assert plan["n_abb"] == 0 and plan["n_add"] == 0
    # a box without any boundary condition: every slot at the rim is one nothing ever writes
    missing = int((g["nghbr"][:, :8] < 0).sum())
    assert plan["n_stale"] == missing


def test_inspection_handle_has_no_compute_path():
    s, _ = make()
    for fn in (s.init, lambda: s.step(1), lambda: s.residual(), lambda: s.f, lambda: s.moments(), lambda: s.stats()):
        c, msg = code_of(fn)
        assert c in (-2, -3)
    assert s.steps_done == 0


def test_pressure_cells_need_two_inward_neighbours():
    # a channel only two cells long in x: the second inward neighbour does not exist
    g = box_grid((2, 8), (False, False))
    s = lbm_b200.Solver(2, 9, g["nghbr"], 1.2, device=-1)
    cells, normals = g["surfaces"]["-x"]
    s.add_pressure(cells, normals, 1.0)
    c, msg = code_of(s.debug_plan)
    assert c == -1 and "two inward neighbours" in msg


def test_order_dependent_pressure_stencil_is_refused():
    """LBMBnd_Pressure reads m_vars of n1 / n2 as earlier entries of the same pass left them (bnd_pressure.h:68-93): when n1 is itself
    an earlier pressure cell the reference's result depends on the loop order; the plan refuses instead of silently reordering."""
    g = box_grid((8, 8), (False, False))
    s = lbm_b200.Solver(2, 9, g["nghbr"], 1.2, device=-1)
    cells, normals = g["surfaces"]["-x"]
    inner = g["nghbr"][cells, 1]  # the +x neighbours: a second "pressure surface" one cell further in, processed first
    s.add_pressure(inner, normals, 1.0)
    s.add_pressure(cells, normals, 1.0)
    c, msg = code_of(s.debug_plan)
    assert c == -1 and "order-dependent" in msg


def test_tangential_wall_velocity_is_2d_only():
    s, g = make(ndim=3, ndist=19, shape=(8, 8, 8), periodic=(True, True, False))
    cells, normals = g["surfaces"]["+z"]
    s.add_wall_bb(cells, normals, 0.1)
    c, msg = code_of(s.debug_plan)
    assert c == -1 and "2D only" in msg  # reference: bnd_wall.h:52-54


def test_halo_lists_are_checked():
    s, g = make(shape=(8, 8))
    s.set_ghosts(8)
    # a receive entry that is not a ghost cell
    s.set_halo([1], [0], np.zeros(0, np.int64), np.zeros(0, np.int32), [1], [3], [1])
    c, msg = code_of(s.debug_plan)
    assert c == -1 and "not a ghost" in msg


def test_wetnode_walls_have_no_fused_plan_and_are_not_partitioned():
    s, g = make(shape=(8, 8), periodic=(True, False))
    cells, normals = g["surfaces"]["+y"]
    s.add_wall_wetnode("equilibrium", cells, normals, np.array([0.1, 0.0]))
    c, msg = code_of(s.debug_plan)
    assert c == -5 and "wet-node" in msg


def test_sfc_index_argument_checks():
    from lbm_b200.capi import sfc_index
    assert sfc_index([0.1, 0.9], 0) == 0
    lib = load_library()
    x = np.array([0.5, 0.5, 0.5])
    assert lib.lbm_b200_sfc_index(3, x, 21) == -1   # 3 * 21 bits do not fit the key
    assert lib.lbm_b200_sfc_index(5, x, 2) == -1
    shape = np.array([4, 0], dtype=np.int64)
    assert lib.lbm_b200_box_ncells(2, shape) == -1


def test_poisson_entry_points_check_their_arguments():
    """Poisson equation types (lbm_b200_set_poisson / lbm_b200_add_poisson_neem): lattice and argument checks, no fused plan"""
    g1 = np.array([[-1, 1], [0, 2], [1, 3], [2, -1]], dtype=np.int64)  # a 1D line of four cells
    s = lbm_b200.Solver(1, 3, g1, 1.0, device=-1)
    s.set_poisson(0.25, 27.79)
    assert s.nvar == 1
    s.add_poisson_neem("dirichlet", [0], np.array([[-1.0]]), [1.0])
    s.add_poisson_neem("neumann", [3], np.array([[1.0]]), [0.0])
    c, msg = code_of(s.debug_plan)
    assert c == -5 and "Poisson" in msg
    c, msg = code_of(s.set_poisson, 0.0, 1.0)
    assert c == -1
    s3, _ = make(ndim=3, ndist=19, shape=(4, 4, 4), periodic=(True, True, True))
    c, msg = code_of(s3.set_poisson, 0.1, 1.0)
    assert c == -1 and "Unsupported model" in msg   # m_canPoisson, constants.h
    c, msg = code_of(lbm_b200.Solver, 1, 5, g1, 1.0, device=-1)
    assert c == -1 and "Unsupported model" in msg


def test_host_alloc_argument_checks_and_no_device():
    """lbm_b200_host_alloc / lbm_b200_host_free: argument errors are return codes; without a CUDA device the allocation fails loudly
    (no silent pageable substitute), freeing NULL is accepted"""
    import torch
    lib = load_library()
    p = C.c_void_p()
    assert lib.lbm_b200_host_alloc(None, 64) == -1
    assert lib.lbm_b200_host_alloc(C.byref(p), 0) == -1 and lib.lbm_b200_host_alloc(C.byref(p), -8) == -1
    assert lib.lbm_b200_host_free(None) == 0
    if not torch.cuda.is_available():
        c, msg = code_of(lbm_b200.HostBuffer, 1 << 20)
        assert c != 0 and "cudaMallocHost" in msg
