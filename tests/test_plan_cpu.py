"""The device plan checked on the CPU: a numpy interpreter (tests/plan_interpreter.py) executes the fused kernel's gather from the
arrays lbm_b200_debug_plan exports and must reproduce the oracle's m_fold (= what preApply -> push -> apply leave behind) exactly.
Covers the inverted push table, bounce-back / moving-wall / periodic-copy resolution, stale slots, chunk templates in device order,
wall descriptors, and -- for partitioned plans -- ghost blocks and the halo index lists."""
import numpy as np

from lbm_b200.capi import pop_scatter
import pytest

import lbm_b200
from casebuilder import CaseSpec, load_golden
from gridgen import box_grid
from lbm_b200 import partition
from plan_interpreter import gather, to_device


def plan_only_solver(spec, **kw):
    s = lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega, device=-1, **kw)
    return spec.apply_to(s)


def box_spec(shape, ndist, periodic, lid):
    g = box_grid(shape, periodic)
    ndim = len(shape)
    spec = CaseSpec(name=f"box{shape}", ndim=ndim, ndist=ndist, nghbr=g["nghbr"], omega=1.0 / 0.6, center=g["center"],
                    bbmin=g["bbmin"], bbmax=g["bbmax"], cell_length=g["cell_length"])
    for nm in sorted(["-x", "+x", "-y", "+y", "-z", "+z"][:2 * ndim]):
        cells, normals = g["surfaces"][nm]
        if len(cells) == 0:
            continue
        if nm == lid[0]:
            spec.bcs.append(dict(kind="dirichlet_bb", cells=cells, normals=normals, value=np.array(lid[1], float)))
        else:
            spec.bcs.append(dict(kind="wall_bb", cells=cells, normals=normals, tangential=0.0))
    return spec, g


def test_inspection_handle_refuses_to_compute():
    spec = load_golden("couette")
    s = plan_only_solver(spec)
    with pytest.raises(lbm_b200.LbmB200Error) as e:
        s.init()
    assert e.value.code == -3 and "no CPU compute path" in str(e.value)


@pytest.mark.parametrize("name", ["couette_ml_p3u5", "couette_ml_u5m6", "couette_ml_p4u5m7", "sphere_ml_p4u6", "step_ml_p3u5"])
def test_plan_gather_on_multi_level_grids(name, oracle_mod):
    """every level is its own lattice; several cells may push into one slot (highest source wins)."""
    from plan_interpreter import extrapolated_velocity, stale_values
    spec = load_golden(name)
    plan = plan_only_solver(spec).debug_plan()
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    values = stale_values(plan, o.fold.copy(), spec.ndist)
    dev2ref = np.full(plan["npad"], -1)
    dev2ref[plan["ref2dev"]] = np.arange(plan["n"])
    for _ in range(3):
        o.step(1)
        uext = extrapolated_velocity(plan, lambda n: o.vars[dev2ref[n], :spec.ndim], None, spec.ndim)
        mine = gather(plan, to_device(plan, o.f, spec.ndist), spec.ndist, values=values, uext=uext)
        assert np.array_equal(mine, o.fold)


@pytest.mark.parametrize("name", ["couette", "couette_bnd", "couette_bnd_bbDirichlet"])
def test_plan_gather_equals_oracle_fold_on_reference_cases(name, oracle_mod):
    spec = load_golden(name)
    plan = plan_only_solver(spec).debug_plan()
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    stale = plan["stale_ref"]
    values = plan["values"].copy()
    init_fold = o.fold.copy()
    dev2ref = np.full(plan["npad"], -1)
    dev2ref[plan["ref2dev"]] = np.arange(plan["n"])
    for k, sr in enumerate(stale):  # slots nothing writes keep their initial value (filled by lbm_b200_init on the device)
        values[k + 1] = init_fold[dev2ref[sr // spec.ndist], sr % spec.ndist]
    for _ in range(3):
        o.step(1)
        mine = gather(plan, to_device(plan, o.f, spec.ndist), spec.ndist, values=values)
        assert np.array_equal(mine, o.fold)


@pytest.mark.parametrize("shape,ndist,periodic,lid", [((16, 32, 32), 19, (True, False, False), ("+z", (0.05, 0.0, 0.0))),
                                                      ((16, 8, 8), 27, (True, True, False), ("+z", (0.05, 0.01, 0.0))),
                                                      ((64, 64), 9, (True, False), ("+y", (0.05, 0.0)))])
def test_plan_gather_on_boxes_with_wall_chunks(shape, ndist, periodic, lid, oracle_mod):
    spec, _ = box_spec(shape, ndist, periodic, lid)
    plan = plan_only_solver(spec).debug_plan()
    assert plan["n_fast_chunks"] > 0, "single-wall chunks should be on the index-free path"
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    o.step(4)
    mine = gather(plan, to_device(plan, o.f, ndist), ndist)
    assert np.array_equal(mine, o.fold)


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_plan_ghost_blocks_and_halo_lists(world, oracle_mod):
    """Every rank's plan, fed with the global post-collision state through its halo lists, must reproduce the owned part of the
    single-domain m_fold; the cut chunks must be on the index-free path (ghost blocks)."""
    shape, ndist, periodic = (32, 16, 16), 19, (True, False, False)
    spec, g = box_spec(shape, ndist, periodic, ("+z", (0.05, 0.0, 0.0)))
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, ndist, spec.nghbr, spec.omega))
    o.init()
    o.step(5)
    plans, lps = [], []
    for r in range(world):
        lp = partition.plan_rank(partition.TableRows(spec.nghbr, ndist), r, world, spec.nghbr.shape[1])
        s = lbm_b200.Solver(spec.ndim, ndist, lp.nghbr, spec.omega, device=-1)
        for bc in spec.bcs:
            cells, normals = lp.restrict(bc["cells"], bc["normals"])
            if len(cells) == 0:
                continue
            if bc["kind"] == "dirichlet_bb":
                s.add_dirichlet_bb(cells, normals, bc["value"])
            else:
                s.add_wall_bb(cells, normals, 0.0)
        lp.apply_halo(s)
        plans.append(s.debug_plan())
        lps.append(lp)
    for r, (plan, lp) in enumerate(zip(plans, lps)):
        assert plan["n_ghost_blocks"] > 0
        # device state of this rank: owned cells from the global state; ghosts only through the halo lists of the peers
        A = np.zeros((ndist, plan["npad"]))
        own = np.arange(lp.lo, lp.hi)
        pop_scatter(plan, A, plan["ref2dev"][:lp.n_owned].astype(np.int64), o.f[own])
        ro = 0
        for k, q in enumerate(lp.peers):
            pq, lq = plans[q], lps[q]
            kq = lq.peers.index(r)
            so = sum(lq.send_count[:kq])
            ns = lq.send_count[kq]
            assert ns == lp.recv_count[k]
            # what rank q packs for me: flat indices into ITS device arrays
            Aq = np.zeros((ndist, pq["npad"]))
            pop_scatter(pq, Aq, pq["ref2dev"][:lq.n_owned].astype(np.int64), o.f[np.arange(lq.lo, lq.hi)])
            wire = Aq.reshape(-1)[pq["send_index"][so:so + ns]]
            A.reshape(-1)[plan["recv_index"][ro:ro + ns]] = wire
            ro += ns
        mine = gather(plan, A, ndist)
        assert np.array_equal(mine, o.fold[own]), f"rank {r}"
    # the chunks at the cut stay on the index-free path wherever a chunk touches at most one wall
    assert sum(p["n_fast_chunks"] for p in plans) > 0 or world > 2


def _pressure_box(shape, ndist):
    from dist_worker import bcs_for
    ndim = len(shape)
    g = box_grid(shape, (False,) * ndim)
    spec = CaseSpec(name=f"pbox{shape}", ndim=ndim, ndist=ndist, nghbr=g["nghbr"], omega=1.0 / 0.6, center=g["center"],
                    bbmin=g["bbmin"], bbmax=g["bbmax"], cell_length=g["cell_length"])
    for kind, cells, normals, val in bcs_for(g, ndim, pressure=True):
        if kind == "pressure":
            spec.bcs.append(dict(kind="pressure", cells=cells, normals=normals, pressure=val))
        elif kind == "dirichlet":
            spec.bcs.append(dict(kind="dirichlet_bb", cells=cells, normals=normals, value=val))
        else:
            spec.bcs.append(dict(kind="wall_bb", cells=cells, normals=normals, tangential=0.0))
    return spec


def _add_restricted(s, spec, lp):
    for bc in spec.bcs:
        cells, normals = lp.restrict(bc["cells"], bc["normals"])
        if len(cells) == 0:
            continue
        if bc["kind"] == "dirichlet_bb":
            s.add_dirichlet_bb(cells, normals, bc["value"])
        elif bc["kind"] == "pressure":
            s.add_pressure(cells, normals, bc["pressure"])
        else:
            s.add_wall_bb(cells, normals, 0.0)


@pytest.mark.parametrize("shape,ndist", [((18, 16, 16), 19), ((34, 64), 9), ((24, 24, 24), 19), ((24, 24, 24), 27), ((96, 96), 9)])
def test_plan_pressure_boundary_single_domain(shape, ndist, oracle_mod):
    """anti-bounce-back entries: the plan's (cell, n1, n2, p) table + the interpreter's extrapolation reproduce the oracle.  With three
    chunks per side the chunk in the middle of a pressure face touches nothing else and must stay on the index-free path: its wall
    descriptor marks the slots as anti-bounce-back and chunk_abb names each cell's pressure entry."""
    spec = _pressure_box(shape, ndist)
    plan = plan_only_solver(spec).debug_plan()
    assert plan["n_abb"] > 0
    if min(shape) >= 3 * (8 if len(shape) == 3 else 32):
        assert plan["n_chunk_abb_rows"] >= 2, "the face-centre chunks of the pressure in-/outlet should be fast chunks (edge chunks too)"
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, ndist, spec.nghbr, spec.omega))
    o.init()
    dev2ref = np.full(plan["npad"], -1)
    dev2ref[plan["ref2dev"]] = np.arange(plan["n"])
    for _ in range(3):
        o.step(1)
        ab = plan["abb_cells"]
        uext = 1.5 * o.vars[dev2ref[ab[:, 1]], :spec.ndim] - 0.5 * o.vars[dev2ref[ab[:, 2]], :spec.ndim]
        mine = gather(plan, to_device(plan, o.f, ndist), ndist, uext=uext)
        assert np.array_equal(mine, o.fold)


@pytest.mark.parametrize("world,shape,ndist", [(2, (18, 16, 16), 19), (3, (10, 10, 10), 27), (2, (34, 64), 9), (2, (26, 24, 24), 19)])
def test_partitioned_plan_pressure_velocity_halo(world, shape, ndist, oracle_mod):
    """A partition cut between a pressure cell and its inward neighbours: the plan refers to slots of the
    received velocity halo, the peers' send lists fill exactly those slots, and the gather reproduces the single-domain m_fold."""
    spec = _pressure_box(shape, ndist)
    ndim = spec.ndim
    pressure = [(bc["cells"], bc["normals"]) for bc in spec.bcs if bc["kind"] == "pressure"]
    o = spec.apply_to(oracle_mod.Oracle(ndim, ndist, spec.nghbr, spec.omega))
    o.init()
    plans, lps = [], []
    for r in range(world):
        lp = partition.plan_rank(partition.TableRows(spec.nghbr, ndist), r, world, spec.nghbr.shape[1], pressure)
        s = lbm_b200.Solver(ndim, ndist, lp.nghbr, spec.omega, device=-1)
        _add_restricted(s, spec, lp)
        lp.apply_halo(s)
        plans.append(s.debug_plan())
        lps.append(lp)
    assert sum(p["n_vrecv"] for p in plans) == sum(len(p["vsend_cells"]) for p in plans)
    if shape == (26, 24, 24):  # chunk-aligned cut: no velocity crossing, but pressure-face chunks next to ghost blocks
        assert sum(p["n_chunk_abb_rows"] for p in plans) > 0
    else:
        assert sum(p["n_vrecv"] for p in plans) > 0
    for _ in range(3):
        o.step(1)
        for r, (plan, lp) in enumerate(zip(plans, lps)):
            own = np.arange(lp.lo, lp.hi)
            dev2glob = np.full(plan["npad"], -1)
            dev2glob[plan["ref2dev"][:lp.n_owned]] = own
            A = np.zeros((ndist, plan["npad"]))
            pop_scatter(plan, A, plan["ref2dev"][:lp.n_owned].astype(np.int64), o.f[own])
            vrecv = np.zeros((plan["n_vrecv"], ndim))
            ro = vo = 0
            for k, q in enumerate(lp.peers):
                pq, lq = plans[q], lps[q]
                kq = lq.peers.index(r)
                q2glob = np.full(pq["npad"], -1)
                q2glob[pq["ref2dev"][:lq.n_owned]] = np.arange(lq.lo, lq.hi)
                Aq = np.zeros((ndist, pq["npad"]))
                pop_scatter(pq, Aq, pq["ref2dev"][:lq.n_owned].astype(np.int64), o.f[np.arange(lq.lo, lq.hi)])
                so, ns = sum(lq.send_count[:kq]), lq.send_count[kq]
                assert ns == lp.recv_count[k]
                A.reshape(-1)[plan["recv_index"][ro:ro + ns]] = Aq.reshape(-1)[pq["send_index"][so:so + ns]]
                ro += ns
                vso, vns = sum(lq.vsend_count[:kq]), lq.vsend_count[kq]
                assert vns == lp.vrecv_count[k]
                vrecv[vo:vo + vns] = o.vars[q2glob[pq["vsend_cells"][vso:vso + vns].astype(np.int64)], :ndim]  # what rank q's k_velocity_pack sends
                vo += vns

            def vel(n):
                out = np.empty((len(n), ndim))
                loc = n >= 0
                out[loc] = o.vars[dev2glob[n[loc]], :ndim]
                out[~loc] = vrecv[-n[~loc] - 1]
                return out
            ab = plan["abb_cells"]
            uext = 1.5 * vel(ab[:, 1]) - 0.5 * vel(ab[:, 2]) if len(ab) else None
            mine = gather(plan, A, ndist, uext=uext)
            assert np.array_equal(mine, o.fold[own]), f"rank {r}"
