"""BASELINE.json configs[3] (flow past a sphere, D3Q27) and configs[4] (channel with a step, D3Q19, pressure outflow) as 3D cases
(tests/cases3d.py), on the CPU: the device plan of the single-domain run and of SFC-partitioned runs (2 and 8 ranks), executed by the
numpy plan interpreter, must reproduce the oracle's m_fold bit for bit.  The GPU versions are in tests/test_baseline_configs_gpu.py."""
import numpy as np
import pytest

import lbm_b200
from cases3d import add_restricted, build_case, pressure_surfaces
from lbm_b200 import partition
from plan_interpreter import extrapolated_velocity, gather, partitioned_fold, stale_values, to_device

_CACHE = {}


def case(name, level):
    if (name, level) not in _CACHE:
        _CACHE[(name, level)] = build_case(name, level)
    return _CACHE[(name, level)]


@pytest.mark.parametrize("name", ["sphere3d", "step3d"])
def test_single_domain_plan_equals_oracle(name, oracle_mod):
    spec = case(name, 5)
    plan = spec.apply_to(lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega, device=-1)).debug_plan()
    assert plan["n_abb"] > 0 and plan["n_stale"] > 0  # pressure entries; cut cells leave slots nothing writes
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    values = stale_values(plan, o.fold.copy(), spec.ndist)
    dev2ref = np.full(plan["npad"], -1)
    dev2ref[plan["ref2dev"]] = np.arange(plan["n"])
    for _ in range(2):
        o.step(1)
        uext = extrapolated_velocity(plan, lambda n: o.vars[dev2ref[n], :spec.ndim], None, spec.ndim)
        mine = gather(plan, to_device(plan, o.f, spec.ndist), spec.ndist, values=values, uext=uext)
        assert np.array_equal(mine, o.fold)


@pytest.mark.parametrize("name,world", [("sphere3d", 2), ("sphere3d", 8), ("step3d", 2), ("step3d", 8)])
def test_partitioned_plan_equals_single_domain_oracle(name, world, oracle_mod):
    spec = case(name, 5)
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    init_fold = o.fold.copy()
    plans, lps = [], []
    provider = partition.TableRows(spec.nghbr, spec.ndist)
    for r in range(world):
        lp = partition.plan_rank(provider, r, world, spec.nghbr.shape[1], pressure_surfaces(spec))
        s = add_restricted(lbm_b200.Solver(spec.ndim, spec.ndist, lp.nghbr, spec.omega, device=-1), spec, lp)
        lp.apply_halo(s)
        plans.append(s.debug_plan())
        lps.append(lp)
    for _ in range(2):
        o.step(1)
        for r in range(world):
            mine = partitioned_fold(r, plans, lps, o.f, o.vars, init_fold, spec.ndist, spec.ndim)
            assert np.array_equal(mine, o.fold[lps[r].lo:lps[r].hi]), f"rank {r}"
