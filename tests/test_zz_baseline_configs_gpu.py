"""BASELINE.json configs[3] / configs[4] on the GPU: flow past a sphere (D3Q27, BGK and MRT) and the channel with a step (D3Q19, TRT,
pressure outflow) as 3D cases (tests/cases3d.py) through the C ABI, against the CPU oracle.  STRICT fp64 must be bit-identical
(indices, flags, populations, moments); FAST fp64 within 1e-12 relative (BASELINE.json north_star).  Parity for 3D is unpinned by
the reference (its executable rejects D3Q19 / D3Q27); the oracle follows the dimension-generic source text.
The CPU half (plans of single-domain and partitioned runs) is tests/test_baseline_configs.py."""
import numpy as np
import pytest

import lbm_b200
from cases3d import build_case, mrt_rates

pytestmark = pytest.mark.gpu

CASES = [("sphere3d", 5, lbm_b200.BGK), ("sphere3d", 5, lbm_b200.MRT), ("sphere3d", 6, lbm_b200.MRT),
         ("step3d", 5, lbm_b200.TRT), ("step3d", 6, lbm_b200.TRT)]


def pair(name, level, collision, oracle_mod, **kw):
    spec = build_case(name, level)
    rates = mrt_rates(spec.ndist, spec.omega)
    om_minus = 1.0 / 0.8
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.set_collision(collision, om_minus, rates)
    g = spec.apply_to(lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega, collision=collision, omega_minus=om_minus,
                                     mrt_rates=rates, **kw))
    o.init()
    g.init()
    return spec, o, g


@pytest.mark.parametrize("name,level,collision", CASES)
def test_strict_fp64_bit_identical(name, level, collision, oracle_mod):
    spec, o, g = pair(name, level, collision, oracle_mod)
    assert np.array_equal(g.f, o.f) and np.array_equal(g.fold, o.fold), "initial condition differs"
    for n in (1, 1, 8, 30):
        o.step(n)
        g.step(n)
        for arr in ("f", "fold", "vars", "varsold"):
            a, b = getattr(g, arr), getattr(o, arr)
            assert np.array_equal(a, b), f"{spec.name}: {arr} differs after {g.steps_done} steps, max abs {np.max(np.abs(a - b))}"
    ro, _ = o.residual()
    rg, bad = g.residual()
    assert not bad and np.allclose(rg, ro, rtol=1e-12, atol=1e-300)
    o.update_moments()
    assert np.array_equal(g.moments(), o.vars)
    st = g.stats()
    assert st["slots_stale"] > 0 and st["slots_bc"] > 0
    if level >= 6:
        assert st["cells_fast"] > 0, "template-indexed chunk path not taken"


@pytest.mark.parametrize("name,level,collision", [("sphere3d", 6, lbm_b200.MRT), ("step3d", 6, lbm_b200.TRT)])
def test_fast_fp64_within_1e12(name, level, collision, oracle_mod):
    spec, o, g = pair(name, level, collision, oracle_mod, arithmetic=lbm_b200.FAST)
    o.step(100)
    g.step(100)
    scale = np.max(np.abs(o.f))
    assert np.max(np.abs(g.f - o.f)) / scale < 1e-12
    assert np.max(np.abs(g.fold - o.fold)) / scale < 1e-12
    assert np.max(np.abs(g.vars[:, -1] - o.vars[:, -1])) < 1e-12                      # density
    assert np.max(np.abs(g.vars[:, :-1] - o.vars[:, :-1])) < 1e-12 / np.sqrt(3.0)    # velocity: 1e-12 of the lattice speed of sound
