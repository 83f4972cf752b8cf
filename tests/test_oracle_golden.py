"""Pin the CPU oracle against the reference's own output.

The golden fixtures are raw dumps of the reference binary (tests/golden/make_golden.py); the oracle has to
reproduce m_fold, m_f, m_vars and m_varsold BIT FOR BIT at every dumped step, on every case.
"""
import hashlib

import numpy as np
import pytest

from casebuilder import load_golden, omega_from_config

CASES = ["couette", "couette_bnd", "couette_bnd_bbDirichlet", "poiseuille", "poiseuille_bnd", "step_ns", "sphere_ns",
         # wet-node wall family + the remaining Navier-Stokes cases of the reference's test/run.sh
         "couette_bnd_eq", "couette_bnd_eq2", "couette_bnd_eq_aligned", "couette_bnd_NEEM", "couette_bnd_NEBB",
         "poiseuille_bnd_eq", "poiseuille_bnd_NEEM", "poiseuille_bnd_NEBB", "poiseuille_bnd_pressure",
         "poiseuille_bnd_pressure_neem2",
         # multi-level grids: every level is stepped as its own lattice
         "couette_ml_p3u5", "couette_ml_u5m6", "couette_ml_p4u5m7", "sphere_ml_p4u6", "step_ml_p3u5",
         # Poisson equation types: the five Poisson cases of the reference's test/run.sh
         "poisson1D", "poisson1D_reaction", "poisson2D", "poisson2D_helmholtz", "poissonD2Q9", "poisson2D_reaction", "step_poisson"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_bit_for_bit(name, oracle_mod):
    spec = load_golden(name)
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    done = 0
    for s in spec.golden["steps"]:
        o.step(int(s) - done)
        done = int(s)
        for arr in ("fold", "f", "vars", "varsold"):
            mine = getattr(o, arr)
            assert sha(mine) == spec.digests[f"{arr}_{s}"], f"{name}: {arr} differs from the reference at step {s}"
            key = f"{arr}_{s}"
            if key in spec.golden:
                assert np.array_equal(mine, spec.golden[key])
    o.close()


@pytest.mark.parametrize("name", [c for c in CASES if not c.startswith("poisson")])
def test_omega_from_config(name):
    """omega derivation, src/lbm/solver.cpp:108-123, against the value the reference printed."""
    spec = load_golden(name)
    assert omega_from_config(spec.config["solver"], int(spec.golden["maxlvl"])) == spec.omega
