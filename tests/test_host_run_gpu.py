"""End to end like the reference's test/run.sh: the reference's OWN configuration files through the host mirror
(grid generator -> transferGrid -> LBMSolver::run on the GPU), pass/fail decided by the in-solver analytic thresholds
(src/lbm/solver.cpp:457-481), and the printed numbers compared with what the reference binary prints."""
import json

import numpy as np
import pytest

from casebuilder import load_golden
from lbm_b200 import host_api

pytestmark = pytest.mark.gpu


def run_case(name, tmp_path, **override):
    spec = load_golden(name)
    cfg = json.loads(str(spec.golden["config_orig_json"]))
    cfg["solver"]["output_dir"] = str(tmp_path / "out")
    cfg["solver"].update(override)
    path = tmp_path / "case.json"
    path.write_text(json.dumps(cfg))
    n = int(spec.golden["ncells"]) * (spec.ndim + 1)
    return host_api.run(str(path), nvars=n), spec


def test_couette_reference_case_converges_like_the_reference(tmp_path):
    (rc, msg, out, vars_), spec = run_case("couette", tmp_path)
    assert rc == 0, msg
    # reference: converged at step 1300 (1301 steps executed), residual 1.41541e-11, max error 1.7692e-12 (limit 4e-11)
    assert out["converged"] == 1.0 and out["steps"] == 1301
    assert abs(out["residual"] - 1.41541e-11) < 1e-16
    assert abs(out["max_error"] - 1.7692e-12) < 1e-16
    assert (tmp_path / "out" / "couette_1300.vtp").exists()
    # analytic Couette profile u = 0.1 * y / 5 on every cell
    v = vars_.reshape(-1, 3)
    assert np.max(np.abs(v[:, 0] - 0.1 / 5.0 * spec.center[:, 1])) < 4e-11


def test_poiseuille_reference_case_meets_its_thresholds(tmp_path):
    (rc, msg, out, _), _ = run_case("poiseuille", tmp_path, solution_interval=10 ** 9)
    assert rc == 0, msg
    # reference: 34 250 steps (cap), max error 1.26064e-07 (limit 1.3e-7), L2 3.23436e-05 (limit 3.3e-5)
    assert out["steps"] == 34250
    assert abs(out["max_error"] - 1.26064e-07) < 1e-12
    assert abs(out["l2_error"] - 3.23436e-05) < 1e-10


def test_couette_dirichlet_bb_case(tmp_path):
    (rc, msg, out, _), _ = run_case("couette_bnd_bbDirichlet", tmp_path)
    assert rc == 0, msg
    # reference: converged at 1300, max error 1.76069e-12
    assert out["steps"] == 1301 and abs(out["max_error"] - 1.76069e-12) < 1e-16


@pytest.mark.parametrize("name,steps,max_error", [
    ("couette_bnd_eq", 14801, 1.39587e-12),          # wet-node equilibrium wall + bounce-back lid, periodic BC
    ("couette_bnd_NEBB", 14801, 6.30746e-13),        # non-equilibrium bounce-back walls
    ("poiseuille_bnd_pressure", 122801, 9.98514e-06),  # pressure in/outlet + equilibrium walls on an aligned grid
])
def test_wet_node_reference_cases_end_to_end(name, steps, max_error, tmp_path):
    """The numbers are what the reference binary prints for its own configuration."""
    (rc, msg, out, _), _ = run_case(name, tmp_path, solution_interval=10 ** 9)
    assert rc == 0, msg
    assert out["converged"] == 1.0 and out["steps"] == steps
    assert abs(out["max_error"] - max_error) < 2e-6 * max_error


def test_failed_threshold_terminates_like_termm(tmp_path):
    (rc, msg, out, _), _ = run_case("couette", tmp_path, errorMax=1e-20)
    assert rc == -1 and "Analytical testcase failed" in msg


def test_invalid_wall_model_is_reported(tmp_path):
    spec = load_golden("couette")
    cfg = json.loads(str(spec.golden["config_orig_json"]))
    cfg["solver"]["boundary"]["cube"]["-y"] = {"type": "wall", "model": "slippery"}
    p = tmp_path / "c.json"
    p.write_text(json.dumps(cfg))
    rc, msg, _, _ = host_api.run(str(p))
    assert rc == -1 and "Invalid wall boundary model: slippery" in msg
