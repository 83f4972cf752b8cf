"""End to end like the reference's test/run.sh: the reference's OWN configuration files through the host mirror
(grid generator -> transferGrid -> LBMSolver::run on the GPU), pass/fail decided by the in-solver analytic thresholds
(src/lbm/solver.cpp:457-481), and the printed numbers compared with what the reference binary prints."""
import hashlib
import json
import os

import numpy as np
import pytest

from casebuilder import load_golden
from lbm_b200 import host_api

pytestmark = pytest.mark.gpu

# final solution file of each full run as the reference binary writes it (tests/golden/make_vtp_golden.py --full)
VTP_FULL = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vtp", "index_full.json")))


def run_case(name, tmp_path, **override):
    spec = load_golden(name)
    cfg = json.loads(str(spec.golden["config_orig_json"]))
    cfg["solver"]["output_dir"] = str(tmp_path / "out")
    cfg["solver"].update(override)
    path = tmp_path / "case.json"
    path.write_text(json.dumps(cfg))
    n = int(spec.golden["ncells"]) * (spec.ndim + 1)
    cwd = os.getcwd()
    os.chdir(tmp_path)  # postprocessing writes ./line.csv, like the reference
    try:
        return host_api.run(str(path), nvars=n), spec
    finally:
        os.chdir(cwd)


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


# The reference's test/run.sh:79-118 runs these 15 Navier-Stokes configurations and passes when the process exits with 0, i.e.
# when the in-solver analytic thresholds hold.  Expected step counts and errors are what the reference binary prints
# ("conv at N" means N+1 executed steps, the loop runs the step in which convergence is detected).
RUN_SH = [
    # name, executed steps, converged, max error, L2 error (None: the reference prints 0 or the case has no L2 limit worth pinning)
    ("couette", 1301, True, 1.7692e-12, None),
    ("couette_bnd", 1301, True, 1.7692e-12, None),
    ("couette_bnd_bbDirichlet", 1301, True, 1.76069e-12, None),
    ("couette_bnd_eq", 14801, True, 1.39587e-12, None),
    ("couette_bnd_eq2", 14801, True, 6.3069e-13, None),
    ("couette_bnd_NEBB", 14801, True, 6.30746e-13, None),
    ("couette_bnd_eq_aligned", 14801, True, 6.30607e-13, None),
    ("couette_bnd_NEEM", 20000, False, 3.45342e-15, None),
    ("poiseuille", 34250, False, 1.26064e-07, 3.23436e-05),
    ("poiseuille_bnd", 75000, False, 7.87881e-06, 2.04532e-03),
    ("poiseuille_bnd_eq", 164301, True, 7.95188e-06, None),
    ("poiseuille_bnd_NEBB", 164501, True, 7.42901e-06, None),
    ("poiseuille_bnd_NEEM", 175001, True, 2.73927e-06, 6.76952e-04),
    ("poiseuille_bnd_pressure", 122801, True, 9.98514e-06, None),
    ("poiseuille_bnd_pressure_neem2", 174001, True, 1.0049e-05, None),
]


@pytest.mark.parametrize("name,steps,converged,max_error,l2_error", RUN_SH)
def test_reference_run_sh_case(name, steps, converged, max_error, l2_error, tmp_path):
    (rc, msg, out, _), _ = run_case(name, tmp_path, solution_interval=10 ** 9)
    assert rc == 0, msg                                  # the reference's own pass criterion
    assert out["steps"] == steps and out["converged"] == float(converged)
    assert abs(out["max_error"] - max_error) <= 6e-6 * max_error   # the reference prints 6 significant digits
    if l2_error is not None:
        assert abs(out["l2_error"] - l2_error) <= 6e-6 * l2_error
    # the solution file the run leaves behind: same name, same bytes as the reference's (fields bit-identical, writer byte-identical)
    ref = VTP_FULL[name]
    written = tmp_path / "out" / ref["file"]
    assert written.exists(), sorted(p.name for p in (tmp_path / "out").iterdir())
    data = written.read_bytes()
    assert len(data) == ref["bytes"] and hashlib.sha256(data).hexdigest() == ref["sha256"]
    if "line_csv" in ref:  # postprocessing type "line", hook atEnd
        assert (tmp_path / "line.csv").read_text() == ref["line_csv"]


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


def test_diverging_run_ends_like_the_reference(tmp_path):
    """src/lbm/solver.cpp:254-260, 327-333, 201-203: a NaN / Inf in the residual sets m_diverged; the loop stops, output() writes one more
    solution file with the suffix "bdiv", and run() ends in TERMM(-1, "Solution diverged") -- exit status 255.  Configuration: the
    reference's sphere case at level 5 with an absurd inlet pressure (3.0) and relaxation 0.501; the REFERENCE BINARY run on it here
    (oracle/_ref/lbm_ref) detects the divergence at step 500 and leaves out/solution_500bdiv.vtp."""
    import subprocess
    cfg = json.loads(str(load_golden("sphere_ns").golden["config_orig_json"]))
    for k in ("partitionLevel", "uniformLevel", "maxRfnmtLvl"):
        cfg[k] = 5
    s = cfg["solver"]
    s.pop("reynoldsnumber")
    s.pop("postprocessing", None)
    s.update(relaxation=0.501, maxSteps=4000, info_interval=100, solution_interval=1000000)
    s["boundary"]["cube"]["-x"]["pressure"] = 3.0
    (tmp_path / "case.json").write_text(json.dumps(cfg))
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lbm_b200", "lbm")
    r = subprocess.run([exe, "case.json"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 255, (r.returncode, r.stderr[-1500:])
    assert "Solution diverged!" in r.stderr and "Solution diverged" in r.stderr.split("Solution diverged!")[-1]
    assert (tmp_path / "out" / "solution_500bdiv.vtp").exists(), os.listdir(tmp_path / "out")
