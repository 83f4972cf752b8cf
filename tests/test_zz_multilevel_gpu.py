"""Multi-level grids on the GPU.  The reference keeps the cells of every level from the partition level up
in one list and steps each level as its own lattice (solver.cpp:525 "todo: skip non-leaf cells"), so the time step is table-driven
exactly as on a single level; the fixtures are dumps of the reference binary on its own couette / sphere / step configurations with
only the three level keys changed (tests/golden/make_golden.py).  (a) STRICT fp64 through the C ABI is bit-identical to those dumps;
(b) the host mirror (grid generator -> LBMSolver::run -> leaf-filtered solution file) leaves the reference's file, byte for byte."""
import hashlib
import json
import os

import numpy as np
import pytest

import lbm_b200
from casebuilder import load_golden
from lbm_b200 import host_api

pytestmark = pytest.mark.gpu

CASES = ["couette_ml_p3u5", "couette_ml_u5m6", "couette_ml_p4u5m7", "sphere_ml_p4u6", "step_ml_p3u5"]
VTP = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vtp", "index.json")))


@pytest.mark.parametrize("name", CASES)
def test_strict_fp64_is_bit_identical_to_the_reference_dump(name, oracle_mod):
    spec = load_golden(name)
    g = spec.apply_to(lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    g.init()
    done = 0
    for s in spec.golden["steps"]:
        g.step(int(s) - done)
        done = int(s)
        for arr in ("f", "fold", "vars", "varsold"):
            a = getattr(g, arr)
            assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == spec.digests[f"{arr}_{s}"], f"{name} step {s}: {arr}"


@pytest.mark.parametrize("name", CASES)
def test_host_run_writes_the_reference_solution_file(name, tmp_path):
    spec = load_golden(name)
    cfg = json.loads(str(spec.golden["config_json"]))  # the shortened run the fixture was dumped from (50 steps, no analytic test)
    cfg["solver"]["output_dir"] = str(tmp_path / "out")
    path = tmp_path / "case.json"
    path.write_text(json.dumps(cfg))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        rc, msg, out, vars_ = host_api.run(str(path), nvars=int(spec.golden["ncells"]) * (spec.ndim + 1))
    finally:
        os.chdir(cwd)
    assert rc == 0, msg
    assert out["steps"] == int(spec.golden["steps"][-1])
    ref = VTP[name]
    written = tmp_path / "out" / ref["file"]
    assert written.exists(), sorted(p.name for p in (tmp_path / "out").iterdir())
    data = written.read_bytes()
    assert len(data) == ref["bytes"] and hashlib.sha256(data).hexdigest() == ref["sha256"]
