"""CPU-side tests: the C ABI loads and exports what include/lbm_b200.h declares, the host-side grid code agrees with the
reference's known answers and with the brute-force numpy restatement, and the library refuses to compute without a GPU."""
import json
import os

import numpy as np
import pytest

import lbm_b200
from gridgen import box_grid, sfc_key, sfc_key_from_unit
from lbm_b200.capi import box_topology, sfc_index

HERE = os.path.dirname(os.path.abspath(__file__))


def test_library_exports_every_declared_symbol():
    lib = lbm_b200.load_library()
    names = lbm_b200.abi_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"liblbm_b200.so does not export {n}"
    assert lib.lbm_b200_abi_version() == 1


def test_no_cpu_fallback():
    """Without a CUDA device creating a solver must fail loudly (there is no CPU path in the product)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    nghbr = np.full((16, 8), -1, dtype=np.int64)
    with pytest.raises(lbm_b200.LbmB200Error) as e:
        lbm_b200.Solver(2, 9, nghbr, 1.0)
    assert e.value.code == -3


def test_sfc_index_reference_known_answers():
    """UnitTest/test_hilbert.cpp of the reference: every EXPECT_EQ, for the library's host code and the numpy restatement."""
    kats = json.load(open(os.path.join(HERE, "golden", "hilbert_kat.json")))["kats"]
    assert len(kats) >= 100
    for k in kats:
        assert sfc_index(k["x"], k["level"]) == k["index"], k
        assert sfc_key_from_unit(k["x"], k["level"]) == k["index"], k


def test_integer_key_equals_unit_cube_key():
    rng = np.random.default_rng(1)
    for ndim, level in ((2, 5), (3, 4), (3, 9)):
        c = rng.integers(0, 2 ** level, size=(200, ndim))
        keys = sfc_key(c, level)
        for row, key in zip(c, keys):
            assert sfc_index((row + 0.5) / 2 ** level, level) == key


@pytest.mark.parametrize("shape,periodic", [((4, 4), (1, 0)), ((12, 7), (0, 1)), ((33, 20), (0, 0)), ((8, 8, 8), (1, 0, 0)),
                                            ((5, 9, 6), (1, 1, 0)), ((16, 16, 16), (1, 1, 1)), ((20, 12, 9), (0, 0, 0))])
def test_box_topology_matches_numpy_restatement(shape, periodic):
    nb, ce, co = box_topology(shape, periodic, True, True)
    g = box_grid(shape, [bool(p) for p in periodic])
    assert np.array_equal(nb, g["nghbr"])
    assert np.array_equal(co, g["coords"])
    assert np.allclose(ce, g["center"], rtol=0, atol=0)


def test_box_order_is_ascending_key():
    _, _, co = box_topology((24, 17, 9), (0, 0, 0), False, True)
    keys = sfc_key(co, 5)
    assert np.all(np.diff(keys) > 0)


def test_push_table_invariants_of_periodic_box():
    nb, _, _ = box_topology((8, 8, 8), (1, 1, 1), False, False)
    opp = [1, 0, 3, 2, 5, 4, 9, 8, 7, 6, 13, 12, 11, 10, 17, 16, 15, 14, 25, 24, 23, 22, 21, 20, 19, 18]
    n = nb.shape[0]
    for i in range(26):
        assert sorted(nb[:, i]) == list(range(n))              # every direction is a permutation
        assert np.array_equal(nb[nb[:, i], opp[i]], np.arange(n))  # and the opposite direction inverts it


def test_navierstokespoisson_ends_like_the_reference(tmp_path):
    """`solver.equation: "navierstokespoisson"` parses (src/lbm/constants.h:43) but every Navier_Stokes_Poisson case of the executor and every
    instantiation is commented out (solverExe.h:45-47,62-64,80-82; solver_inst_*.cpp), so the reference ends in
    TERMM(-1, "Unsupported equation type") (solverExe.h:86; confirmed on the reference binary) -- there is no such solver to build."""
    import json
    import subprocess
    from casebuilder import load_golden
    cfg = json.loads(str(load_golden("couette").golden["config_orig_json"]))
    cfg["solver"]["equation"] = "navierstokespoisson"
    (tmp_path / "case.json").write_text(json.dumps(cfg))
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lbm_b200", "lbm")
    r = subprocess.run([exe, "case.json"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 255 and "Unsupported equation type" in r.stderr


TIMER_LINE = r"^ *\[\d{2,3}\.\d%\] .{1,44}? +[0-9.e+-]+ \[sec\]$"


def read_run_log(path):
    """messages of a reference-style log file (include/common/log.h): well-formed XML, <meta> entries, <m d="0" >text\\n</m> elements"""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    assert root.tag == "root"
    meta = {m.get("name"): m.get("content") for m in root.findall("meta")}
    assert {"noDomains", "dateCreation", "fileFormatVersion", "user", "host", "dir", "executionCommand", "revision", "build", "dateClosing"} <= set(meta)
    msgs = [m.text for m in root.findall("m")]
    assert all(m.get("d") == "0" for m in root.findall("m")) and all(t.endswith("\n") for t in msgs)
    return meta, [t[:-1] for t in msgs]


def test_lbm_executable_leaves_the_reference_style_run_logs(tmp_path):
    """`gridgen_log` and `lbm_log` in the working directory (lbm_b200/host/run_log.hpp): XML envelope, messages and
    the timer table in the reference's layout.  Without a device the solver stops at lbm_b200_create (exit status 255, like TERMM) and the
    logs are still closed properly; with one the run goes through and lbm_log ends with the whole table."""
    import json
    import re
    import subprocess
    from casebuilder import load_golden
    cfg = json.loads(str(load_golden("couette").golden["config_orig_json"]))
    (tmp_path / "case.json").write_text(json.dumps(cfg))
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lbm_b200", "lbm")
    r = subprocess.run([exe, "case.json"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode in (0, 255), r.stderr[-1500:]
    meta, msgs = read_run_log(tmp_path / "gridgen_log")
    assert meta["noDomains"] == "1" and meta["executionCommand"].endswith("lbm case.json") and meta["dir"] == str(tmp_path)
    assert msgs[0] == "Grid generator started ||>" and "Loading configuration file [case.json]" in msgs and "Generating a grid[2D]" in msgs
    assert "      * grid has 640 cells" in msgs and "Grid generator finished <||" in msgs
    table = msgs[msgs.index("-" * 80):]
    assert table[1] == "Group".ljust(50) + "Application".ljust(40)
    names = [re.sub(r"^ *\[[0-9.]+%\] ", "", t)[:-26].rstrip() for t in table[2:]]
    assert names == ["Total", "Total run time of the grid generator", "Init", "Create the grid.", "Grid IO."]   # the solver's timers do not exist yet
    assert all(re.match(TIMER_LINE, t) for t in table[2:]), table
    assert table[2].startswith("[100.0%] Total") and table[3].startswith("  [") and table[4].startswith("    [")
    assert len(table[2]) == 50 + 20 + len(" [sec]")
    meta, msgs = read_run_log(tmp_path / "lbm_log")
    assert msgs[:3] == ["2D LBM Solver started ||>", "Loading configuration file [case.json]", "Transferring 2D Grid to LBM solver"]
    if r.returncode == 0:
        assert "Reached convergence to: 1.41541e-11" in msgs and "max. Error: 1.7692e-12" in msgs and "LBM Solver finished <||" in msgs
        assert "1300: dU=1.41541e-11 dV=2.44249e-14 drho=5.77316e-14 " in msgs       # the reference's lbm_log holds the same line
        assert any(m.startswith("  Writing ") and m.endswith("couette_1300.vtp with #640 cells") for m in msgs)
        table = msgs[msgs.index("-" * 80):]
        names = [re.sub(r"^ *\[[0-9.]+%\] ", "", t)[:-26].rstrip() for t in table[2:]]
        assert names == ["Total", "Total run time of the grid generator", "Init", "Create the grid.", "Grid IO.", "Total run time of the LBM Solver.",
                         "Initialization of the LBM solver!", "Main Loop of the LBM solver!", "Computation", "Postprocessing", "IO"]
        assert all(re.match(TIMER_LINE, t) for t in table[2:]), table


def test_run_log_escapes_the_five_xml_characters(tmp_path):
    """log.h:45-75"""
    import ctypes as C
    import subprocess
    src = tmp_path / "t.cpp"
    src.write_text('#include "run_log.hpp"\nint main(int argc, char** argv) { lbmhost::RunLog l; l.open("x_log", argc, argv); '
                   'l("a<b & \\"c\\" > \'d\'"); return 0; }\n')
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lbm_b200", "host")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", host, str(src), "-o", str(tmp_path / "t")])
    subprocess.check_call([str(tmp_path / "t"), "<arg>"], cwd=tmp_path)
    text = (tmp_path / "x_log").read_text()
    assert "a&lt;b &amp; &quot;c&quot; &gt; &apos;d&apos;\n</m>" in text and "&lt;arg&gt;" in text
    meta, msgs = read_run_log(tmp_path / "x_log")
    assert msgs == ["a<b & \"c\" > 'd'"]
