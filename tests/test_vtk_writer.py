"""The solution file: lbm_b200/host/vtk_writer.hpp against files written by the reference binary.

Input of the writer = moments of the final m_fold (LBMSolver::output recomputes them, /root/reference/src/lbm/solver.cpp:336),
reproduced here by the oracle, which is bit-exact on these cases; output = the bytes of out/<name>_<step>.vtp, compared by
SHA-256 on all 17 reference cases and byte by byte on the two files kept whole (tests/golden/make_vtp_golden.py).
CPU only: this is host-side formatting.
"""
import ctypes
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from casebuilder import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INDEX = json.load(open(os.path.join(HERE, "golden", "vtp", "index.json")))


@pytest.fixture(scope="module")
def hostlib():
    path = os.path.join(ROOT, "lbm_b200", "liblbm_host.so")
    if not os.path.exists(path):
        pytest.skip("liblbm_host.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib = ctypes.CDLL(path)
    lib.lbmhost_write_points.restype = ctypes.c_int
    lib.lbmhost_write_points.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p)]
    lib.lbmhost_round15.restype = None
    lib.lbmhost_round15.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    return lib


def write(lib, path, center, vars_, names, keep=None):
    center = np.ascontiguousarray(center, dtype=np.float64)
    vars_ = np.ascontiguousarray(vars_, dtype=np.float64)
    n, ndim = center.shape
    arr = (ctypes.c_char_p * len(names))(*[s.encode() for s in names])
    kp = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    return lib.lbmhost_write_points(path.encode(), ndim, n, center.ctypes.data, None if kp is None else kp.ctypes.data, vars_.shape[1],
                                    vars_.ctypes.data, arr)


def final_moments(spec, oracle_mod):
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    o.step(int(spec.golden["steps"][-1]))
    o.update_moments()  # output() recomputes m_vars from the final m_fold
    v = np.array(o.vars)
    o.close()
    return v


@pytest.mark.parametrize("name", sorted(INDEX))
def test_solution_file_is_byte_identical_to_the_reference(name, hostlib, oracle_mod, tmp_path):
    spec = load_golden(name)
    v = final_moments(spec, oracle_mod)
    out = str(tmp_path / INDEX[name]["file"])
    # default cell filter "leafCells" (solver.cpp:86, cell_filter.h:60-96): on multi-level grids only the childless cells are written
    keep = ((spec.golden["props"] >> 14) & 1).astype(np.uint8)
    names = ["V"] if spec.equation == "poisson" else ["U", "V", "rho"]  # the potential is written as "V" (solver.cpp:368-378)
    assert write(hostlib, out, spec.center, v, names, None if keep.all() else keep) == 0
    data = open(out, "rb").read()
    whole = os.path.join(HERE, "golden", "vtp", f"{name}.vtp.gz")
    if os.path.exists(whole):
        ref = gzip.open(whole, "rb").read()
        assert len(data) == len(ref)
        diff = next((i for i in range(len(ref)) if data[i] != ref[i]), None)
        assert diff is None, f"first difference at byte {diff}: {data[max(0, diff - 40):diff + 40]!r} vs {ref[max(0, diff - 40):diff + 40]!r}"
    assert len(data) == INDEX[name]["bytes"]
    assert hashlib.sha256(data).hexdigest() == INDEX[name]["sha256"]


# executed steps of the full runs (tests/test_host_run_gpu.py::RUN_SH); the oracle reaches them in seconds for these three
LINE_CASES = [("poisson2D", 10000), ("poissonD2Q9", 10000), ("poiseuille", 34250), ("couette_bnd_eq", 14801), ("couette_bnd_eq_aligned", 14801)]


@pytest.mark.parametrize("name,steps", LINE_CASES)
def test_line_csv_matches_the_reference(name, steps, hostlib, oracle_mod, tmp_path):
    """postprocessing type "line", hook atEnd (postprocessing.h:104-114): ./line.csv of the reference's full run, text for text.
    Cell selection and order come from the host's grid pipeline, u from the oracle's final moments."""
    from lbm_b200 import host_api
    full = json.load(open(os.path.join(HERE, "golden", "vtp", "index_full.json")))
    spec = load_golden(name)
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    o.step(steps)
    o.update_moments()
    v = np.array(o.vars)
    o.close()
    cfg = tmp_path / "case.json"
    cfg.write_text(str(spec.golden["config_orig_json"]))
    out = tmp_path / "line.csv"
    n = host_api.postprocess_line(str(cfg), v, str(out))
    ref = full[name]["line_csv"]
    assert n == ref.count("\n") - 1
    assert out.read_text() == ref


def test_round15_equals_the_decimal_round_trip(hostlib):
    """round15(x) == float('%.15f' % x) (what toStringVector + std::stod do), including ties, signed zeros, subnormals, big values."""
    rng = np.random.default_rng(7)
    x = np.concatenate([
        rng.standard_normal(20000), rng.standard_normal(20000) * 1e-3, rng.standard_normal(20000) * 1e-14, rng.standard_normal(5000) * 1e-17,
        1.0 + rng.standard_normal(20000) * 1e-9, rng.standard_normal(5000) * 50.0, rng.standard_normal(2000) * 1e6, rng.standard_normal(500) * 1e18,
        np.arange(-2000, 2000) / 65536.0,                  # exact ties of the 15th decimal (x * 1e15 ends in .5)
        (2 * np.arange(0, 500) + 1) * 0.5e-15,              # near-ties
        np.array([0.0, -0.0, 5e-324, -5e-324, 2.2e-308, 4.9e-16, 5e-16, 5.1e-16, -4e-16, 9.007199254740991, 9.007199254740993, -9.1,
                  1e300, np.inf, -np.inf, np.nan]),
    ])
    out = np.empty_like(x)
    hostlib.lbmhost_round15(x.ctypes.data, out.ctypes.data, x.size)
    want = np.array([float("%.15f" % v) for v in x])
    same = (out.view(np.uint64) == want.view(np.uint64)) | (np.isnan(out) & np.isnan(want))
    assert same.all(), f"{x[~same][:5]} -> {out[~same][:5]} expected {want[~same][:5]}"


def test_cell_filter_and_padding_rule(hostlib, tmp_path):
    """A filtered file holds only the kept cells, renumbered; the '=' count follows ceil(bytes*8/6) mod 4 (IO.h:395-399)."""
    import base64
    import re
    import struct
    rng = np.random.default_rng(3)
    for n_keep in (1, 2, 3, 4, 5, 6, 7):
        n = 9
        center = rng.random((n, 3))
        v = rng.standard_normal((n, 4))
        keep = np.zeros(n, np.uint8)
        keep[rng.permutation(n)[:n_keep]] = 1
        out = str(tmp_path / f"f{n_keep}.vtp")
        assert write(hostlib, out, center, v, ["U", "V", "W", "rho"], keep) == 0
        s = open(out).read()
        assert f'NumberOfPoints="{n_keep}"' in s
        arrays = re.findall(r'<DataArray ([^>]*format="binary")>\s*\n([A-Za-z0-9+/=]*)\n', s)
        assert len(arrays) == 6
        for attr, text in arrays:
            width = 4 if "Float32" in attr else 8
            count = n_keep * 3 if "Points" in attr else n_keep
            nbytes = 8 + width * count
            chars = -(-nbytes * 8 // 6)
            assert len(text.rstrip("=")) == chars
            assert len(text) - chars == (0 if chars % 4 == 0 else 4 - chars % 4)
            raw = base64.b64decode(text + "=" * (-len(text) % 4))
            assert struct.unpack("<Q", raw[:8])[0] == 8 * count  # element count x 8, not the byte count (base64.h:236)
            if "Name=\"rho\"" in attr:
                got = np.frombuffer(raw[8:8 + 8 * count], dtype=np.float64)
                want = np.array([float("%.15f" % t) for t in v[keep.astype(bool), 3]])
                assert np.array_equal(got, want)
