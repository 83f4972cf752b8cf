"""bench.py's JSON line (the driver's contract): the reference arm runs on the CPU, so its line can be checked here; the fields of the
GPU arm that do not depend on a device (config, metric names) come from the same helpers."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [["--size", "32"], ["--workload", "step"]])
def test_reference_arm_prints_one_json_line(extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3", "--cpu-size", "32", *extra],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MLUPS" and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                                   # BASELINE.md holds no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if "--size" in extra:
        assert "32^3" in d["config"]["workload"] and "32^3" in cb["sample"]      # the arm runs the size it prints


def test_reference_arm_never_loads_the_cuda_library():
    """the box workload of the reference arm builds its tables in numpy (bench.box_table_numpy) -- equal to lbm_b200_box_topology"""
    code = ("import sys, json; sys.argv=['bench.py','--impl','reference','--size','16','--steps','3','--warmup','3']; import runpy;"
            "runpy.run_path('bench.py', run_name='__main__'); import ctypes;"
            "maps=open('/proc/self/maps').read(); assert 'liblbm_b200' not in maps, 'CUDA library loaded'; assert 'liblbm_oracle' in maps")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("shape,ndist", [((16, 8, 8), 19), ((8, 8, 8), 27), ((32, 16), 9)])
def test_numpy_box_table_equals_the_native_one(shape, ndist):
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from lbm_b200.capi import box_topology
    periodic = (1,) + (0,) * (len(shape) - 1)
    native = box_topology(shape, periodic)[0]
    assert np.array_equal(bench.box_table_numpy(shape, periodic, ndist), native[:, :ndist - 1])


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--size", "16", "--no-cpu", "--no-e2e"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and r.stdout.strip() == ""              # no CPU fallback, no JSON line


def test_output_path_sample_fails_with_an_ordinary_exception_without_a_device():
    """bench.py wraps its optional output-path measurement (output_path field) in try/except: whatever goes wrong there must arrive as a
    Python exception, never take the bench line with it.  Without a device the page-locked allocation is the first thing to fail."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sys.path.insert(0, ROOT)
    import bench
    import lbm_b200
    nghbr = np.full((64, 8), -1, dtype=np.int64)
    s = lbm_b200.Solver(2, 9, nghbr, 1.2, device=-1)
    with pytest.raises(Exception) as e:
        bench.output_path_sample(s)
    assert "cudaMallocHost" in str(e.value)
