"""The C++ host's multi-GPU set-up of one rank (lbm_b200/host/lbm_solver.hpp: setupGpuPartitioned -- on-demand grid rows, native
partition, restricted boundary conditions, halo lists) against the Python path that bench.py and the NCCL tests use
(lbm_b200.partition + lbm_b200.cases): the device plans must be identical, array by array.  CPU only (inspection handles); what is left
for hardware is the NCCL bootstrap through the id file and the run loop."""
import json

import numpy as np
import pytest

import lbm_b200
from lbm_b200 import cases, host_api, partition

ARRAYS = ("ref2dev", "tmpl", "chunk_nb", "codes", "copytab", "addtab", "wall_desc", "abb_p", "abb_cells", "values", "stale_ref", "send_index",
          "recv_index", "vsend_cells", "chunk_abb_base", "chunk_abb")


def python_plan(cfg, path, rank, world):
    ndim, ndist = cfg["dim"], cases.NDIST[cfg["solver"]["model"]]
    u = host_api.UniformGrid(str(path))
    surfaces = {nm: (c, n) for nm, c, n in u.surfaces()}
    bcs, _ = cases.bcs_from_config(cfg["solver"], surfaces, ndim)
    pressure = [(bc["cells"], bc["normals"]) for bc in bcs if bc["kind"] == "pressure"]
    lp = partition.plan_rank(partition.GridRows(u, ndist), rank, world, u.stride, pressure)
    s = lbm_b200.Solver(ndim, ndist, lp.nghbr, 1.0 / float(cfg["solver"]["relaxation"]), device=-1)
    cases.apply_bcs(s, cases.restrict_bcs(bcs, lp))
    lp.apply_halo(s)
    return s.debug_plan()


@pytest.mark.parametrize("name,level,world", [("sphere3d", 5, 2), ("sphere3d", 5, 4), ("step3d", 5, 3), ("step3d", 6, 8)])
def test_cpp_host_rank_setup_equals_the_python_path(name, level, world, tmp_path, monkeypatch):
    cfg = cases.CONFIGS[name](level)
    path = tmp_path / "case.json"
    path.write_text(json.dumps(cfg))
    monkeypatch.chdir(tmp_path)
    ranks = range(world) if world <= 4 else (0, 3, 7)
    for r in ranks:
        mine = host_api.partitioned_plan(path, r, world, 3, cases.NDIST[cfg["solver"]["model"]])
        ref = python_plan(cfg, path, r, world)
        for k, v in ref.items():
            if isinstance(v, int):
                assert mine[k] == v, (k, mine[k], v)
        for k in ARRAYS:
            assert np.array_equal(mine[k], ref[k]), f"rank {r}: {k}"
        assert mine["n_send"] > 0 and mine["n_abb"] >= 0


def test_unsupported_configurations_are_refused(tmp_path, monkeypatch):
    cfg = cases.CONFIGS["sphere3d"](4)
    cfg["solver"]["boundary"]["cube"]["+z"] = {"type": "wall", "model": "equilibrium"}
    path = tmp_path / "case.json"
    path.write_text(json.dumps(cfg))
    monkeypatch.chdir(tmp_path)
    with pytest.raises(RuntimeError) as e:
        host_api.partitioned_plan(path, 0, 2, 3, 27)
    assert "not partitioned" in str(e.value)


def test_lbm_executable_reads_its_rank_from_the_environment(tmp_path):
    """`lbm` as rank 1 of 2: the grid generator skips the whole-tree build, the solver partitions through the on-demand provider and
    asks for CUDA device LOCAL_RANK -- on a machine without a GPU that is where it must stop, loudly (no CPU fallback)."""
    import os
    import subprocess

    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run would wait for its peer")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "lbm_b200", "lbm")
    path = tmp_path / "case.json"
    path.write_text(json.dumps(cases.CONFIGS["step3d"](4)))
    env = dict(os.environ, LBM_B200_RANK="1", LBM_B200_WORLD="2", LBM_B200_LOCAL_RANK="1", LBM_B200_ID_FILE=str(tmp_path / "id"))
    r = subprocess.run([exe, str(path)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 255
    assert "partitioned run: grid rows are generated per rank" in r.stdout
    assert "Rank 1 of 2: partitioned run" in r.stderr and "no CUDA device 1" in r.stderr
    # the same file without rank variables is a single-process run and asks for device 0
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, str(path)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 255 and "no CUDA device 0" in r.stderr and "partitioned run" not in r.stdout
