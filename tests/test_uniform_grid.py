"""The on-demand row provider for single-level grids (lbm_b200/host/uniform_grid.hpp, used by partitioned runs of BASELINE.json
configs[3] / [4]) against the host's full grid pipeline (grid.hpp, itself bit-exact against the reference's dumps,
tests/test_host_grid.py): same cell count, centres, neighbour table incl. composed diagonals, pull sources and boundary surfaces."""
import json

import numpy as np
import pytest

from casebuilder import load_golden
from cases3d import CONFIGS
from lbm_b200 import host_api, partition


def both(cfg, tmp_path):
    path = tmp_path / "case.json"
    path.write_text(json.dumps(cfg))
    import os
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        return host_api.build_grid(str(path)), host_api.UniformGrid(str(path))
    finally:
        os.chdir(cwd)


def check(g, u):
    assert u.n == g["n"] and u.ndim == g["ndim"] and u.cell_length == g["cell_length"]
    ids = np.arange(u.n, dtype=np.int64)
    rows, center = u.rows(ids, want_center=True)
    assert np.array_equal(center, g["center"]), "cell order / centres differ"
    assert np.array_equal(rows, g["nghbr"]), "neighbour table differs"
    # pull sources = explicit inverse of the push table
    qm = g["nghbr"].shape[1]
    inv = partition.TableRows(g["nghbr"], qm + 1).pull
    assert np.array_equal(u.sources(ids)[:, :qm], inv)
    # a scattered subset gives the same rows
    sub = ids[::7][::-1].copy()
    assert np.array_equal(u.rows(sub)[0], g["nghbr"][sub])
    names = [s[0] for s in g["surfaces"]]
    assert [s[0] for s in u.surfaces()] == names
    for (nm, cells, normals), (_, c2, n2) in zip(g["surfaces"], u.surfaces()):
        assert np.array_equal(cells, c2), f"surface {nm}: cell list differs"
        assert np.array_equal(normals, n2), f"surface {nm}: normals differ"


@pytest.mark.parametrize("name,level", [("sphere3d", 4), ("sphere3d", 5), ("step3d", 4), ("step3d", 5)])
def test_3d_cases_equal_the_full_pipeline(name, level, tmp_path):
    g, u = both(CONFIGS[name](level), tmp_path)
    check(g, u)


@pytest.mark.parametrize("name", ["sphere_ns", "step_ns", "couette_bnd"])
def test_reference_2d_cases_equal_the_full_pipeline(name, tmp_path):
    """the reference's own configurations (cell order, tables and surfaces of grid.hpp are pinned by the reference's dumps)"""
    spec = load_golden(name)
    g, u = both(spec.config, tmp_path)
    check(g, u)
    assert np.array_equal(u.rows(np.arange(u.n))[0], spec.golden["nghbr"].astype(np.int64))


@pytest.mark.parametrize("name,world", [("sphere3d", 4), ("step3d", 3)])
def test_partition_from_on_demand_rows_equals_partition_from_the_table(name, world, tmp_path):
    """lbm_b200.partition.GridRows (rows generated per rank) gives every rank the same local problem as TableRows over the full table."""
    from cases3d import build_case, pressure_surfaces
    spec = build_case(name, 5)
    cfg = tmp_path / "case.json"
    cfg.write_text(json.dumps(spec.config))
    u = host_api.UniformGrid(str(cfg))
    for r in range(world):
        a = partition.plan_rank(partition.TableRows(spec.nghbr, spec.ndist), r, world, 26, pressure_surfaces(spec))
        b = partition.plan_rank(partition.GridRows(u, spec.ndist), r, world, 26, pressure_surfaces(spec))
        assert np.array_equal(a.nghbr, b.nghbr) and np.array_equal(a.ghosts, b.ghosts) and a.peers == b.peers
        for f in ("send_cell", "send_dir", "recv_cell", "recv_dir", "vsend_cell", "vrecv_cell"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
        assert a.send_count == b.send_count and a.recv_count == b.recv_count
