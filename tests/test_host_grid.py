"""The host topology builder (lbm_b200/host/grid.hpp, C++) against dumps of the reference's own grid pipeline:
cell order / centres (G1), axis neighbours incl. grid-level periodic links (G2, H9), property bits (G3), boundary surfaces with
per-entry normals in creation order (G4) and diagonal neighbours (G5) must be BIT-EXACT."""
import json
import os

import numpy as np
import pytest

from casebuilder import load_golden
from lbm_b200 import host_api

CASES = ["couette", "couette_bnd", "couette_bnd_bbDirichlet", "poiseuille", "poiseuille_bnd", "step_ns", "sphere_ns",
         "couette_bnd_eq", "couette_bnd_eq2", "couette_bnd_eq_aligned", "couette_bnd_NEEM", "couette_bnd_NEBB",
         "poiseuille_bnd_eq", "poiseuille_bnd_NEEM", "poiseuille_bnd_NEBB", "poiseuille_bnd_pressure",
         "poiseuille_bnd_pressure_neem2",  # the *_aligned / poiseuille_bnd_* cases use alignNodesWithSurface
         # multi-level grids: partitionLevel < uniformLevel and / or boundary refinement
         "couette_ml_p3u5", "couette_ml_u5m6", "couette_ml_p4u5m7", "sphere_ml_p4u6", "step_ml_p3u5",
         # Poisson cases of test/run.sh: 1D grids (D1Q3; the curve's 1D keys are 0, 3, 6, ...), aligned 1D / 2D grids
         "poisson1D", "poisson1D_reaction", "poisson2D", "poisson2D_helmholtz", "poissonD2Q9", "poisson2D_reaction", "step_poisson"]


@pytest.mark.parametrize("name", CASES)
def test_host_grid_tables_bit_exact(name, tmp_path):
    spec = load_golden(name)
    cfg = tmp_path / "case.json"
    cfg.write_text(json.dumps(spec.config))
    g = host_api.build_grid(str(cfg))
    gold = spec.golden
    assert g["n"] == int(gold["ncells"])
    assert np.array_equal(g["center"], gold["center"]), "cell order / centres differ"
    assert np.array_equal(g["nghbr"], gold["nghbr"].astype(np.int64)), "neighbour table differs"
    assert np.array_equal(g["props"], gold["props"]), "property bits differ"
    names = [str(s) for s in gold["surface_names"]]
    assert [s[0] for s in g["surfaces"]] == names
    for k, (sname, cells, normals) in enumerate(g["surfaces"]):
        assert np.array_equal(cells, gold[f"surf{k}_cells"].astype(np.int64)), f"surface {sname}: cell list differs"
        assert np.array_equal(normals, gold[f"surf{k}_normals"]), f"surface {sname}: normals differ"
    assert g["cell_length"] == spec.cell_length


def test_reference_configurations_that_abort_abort_the_same_way(tmp_path, monkeypatch):
    """test/poiseuille/poiseuille_eq.json and test/step/step_ns_double.json name periodic connections that do not exist: the reference
    ends with TERMM(-1, "Invalid periodic setup!") (src/cartesiangrid.h:636).  Same message here."""
    broken = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "broken_configs.json")))
    monkeypatch.chdir(tmp_path)
    for name, item in broken.items():
        cfg = tmp_path / f"{name}.json"
        cfg.write_text(json.dumps(item["config"]))
        with pytest.raises(RuntimeError) as e:
            host_api.build_grid(str(cfg))
        assert item["error"] in str(e.value), name
