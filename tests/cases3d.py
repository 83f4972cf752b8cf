"""3D versions of the reference's sphere and step cases (test helper): BASELINE.json configs[3] and configs[4].

The reference's own test/sphere/sphere_ns.json and test/step/step_ns.json are 2D (D2Q9); its executable rejects D3Q19 / D3Q27
(src/lbm/solverExe.h:37-90).  The configurations below keep their geometry objects, boundary keys and pressure values and add the
third dimension (walls on -z / +z); the host-side grid pipeline (lbm_b200/host, bit-exact against the reference's 2D dumps) turns
them into the tables the solver consumes.  Parity for these cases is UNPINNED by the reference: the CUDA path is
compared with the CPU oracle, which follows the reference's dimension-generic source text.
"""
import json
import os
import tempfile

import numpy as np

from casebuilder import CaseSpec, bcs_from_config, geometry_bbox
from lbm_b200.cases import CONFIGS, NDIST, sphere3d_config, step3d_config  # noqa: F401





def build_case(name, level, model=None):
    """Configuration -> host grid pipeline -> CaseSpec (tables + boundary conditions in the reference's application order)."""
    from lbm_b200 import host_api
    cfg = CONFIGS[name](level) if model is None else CONFIGS[name](level, model)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, f"{name}.json")
        with open(path, "w") as fh:
            json.dump(cfg, fh)
        cwd = os.getcwd()
        os.chdir(tmp)  # the grid generator writes its log into the working directory, like the reference
        try:
            g = host_api.build_grid(path)
        finally:
            os.chdir(cwd)
    ndim = 3
    surfaces = {nm: (cells, normals) for nm, cells, normals in g["surfaces"]}
    lo, hi = geometry_bbox(cfg["geometry"], ndim)
    spec = CaseSpec(name=f"{name}_l{level}", ndim=ndim, ndist=NDIST[cfg["solver"]["model"]], nghbr=g["nghbr"],
                    omega=1.0 / float(cfg["solver"]["relaxation"]), center=g["center"], bbmin=lo, bbmax=hi, cell_length=g["cell_length"])
    spec.bcs, spec.forcing = bcs_from_config(cfg["solver"], surfaces, ndim)
    spec.config = cfg
    spec.surfaces = surfaces
    return spec


def pressure_surfaces(spec):
    return [(bc["cells"], bc["normals"]) for bc in spec.bcs if bc["kind"] == "pressure"]


def add_restricted(solver, spec, lp):
    """the boundary-condition entries of the cells rank `lp.rank` owns, local ids, order kept"""
    for bc in spec.bcs:
        cells, normals = lp.restrict(bc["cells"], bc["normals"])
        if len(cells) == 0:
            continue
        if bc["kind"] == "pressure":
            solver.add_pressure(cells, normals, bc["pressure"])
        elif bc["kind"] == "wall_bb":
            solver.add_wall_bb(cells, normals, bc["tangential"])
        elif bc["kind"] == "dirichlet_bb":
            solver.add_dirichlet_bb(cells, normals, bc["value"])
        else:
            raise NotImplementedError(bc["kind"])
    return solver


# per-moment MRT rates used by the tests: every non-conserved moment relaxes at its own rate, so that a mix-up shows
def mrt_rates(ndist, omega):
    r = np.full(27, omega)
    r[:ndist] = omega * (1.0 + 0.01 * np.arange(ndist) / ndist)
    return r
