"""Turn a reference configuration + dumped tables into solver set-up calls (test helper).

Mirrors what LBMSolver::loadConfiguration / LBMBndManager::setupBndryCnds do in the reference
(src/lbm/solver.cpp:71-144, src/lbm/bnd/bnd.h:71-142): boundary conditions are created per geometry and per
surface key in byte-lexicographic order (nlohmann::json objects are std::map), surfaces without cells are
skipped, `generateBndry:false` produces a dummy.  The same spec drives the CPU oracle and the CUDA product,
so both see identical inputs.
"""
import json
import os
from dataclasses import dataclass, field

import numpy as np

from lbm_b200.cases import bcs_from_config, omega_from_config  # noqa: F401  (re-exported: the tests use the product's own mapping)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@dataclass
class CaseSpec:
    name: str
    ndim: int
    ndist: int
    nghbr: np.ndarray            # [n, stride] int64 push table, -1 = none
    omega: float
    center: np.ndarray = None    # [n, ndim]
    bbmin: np.ndarray = None
    bbmax: np.ndarray = None
    cell_length: float = 0.0
    bcs: list = field(default_factory=list)
    forcing: dict = None
    golden: dict = None
    equation: str = "navierstokes"
    poisson: dict = None         # dt, rate (lbm_b200.cases.poisson_parameters)

    @property
    def n(self):
        return self.nghbr.shape[0]

    def apply_to(self, solver):
        """Issue the set-up calls on `solver` (oracle.Oracle or lbm_b200.Solver: same method names)."""
        if self.center is not None:
            solver.set_geometry(self.center, self.bbmin, self.bbmax, self.cell_length)
        if self.equation == "poisson":
            solver.set_poisson(self.poisson["dt"], self.poisson["rate"])
        for bc in self.bcs:
            k = bc["kind"]
            if k == "wall_bb":
                solver.add_wall_bb(bc["cells"], bc["normals"], bc["tangential"])
            elif k == "dirichlet_bb":
                solver.add_dirichlet_bb(bc["cells"], bc["normals"], bc["value"])
            elif k == "pressure":
                solver.add_pressure(bc["cells"], bc["normals"], bc["pressure"])
            elif k == "periodic":
                solver.add_periodic(bc["cells"], bc["normals"], bc["connected"], bc["pressure"])
            elif k == "wall_wetnode":
                solver.add_wall_wetnode(bc["model"], bc["cells"], bc["normals"], bc["velocity"])
            elif k == "poisson_neem":
                solver.add_poisson_neem("neumann" if bc["neumann"] else "dirichlet", bc["cells"], bc["normals"], bc["values"], bc["grad"])
            else:
                raise ValueError(k)
        if self.forcing is not None:
            solver.set_forcing(self.forcing["inlet"], self.forcing["outlet"], self.forcing["gradient"])
        return solver


def geometry_bbox(geometry_cfg, ndim):
    """GeometryManager::getBoundingBox (src/geometry.h) for analytic objects."""
    lo = np.full(ndim, np.inf)
    hi = np.full(ndim, -np.inf)
    for _, g in sorted(geometry_cfg.items()):
        if g["type"] == "box":
            a, b = np.array(g["A"], float), np.array(g["B"], float)
        elif g["type"] == "sphere":
            c = np.array(g["center"], float)
            a, b = c - g["radius"], c + g["radius"]
        elif g["type"] == "cube":
            c = np.array(g["center"], float)
            r = np.sqrt(ndim) * g["length"]
            a, b = c - r, c + r
        else:
            raise ValueError(g["type"])
        lo, hi = np.minimum(lo, a), np.maximum(hi, b)
    return lo, hi


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False))
    cfg = json.loads(str(g["config_json"]))
    ndim, ndist = int(g["ndim"]), int(g["ndist"])
    surfaces = {}
    for k, sname in enumerate(g["surface_names"]):
        surfaces[str(sname)] = (g[f"surf{k}_cells"].astype(np.int64), g[f"surf{k}_normals"])
    lo, hi = geometry_bbox(cfg["geometry"], ndim)
    l0 = float(np.max(hi - lo))
    spec = CaseSpec(name=name, ndim=ndim, ndist=ndist, nghbr=g["nghbr"].astype(np.int64), omega=float(g["omega"]),
                    center=g["center"], bbmin=lo, bbmax=hi, cell_length=float(g["cell_length"]), golden=g)
    expr_values = {str(s): g[f"surf{k}_values"] for k, s in enumerate(g["surface_names"]) if f"surf{k}_values" in g}
    spec.bcs, spec.forcing = bcs_from_config(cfg["solver"], surfaces, ndim, expr_values)
    if cfg["solver"].get("equation", "navierstokes") == "poisson":
        from lbm_b200.cases import poisson_parameters
        spec.equation = "poisson"
        dt, rate = poisson_parameters(cfg["solver"], spec.n, ndim)
        spec.poisson = dict(dt=dt, rate=rate)
    spec.config = cfg
    spec.surfaces = surfaces
    spec.digests = json.loads(str(g["digests_json"]))
    return spec
