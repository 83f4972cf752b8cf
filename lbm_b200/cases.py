"""Configurations -> solver set-up calls (product side; used by bench.py, the partitioned workloads and the tests).

`bcs_from_config` mirrors what LBMSolver::loadConfiguration / LBMBndManager::setupBndryCnds do in the reference
(src/lbm/solver.cpp:71-144, src/lbm/bnd/bnd.h:71-142): boundary conditions are created per geometry and per surface key in
byte-lexicographic order (nlohmann::json objects are std::map), surfaces without cells are skipped, `generateBndry:false` produces
a dummy.  The C++ host (lbm_b200/host/lbm_solver.hpp) does the same for the `lbm` executable; this is the Python twin for callers
that drive the C ABI through lbm_b200.Solver.

`sphere3d_config` / `step3d_config` are BASELINE.json configs[3] / configs[4]: the reference's test/sphere/sphere_ns.json and
test/step/step_ns.json (2D, D2Q9) with their geometry objects, boundary keys and pressure values kept and the third dimension added
(walls on -z / +z).  The reference's executable rejects D3Q19 / D3Q27 (src/lbm/solverExe.h:37-90), so these are extensions.
"""
import numpy as np

WALL = {"type": "wall", "model": "bounceback"}
NDIST = {"D2Q9": 9, "D3Q19": 19, "D3Q27": 27}


def omega_from_config(solver_cfg, maxlvl):
    """src/lbm/solver.cpp:102-123"""
    if "relaxation" in solver_cfg:
        return 1.0 / float(solver_cfg["relaxation"])
    ma = float(solver_cfg["ma"])
    re = float(solver_cfg["reynoldsnumber"])
    ref_length = float(solver_cfg.get("refLength", 1.0))
    nu = ma / re * ref_length
    return 2.0 / (1.0 + 2.0 * nu * 2.0 ** maxlvl)



def poisson_parameters(solver_cfg, ncells, ndim):
    """m_dt and poisson_D of the Poisson equation types (src/lbm/solver.cpp:100,128-137,589-599): the time step follows from the cell
    count (m_finestGridSpacing = 1 / (size^(1/NDIM) - 1)), the reaction rate is `equation_th` for "simple_diff_reaction" and the
    Debye-Hueckel constant 27.79 otherwise."""
    spacing = 1.0 / (float(ncells) ** (1.0 / ndim) - 1)
    lbm_cs = 1.0 / np.sqrt(3.0)
    ma = 1.0 / lbm_cs
    dt = spacing * ma * lbm_cs / float(solver_cfg.get("refLength", 1.0))
    if solver_cfg.get("equation_application", "debye_huckel") == "simple_diff_reaction":
        rate = float(solver_cfg["equation_th"])
    else:
        rate = 27.79
    return dt, rate


def bcs_from_config(solver_cfg, surfaces, ndim, expr_values=None):
    """surfaces: name -> (cells, normals).  Returns (bc list in application order, forcing or None).
    expr_values: name -> per-entry values for boundary conditions whose "value" is a math expression (the reference evaluates it
    with exprtk at the cell centres, bnd_dirichlet.h:268-281); the caller supplies the evaluated numbers."""
    bcs = []
    boundary = solver_cfg["boundary"]
    for geom in sorted(boundary):
        keys = boundary[geom]
        for key in sorted(keys):
            conf = keys[key]
            sname = f"{geom}_{key}" if len(keys) > 1 else geom
            cells, normals = surfaces.get(sname, (np.zeros(0, np.int64), np.zeros((0, ndim))))
            if len(cells) == 0:
                continue  # bnd.h:83-86
            if not conf.get("generateBndry", True):
                continue  # LBMBnd_dummy, bnd.h:149-159
            t = conf["type"]
            if t == "periodic":
                conn = surfaces[conf["connection"]][0]
                bcs.append(dict(kind="periodic", cells=cells, normals=normals, connected=conn,
                                pressure=float(conf.get("pressure", "nan"))))
            elif t == "wall":
                if conf["model"] == "bounceback":
                    bcs.append(dict(kind="wall_bb", cells=cells, normals=normals,
                                    tangential=float(conf.get("tangentialVelocity", 0.0))))
                elif conf["model"] in ("equilibrium", "neem", "nebb"):
                    vel = np.array(conf["velocity"], float)[:ndim] if "velocity" in conf else None
                    bcs.append(dict(kind="wall_wetnode", model=conf["model"], cells=cells, normals=normals, velocity=vel))
                else:
                    raise ValueError(f"Invalid wall boundary model: {conf['model']}")
            elif t == "pressure":
                bcs.append(dict(kind="pressure", cells=cells, normals=normals, pressure=float(conf["pressure"])))
            elif t in ("dirichlet", "neumann") and conf["model"] == "neem":  # Poisson equation types, bnd.h:116-137
                val = conf["value"]
                if isinstance(val, str):
                    if expr_values is None or sname not in expr_values:
                        raise ValueError(f"boundary {sname}: the expression {val!r} must be evaluated by the caller (expr_values)")
                    values = np.asarray(expr_values[sname], dtype=np.float64)
                else:
                    values = np.full(len(cells), float(val))
                bcs.append(dict(kind="poisson_neem", neumann=t == "neumann", cells=cells, normals=normals, values=values, grad=0.0))
            elif t == "dirichlet" and conf["model"] == "bounceback":
                bcs.append(dict(kind="dirichlet_bb", cells=cells, normals=normals,
                                value=np.array(conf["value"], float)[:ndim]))
            else:
                raise NotImplementedError(f"boundary type {t}")
    forcing = None
    if solver_cfg.get("forcing", ""):
        forcing = dict(inlet=surfaces["cube_-x"][0], outlet=surfaces["cube_+x"][0],
                       gradient=float(solver_cfg["poiseuillePressureGradient"]))
    return bcs, forcing



def sphere3d_config(level, model="D3Q27"):
    """sphere_ns.json (box [0,10]^2 minus a sphere of radius 1 at the centre, pressure in-/outlet on -x/+x) in 3D"""
    return {"dim": 3, "partitionLevel": level, "uniformLevel": level, "maxRfnmtLvl": level, "maxNoCells": 100000000,
            "outputDir": "out", "gridFileName": "gridD",
            "geometry": {"cube": {"type": "box", "body": "flowregion", "A": [0.0, 0.0, 0.0], "B": [10.0, 10.0, 10.0]},
                         "sphere": {"type": "sphere", "body": "flowregion", "subtract": True, "center": [5.0, 5.0, 5.0], "radius": 1.0}},
            "solver": {"type": "lbm", "model": model, "relaxation": 0.6, "ma": 0.01, "maxSteps": 10,
                       "boundary": {"cube": {"+x": {"type": "pressure", "pressure": 1.0}, "-x": {"type": "pressure", "pressure": 1.0000008},
                                             "-y": WALL, "+y": WALL, "-z": WALL, "+z": WALL},
                                    "sphere": {"all": WALL}}}}



def step3d_config(level, model="D3Q19"):
    """step_ns.json (channel [0,10]x[0,9] with two side pockets, i.e. a block on the upper wall, pressure in-/outlet) extruded in z"""
    return {"dim": 3, "partitionLevel": level, "uniformLevel": level, "maxRfnmtLvl": level, "maxNoCells": 100000000,
            "outputDir": "out", "gridFileName": "gridD",
            "geometry": {"cube": {"type": "box", "body": "flowregion", "A": [0.0, 0.0, 0.0], "B": [10.0, 9.0, 10.0]},
                         "step_a": {"type": "box", "body": "flowregion", "subtract": False, "A": [0.0, 9.0, 0.0], "B": [4.0, 10.0, 10.0]},
                         "step_b": {"type": "box", "body": "flowregion", "subtract": False, "A": [6.0, 9.0, 0.0], "B": [10.0, 10.0, 10.0]}},
            "solver": {"type": "lbm", "model": model, "relaxation": 0.6, "ma": 0.01, "maxSteps": 10,
                       "boundary": {"cube": {"+x": {"type": "pressure", "pressure": 1.0}, "-x": {"type": "pressure", "pressure": 1.0000008},
                                             "-y": WALL, "+y": WALL, "-z": WALL, "+z": WALL},
                                    "step_a": {"+x": WALL, "-x": {"type": "pressure", "pressure": 1.0000008}, "+y": WALL, "-y": WALL,
                                               "-z": WALL, "+z": WALL},
                                    "step_b": {"+x": {"type": "pressure", "pressure": 1.0}, "-x": WALL, "+y": WALL, "-y": WALL,
                                               "-z": WALL, "+z": WALL}}}}



CONFIGS = {"sphere3d": sphere3d_config, "step3d": step3d_config}


def apply_bcs(solver, bcs, forcing=None):
    """Issue the boundary-condition calls on `solver` (lbm_b200.Solver, or the test oracle: same method names)."""
    for bc in bcs:
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
    if forcing is not None:
        solver.set_forcing(forcing["inlet"], forcing["outlet"], forcing["gradient"])
    return solver


def restrict_bcs(bcs, lp):
    """The entries of the cells rank `lp.rank` of a partitioned run owns, local ids, order kept (lbm_b200/partition.py)."""
    out = []
    for bc in bcs:
        if bc["kind"] not in ("wall_bb", "dirichlet_bb", "pressure"):
            raise NotImplementedError(f"boundary condition {bc['kind']} is not partitioned")
        cells, normals = lp.restrict(bc["cells"], bc["normals"])
        if len(cells):
            out.append(dict(bc, cells=cells, normals=normals))
    return out


def trt_omega_minus(omega, magic=3.0 / 16.0):
    """odd-moment rate of the two-relaxation-time operator from the 'magic' parameter (1/w+ - 1/2)(1/w- - 1/2) = magic"""
    return 1.0 / (magic / (1.0 / omega - 0.5) + 0.5)
