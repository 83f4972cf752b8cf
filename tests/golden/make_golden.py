#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the reference binary.

Runs oracle/_ref/lbm_ref (the reference's own solver built by oracle/ref_build/Makefile, arithmetic
untouched) on the reference's own test configurations (/root/reference/test/...), with only
`solver.maxSteps` shortened, and stores what the dump hook wrote:

  * tables : neighbour table (cartesiangrid.h:111-124), property bits, cell centres, boundary surfaces
             (cell list + per-entry normal) in boundary-condition application order
  * state  : raw m_fold, m_f, m_vars, m_varsold right after timeStep() number s for the listed steps

Small cases keep the full arrays; the sphere case (63 576 cells) keeps the tables, `vars`, and SHA-256
digests of the population arrays (the C oracle is bit-exact, so a digest pins it).

Needs /root/reference and oracle/_ref/lbm_ref, i.e. it only runs in the build container.  The fixtures it
writes are committed, so tests never need the reference at run time.

usage: python tests/golden/make_golden.py [case ...]
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref", "lbm_ref")

# name -> (config relative to REF/test, maxSteps, dump steps, keep full populations?)
CASES = {
    "couette": ("couette/couette.json", 200, [1, 2, 10, 100, 200], True),
    "couette_bnd": ("couette/couette_bnd.json", 100, [1, 10, 100], True),
    "couette_bnd_bbDirichlet": ("couette/couette_bnd_bbDirichlet.json", 100, [1, 10, 100], True),
    "poiseuille": ("poiseuille/poiseuille.json", 200, [1, 2, 10, 200], True),
    "poiseuille_bnd": ("poiseuille/poiseuille_bnd.json", 100, [1, 10, 100], True),
    "step_ns": ("step/step_ns.json", 200, [1, 10, 200], True),
    "sphere_ns": ("sphere/sphere_ns.json", 300, [1, 300], False),
    # wet-node wall family and the remaining run.sh Navier-Stokes cases
    "couette_bnd_eq": ("couette/couette_bnd_eq.json", 100, [1, 2, 10, 100], True),
    "couette_bnd_eq2": ("couette/couette_bnd_eq2.json", 100, [1, 10, 100], True),
    "couette_bnd_eq_aligned": ("couette/couette_bnd_eq_aligned.json", 100, [1, 10, 100], True),
    "couette_bnd_NEEM": ("couette/couette_bnd_NEEM.json", 100, [1, 2, 10, 100], True),
    "couette_bnd_NEBB": ("couette/couette_bnd_NEBB.json", 100, [1, 2, 10, 100], True),
    "poiseuille_bnd_eq": ("poiseuille/poiseuille_bnd_eq.json", 100, [1, 10, 100], True),
    "poiseuille_bnd_NEEM": ("poiseuille/poiseuille_bnd_NEEM.json", 100, [1, 10, 100], True),
    "poiseuille_bnd_NEBB": ("poiseuille/poiseuille_bnd_NEBB.json", 100, [1, 10, 100], True),
    "poiseuille_bnd_pressure": ("poiseuille/poiseuille_bnd_pressure.json", 100, [1, 2, 10, 100], True),
    "poiseuille_bnd_pressure_neem2": ("poiseuille/poiseuille_bnd_pressure_neem2.json", 100, [1, 10, 100], True),
    # multi-level grids: the reference's configurations with only the three level keys changed.  The
    # reference keeps the cells of ALL levels >= partitionLevel in the list and steps every level as its own lattice
    # ("todo: skip non-leaf cells", solver.cpp:525); boundary refinement (maxRfnmtLvl > uniformLevel) adds children of cut cells.
    "couette_ml_p3u5": ("couette/couette.json", 50, [1, 10, 50], True, {"partitionLevel": 3, "uniformLevel": 5, "maxRfnmtLvl": 5}),
    "couette_ml_u5m6": ("couette/couette.json", 50, [1, 10, 50], True, {"partitionLevel": 5, "uniformLevel": 5, "maxRfnmtLvl": 6}),
    "couette_ml_p4u5m7": ("couette/couette.json", 50, [1, 10, 50], True, {"partitionLevel": 4, "uniformLevel": 5, "maxRfnmtLvl": 7, "_threads": 1}),
    # (boundary refinement next to a pressure surface is not a valid reference configuration: the refined layer is two cells wide, so
    # the second inward neighbour of LBMBnd_Pressure does not exist and the reference reads m_vars[-1], bnd_pressure.h:68-84)
    "sphere_ml_p4u6": ("sphere/sphere_ns.json", 50, [1, 50], True, {"partitionLevel": 4, "uniformLevel": 6, "maxRfnmtLvl": 6}),
    # Poisson equation types: the five Poisson cases of the reference's test/run.sh.  D1Q3 / D2Q5 / D2Q9, one
    # variable (the potential), Dirichlet / Neumann NEEM boundary conditions; poisson2D_helmholtz has expression-valued boundaries
    # (exprtk in the reference): the evaluated per-entry numbers are read back from the reference's own m_vars after step 1
    "poisson1D": ("poisson/poisson1D.json", 100, [1, 2, 10, 100], True),
    "poisson1D_reaction": ("poisson/poisson1D_reaction.json", 100, [1, 2, 10, 100], True),
    "poisson2D": ("poisson/poisson2D.json", 100, [1, 10, 100], False),
    "poisson2D_helmholtz": ("poisson/poisson2D_helmholtz.json", 50, [1, 50], False),
    "poissonD2Q9": ("poisson/poissonD2Q9.json", 100, [1, 10, 100], False),
    # two more Poisson configurations of the reference's test directory (not part of run.sh): Neumann + Dirichlet in 2D, and the step
    # geometry with twelve NEEM surfaces (concave corners, cells on several surfaces)
    "poisson2D_reaction": ("poisson/poisson2D_reaction.json", 100, [1, 10, 100], True),
    "step_poisson": ("step/step_poisson.json", 100, [1, 10, 100], True),
    "step_ml_p3u5": ("step/step_ns.json", 50, [1, 50], True, {"partitionLevel": 3, "uniformLevel": 5, "maxRfnmtLvl": 5}),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def parse_surfaces(path, ndim):
    surfaces = []
    cur = None
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "surface":
            cur = {"name": t[1], "cells": [], "normals": [], "declared": int(t[2])}
            surfaces.append(cur)
        else:
            cur["cells"].append(int(t[0]))
            cur["normals"].append([float(x) for x in t[1:1 + ndim]])
    return surfaces


def run_case(name):
    rel, max_steps, steps, full = CASES[name][:4]
    cfg = json.load(open(os.path.join(REF, "test", rel)))
    variant = dict(CASES[name][4]) if len(CASES[name]) > 4 else {}
    # "_threads": with three or more levels the grid-level periodic pairing (cartesiangrid.h:608-706, "centres differ in exactly one
    # coordinate") links several cells of one level to the same neighbour, so two cells push into one slot and the reference's
    # `#pragma omp parallel for` propagation (solver.cpp:728) races; with one thread the highest source wins deterministically
    threads = str(variant.pop("_threads", 2))
    cfg.update(variant)  # variants: top-level keys of the grid generator only
    original = json.dumps(cfg)
    cfg["solver"]["maxSteps"] = max_steps
    # keep the run quiet and free of early termination; none of these keys touches the arithmetic
    cfg["solver"]["solution_interval"] = 10 ** 9
    cfg["solver"]["convergence"] = 0.0
    cfg["solver"].pop("analyticalSolution", None)
    cfg["solver"].pop("postprocessing", None)
    tmp = tempfile.mkdtemp(prefix="lbm_golden_")
    try:
        os.makedirs(os.path.join(tmp, "dump"))
        json.dump(cfg, open(os.path.join(tmp, "case.json"), "w"), indent=1)
        env = dict(os.environ, SFCMM_DUMP="dump", SFCMM_DUMP_STEPS=",".join(map(str, steps)), OMP_NUM_THREADS=threads)
        r = subprocess.run([BIN, "case.json"], cwd=tmp, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{name}: reference exited {r.returncode}\n{r.stderr[-2000:]}")
        d = os.path.join(tmp, "dump")
        meta = dict(l.split() for l in open(os.path.join(d, "meta.txt")))
        n, ndim, q, nvar, nn = (int(meta[k]) for k in ("ncells", "ndim", "ndist", "nvar", "nnghbr"))
        out = {
            "config_json": np.array(json.dumps(cfg)),
            "config_orig_json": np.array(original),  # the reference's test configuration, unmodified
            "ncells": n, "ndim": ndim, "ndist": q, "nvar": nvar, "nnghbr": nn,
            "omega": float(meta["omega"]), "nu": float(meta["nu"]), "maxlvl": int(meta["maxlvl"]),
            "cell_length": float(meta["celllength"]),  # grid().lengthOnLvl(maxLvl()) of the solver grid (after alignment)
            "nghbr": np.fromfile(os.path.join(d, "nghbr.i64"), dtype=np.int64).reshape(n, nn).astype(np.int32),
            "props": np.fromfile(os.path.join(d, "props.u64"), dtype=np.uint64).astype(np.uint16),
            "center": np.fromfile(os.path.join(d, "center.f64"), dtype=np.float64).reshape(n, ndim),
            "steps": np.array(steps, dtype=np.int64),
        }
        surfaces = parse_surfaces(os.path.join(d, "surfaces.txt"), ndim)
        out["surface_names"] = np.array([s["name"] for s in surfaces])
        for k, s in enumerate(surfaces):
            out[f"surf{k}_cells"] = np.array(s["cells"], dtype=np.int32)
            out[f"surf{k}_normals"] = np.array(s["normals"], dtype=np.float64).reshape(len(s["cells"]), ndim)
        if cfg["solver"].get("equation", "navierstokes") == "poisson":
            # per-entry boundary values: the constant of the configuration, or -- for math expressions -- what the reference's exprtk
            # evaluation produced, which the Dirichlet condition writes into m_vars of its cells in every apply (bnd_dirichlet.h:352)
            v1 = np.fromfile(os.path.join(d, f"vars_{steps[0]}.f64"), dtype=np.float64).reshape(n, nvar)[:, 0]
            bnd = cfg["solver"]["boundary"]
            confs = {(f"{gk}_{sk}" if len(bnd[gk]) > 1 else gk): bnd[gk][sk] for gk in bnd for sk in bnd[gk]}
            for k, s in enumerate(surfaces):
                val = confs[s["name"]].get("value", 0)
                cells = np.array(s["cells"], dtype=np.int64)
                out[f"surf{k}_values"] = v1[cells].copy() if isinstance(val, str) else np.full(len(cells), float(val))
        digests = {}
        for s in steps:
            for arr, width in (("fold", q), ("f", q), ("vars", nvar), ("varsold", nvar)):
                a = np.fromfile(os.path.join(d, f"{arr}_{s}.f64"), dtype=np.float64).reshape(n, width)
                digests[f"{arr}_{s}"] = sha(a)
                if full or arr == "vars":
                    out[f"{arr}_{s}"] = a
        out["digests_json"] = np.array(json.dumps(digests))
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        sz = os.path.getsize(os.path.join(HERE, f"{name}.npz"))
        print(f"{name}: {n} cells, {len(surfaces)} surfaces, steps {steps} -> {sz / 1024:.0f} KiB")
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
