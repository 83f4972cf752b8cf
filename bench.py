#!/usr/bin/env python3
"""bench.py -- MLUPS of the lattice-Boltzmann time step on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], SURVEY.md section 8d "S3 bench-3D"): cube of S^3 cells, D3Q19, BGK, fp64,
omega = 1/0.6, periodic in x (grid level), no-slip bounce-back walls on -y/+y/-z, moving lid on +z
(Dirichlet bounce-back, u = (0.05, 0, 0)), reference initial condition (rho = 1, u = 0 except the lid preset).
A "step" is one LBM time step of the whole box.  Default S = 256 (16.8 M cells, 5.1 GB of populations: far
larger than the 126 MB L2, so no L2 flush is needed between timed steps).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events on the solver's stream, max over
ranks); `e2e` is the same metric through the C ABI with the state in pinned HOST buffers: upload m_f/m_fold,
K steps, download the macroscopic fields -- all inside the timed region.  `--impl reference` times the CPU
restatement of the reference's own time step (oracle/lbm_oracle.c: the reference's algorithm, pass by pass,
with its OpenMP pragmas) on the host cores; the reference binary itself cannot run D3Q19.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# torchrun exports OMP_NUM_THREADS=1; the host-side set-up (table generation, device plan) is OpenMP code, so give every
# rank its share of the host cores before any OpenMP runtime is loaded
# the reference arm runs on rank 0 alone (the other ranks exit at once): it gets ALL host cores whatever the launcher's world size is
if "reference" in sys.argv and int(os.environ.get("RANK", "0")) == 0:
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
elif os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))

# stdout carries exactly ONE JSON line: everything libraries print (NCCL banner, torchrun notes) goes to stderr
_JSON_FD = 1
if __name__ == "__main__":
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)

import numpy as np


def emit(line):
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LATTICES = {"D2Q9": (2, 9), "D3Q19": (3, 19), "D3Q27": (3, 27)}
OMEGA = 1.0 / 0.6
LID_U = 0.05


def global_shape(size, ndim, world):
    """weak scaling: every rank owns one size^ndim cube; the global box doubles along x, then y, then z"""
    mult = [1] * ndim
    w, d = world, 0
    while w > 1:
        if w % 2:
            raise SystemExit("--gpus must be a power of two")
        mult[d % ndim] *= 2
        w //= 2
        d += 1
    return tuple(size * m for m in mult)


def box_table_numpy(shape, periodic, ndist):
    """The benchmark box's push table in the reference's format, in plain numpy: cells in ascending order of the reference's curve key
    (include/common/math/hilbert.h:16-48), column i = the cell one step along direction i of LBMethod<>::m_dirs (src/lbm/constants.h),
    -1 outside.  Used by the reference arm so that it never loads the CUDA library; equal to lbm_b200_box_topology
    (tests/test_bench_contract.py)."""
    shape = tuple(int(v) for v in shape)
    ndim = len(shape)
    level = max(int(np.ceil(np.log2(max(shape)))), 1)
    lut = np.array([0, 3, 1, 2, 5, 4, 6, 7], dtype=np.int64)
    d2 = [[-1, 0], [1, 0], [0, -1], [0, 1], [1, 1], [1, -1], [-1, -1], [-1, 1]]
    d3 = [[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1], [-1, -1, 0], [-1, 1, 0], [1, -1, 0], [1, 1, 0],
          [-1, 0, -1], [-1, 0, 1], [1, 0, -1], [1, 0, 1], [0, -1, -1], [0, -1, 1], [0, 1, -1], [0, 1, 1], [-1, -1, -1], [-1, -1, 1],
          [-1, 1, -1], [-1, 1, 1], [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]]
    coords = [g.ravel() for g in np.meshgrid(*[np.arange(v, dtype=np.int32) for v in shape], indexing="ij")]
    key = np.zeros(coords[0].shape[0], dtype=np.int64)
    for l in range(level):
        q = np.zeros(key.shape[0], dtype=np.int64)
        for d in range(ndim):
            q |= ((coords[d] >> (level - 1 - l)) & 1).astype(np.int64) << d
        key = (key << ndim) | lut[q]
    order = np.argsort(key, kind="stable")
    del key
    coords = [c[order] for c in coords]
    n = order.shape[0]
    strides = [int(np.prod(shape[d + 1:])) for d in range(ndim)]
    lin2cell = np.empty(n, dtype=np.int64)
    lin2cell[order] = np.arange(n)
    nghbr = np.full((n, ndist - 1), -1, dtype=np.int64)
    for i, dv in enumerate((d2 if ndim == 2 else d3)[:ndist - 1]):
        lin = np.zeros(n, dtype=np.int64)
        ok = np.ones(n, dtype=bool)
        for d in range(ndim):
            c = coords[d].astype(np.int64) + dv[d]
            if periodic[d]:
                c %= shape[d]
            else:
                ok &= (c >= 0) & (c < shape[d])
            lin += np.clip(c, 0, shape[d] - 1) * strides[d]
        col = lin2cell[lin]
        col[~ok] = -1
        nghbr[:, i] = col
    return nghbr


def workload(size, lattice, rank=0, world=1, native=True, shape=None):
    """Tables of the benchmark box in the reference's format + boundary conditions in application order.
    world > 1: this rank's contiguous range of the SFC-ordered list plus ghost cells (lbm_b200/partition.py).
    native = False: tables from plain numpy (the reference arm must not touch the CUDA library)."""
    ndim, ndist = LATTICES[lattice]
    periodic = (1,) + (0,) * (ndim - 1)
    lp = None
    if world == 1:
        shape = (size,) * ndim if shape is None else tuple(shape)
        center = None
        if native:
            from lbm_b200.capi import box_topology
            nghbr, center, _ = box_topology(shape, periodic, want_center=True)
        else:
            nghbr = box_table_numpy(shape, periodic, ndist)
        n_owned = nghbr.shape[0]
    else:
        from lbm_b200 import partition
        shape = global_shape(size, ndim, world) if shape is None else tuple(shape)
        from lbm_b200.capi import NativePartition
        # the library's own partition code (lbm_b200_partition_*, csrc/partition.hpp) over on-demand box rows
        lp = NativePartition(partition.BoxRows(shape, periodic, ndist), ndim, ndist, 8 if ndim == 2 else 26, rank, world)
        nghbr, center, n_owned = lp.nghbr, None, lp.n_owned
    names = ["-x", "+x", "-y", "+y", "-z", "+z"][:2 * ndim]
    lid = names[-1]
    bcs = []
    for d, nm in sorted(enumerate(names), key=lambda t: t[1]):  # lexicographic, like the reference (bnd.h:71-142)
        cells = np.nonzero(nghbr[:n_owned, d] < 0)[0].astype(np.int64)
        if len(cells) == 0:
            continue
        normal = np.zeros(ndim)
        normal[d // 2] = -1.0 if d % 2 == 0 else 1.0
        normals = np.tile(normal, (len(cells), 1))
        if nm == lid:
            value = np.zeros(ndim)
            value[0] = LID_U
            bcs.append(("dirichlet_bb", cells, normals, value))
        else:
            bcs.append(("wall_bb", cells, normals, 0.0))
    return dict(ndim=ndim, ndist=ndist, nghbr=nghbr, center=center, bcs=bcs, shape=shape, lp=lp, n_owned=n_owned)


def apply_bcs(solver, wl):
    for kind, cells, normals, val in wl["bcs"]:
        if kind == "dirichlet_bb":
            solver.add_dirichlet_bb(cells, normals, val)
        elif kind == "pressure":
            solver.add_pressure(cells, normals, val)
        else:
            solver.add_wall_bb(cells, normals, val)
    return solver


def case_workload(name, size, rank=0, world=1):
    """BASELINE.json configs[3] ("sphere": flow past a sphere, D3Q27, MRT) and configs[4] ("step": channel with a step, D3Q19, TRT,
    pressure outflow) as 3D cases on a cube of size^3 cells (lbm_b200/cases.py), cut into `world` contiguous SFC ranges.  The grid
    comes from the host pipeline's on-demand row provider (lbm_b200/host/uniform_grid.hpp), so no rank ever holds the whole table."""
    import math
    from lbm_b200 import cases, host_api, partition
    level = int(round(math.log2(size)))
    if 2 ** level != size:
        raise SystemExit("--size must be a power of two for the sphere / step workloads (cells per side of the cube)")
    cfg = cases.CONFIGS[name + "3d"](level)
    ndim, ndist = 3, cases.NDIST[cfg["solver"]["model"]]
    tmp = tempfile.mkdtemp(prefix="lbm_case_")
    path = os.path.join(tmp, "case.json")
    with open(path, "w") as fh:
        json.dump(cfg, fh)
    ug = host_api.UniformGrid(path)
    surfaces = {nm: (cells, normals) for nm, cells, normals in ug.surfaces()}
    bcs, _ = cases.bcs_from_config(cfg["solver"], surfaces, ndim)
    lp = None
    if world == 1:
        nghbr = ug.rows(np.arange(ug.n, dtype=np.int64))[0]
        n_owned = ug.n
    else:
        pressure = [(bc["cells"], bc["normals"]) for bc in bcs if bc["kind"] == "pressure"]
        lp = partition.plan_rank(partition.GridRows(ug, ndist), rank, world, ug.stride, pressure)
        bcs = cases.restrict_bcs(bcs, lp)
        nghbr, n_owned = lp.nghbr, lp.n_owned
    flat = [(bc["kind"], bc["cells"], bc["normals"], bc.get("pressure", bc.get("tangential", 0.0))) for bc in bcs]
    n_global = int(ug.n)
    ug.close()
    return dict(ndim=ndim, ndist=ndist, nghbr=nghbr, center=None, bcs=flat, shape=(size,) * 3, lp=lp, n_owned=n_owned, n_global=n_global,
                lattice=cfg["solver"]["model"], collision={"sphere": "mrt", "step": "trt"}[name])


class ClockSampler:
    """SM clock / power / throttle reasons of the device DURING the measurement, sampled in-process through NVML every few milliseconds
    by a thread (the timed region of the default run lasts ~20 ms, too short for an `nvidia-smi -lms` loop; ctypes releases the GIL
    while the C ABI call runs).  Covers warm-up and the timed steps."""

    def __init__(self, device):
        import threading
        self.samples, self.reasons, self.ok, self._stop = [], set(), False, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            import torch
            # NVML enumerates physical devices: map through the UUID so that CUDA_VISIBLE_DEVICES does not shift the index
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            self.h = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.nv = pynvml
            self.ok = True
        except Exception as e:  # noqa: BLE001 -- any NVML problem just means "no clock record from here"
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, pw, util))
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.004)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["NVML unavailable: " + getattr(self, "err", "?")]}
        self._stop.set()
        self.t.join(timeout=2)
        try:
            smax = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            smax = None
        sm = [x[0] for x in self.samples]
        busy = [x[0] for x in self.samples if x[1] > 250.0]   # samples with the device visibly under load (power above idle)
        return {"sm_mhz": float(np.median(busy if busy else sm)) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max((x[1] for x in self.samples), default=None), "samples": len(sm), "samples_under_load": len(busy),
                "reasons": sorted(self.reasons), "how": "NVML in-process (one query set takes several ms) from the first warm-up step to the end of the timed steps, incl. the residual-mode run"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def kernel_source_hash():
    """SHA-256 (first 16 hex digits) of the device code the main kernels are built from: an ncu capture is only quoted for the binary
    it was taken on"""
    import hashlib
    h = hashlib.sha256()
    for name in ("kernels.cuh", "lattice.h", "mrt_tables.h"):
        h.update(open(os.path.join(ROOT, "lbm_b200", "csrc", name), "rb").read())
    return h.hexdigest()[:16]


def output_path_sample(s, reps=3):
    """SURVEY.md section 8f N2 beside the step: the fields of one solution file (all cells kept) through lbm_b200_encode_output -- filter
    gather, 15-decimal rounding and base64 on the device, the text into page-locked memory -- and, for scale, lbm_b200_get_moments into
    pageable host memory, which is where the host writer's route only starts (tools/bench_output.py times that route completely)"""
    import lbm_b200
    n = s.n
    total = s.nvar * int(s._lib.lbm_b200_output_chars(n))
    mem = lbm_b200.HostBuffer(total)
    try:
        enc, mom = [], []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            s.encode_output(None, out=mem.array, raw=True)
            enc.append(time.perf_counter() - t0)
        for _ in range(reps):
            t0 = time.perf_counter()
            s.moments()
            mom.append(time.perf_counter() - t0)
    finally:
        mem.close()
    enc_ms, mom_ms = float(np.median(enc[1:])) * 1e3, float(np.median(mom)) * 1e3      # the first call builds the selection list
    return {"what": "fields of one solution file, all cells kept: lbm_b200_encode_output into page-locked memory (C ABI, host buffer out)",
            "cells": int(n), "fields": int(s.nvar), "text_bytes": int(total), "device_encode_ms": enc_ms,
            "Mcells_per_s": n / enc_ms / 1e3, "moments_to_pageable_host_ms": mom_ms,
            "host_route": "profiles/r02_bench_output_path.json: moments to host + rounding + base64 on 16 cores = 480 ms at 256^3"}


def ncu_traffic(lattice, size):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the chunk kernel from the committed ncu capture
    (profiles/traffic.json), or None when the capture is of other device code than the one in the tree"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        t = json.load(open(p))
        key = f"{lattice}_{size}"
        if t.get("source_sha16") == kernel_source_hash() and key in t:
            return t[key]
    return None


def mrt_kinds(lattice):
    """kind of every moment of the MRT basis (0 conserved, 1 shear, 2 bulk, 3 ghost), read from the generated table header so that the
    reference arm does not have to load the CUDA library for it"""
    import re
    text = open(os.path.join(ROOT, "oracle", "mrt_tables.h")).read()
    m = re.search(r"#define LBM_MRT_%s_KIND \{([^}]*)\}" % lattice, text)
    return np.array([int(x) for x in m.group(1).split(",")])


def collision_setup(name, lattice):
    """(C ABI collision id, omega_minus, MRT rates) of a workload's collision operator.  TRT: odd-moment rate from the 'magic' parameter
    3/16.  MRT (moment space): shear moments relax with omega (same viscosity as the BGK / TRT runs), the bulk moment
    with 1.19, every ghost moment with 1.2 -- a common three-rate parametrisation; the kernel then projects the six shear / bulk rows only
    (rows at the most common rate are skipped), all conserved moments untouched.  LBM_BENCH_MRT_GHOSTS=cycle gives every third ghost
    moment 1.2 / 1.4 / 1.98 instead (d'Humieres et al. 2002): 17 of 23 rows projected."""
    from lbm_b200.cases import trt_omega_minus
    om_minus = trt_omega_minus(OMEGA)
    kinds = mrt_kinds(lattice)
    rates = np.full(27, OMEGA)
    rates[:len(kinds)][kinds == 2] = 1.19
    ghost = np.nonzero(kinds == 3)[0]
    if os.environ.get("LBM_BENCH_MRT_GHOSTS") == "cycle":
        rates[ghost] = np.array([1.2, 1.4, 1.98])[np.arange(len(ghost)) % 3]
    else:
        rates[ghost] = 1.2
    return {"bgk": 0, "trt": 1, "mrt": 2}[name], om_minus, rates


def cpu_sample(lattice, sample_size, budget_s, omp_collide=False, workload_name="box"):
    """Time the CPU restatement of the reference step (oracle port) on a bounded sample of the same workload."""
    from oracle import oracle
    wl = workload(sample_size, lattice) if workload_name == "box" else case_workload(workload_name, sample_size)
    o = oracle.Oracle(wl["ndim"], wl["ndist"], wl["nghbr"], OMEGA)
    if workload_name != "box":
        coll, om_minus, rates = collision_setup(wl["collision"], wl["lattice"])
        o.set_collision(coll, om_minus, rates)
    apply_bcs(o, wl)
    o.set_omp_collide(omp_collide)
    o.init()
    o.step(1)
    t0 = time.perf_counter()
    o.step(2)
    per = (time.perf_counter() - t0) / 2
    steps = int(max(3, min(200, budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    n = wl["nghbr"].shape[0]
    o.close()
    return n * steps / dt / 1e6, steps, n, oracle.threads()


def reference_binary_d2q9(level=9, steps=60):
    """The reference binary itself (oracle/_ref/lbm_ref, built from the reference's sources) on the largest kind of case it can
    run: D2Q9 BGK fp64, (2^level)^2 box, periodic x, bounce-back walls, moving top wall -- timed by the reference's own
    'Computation' timer (src/lbm/solver.cpp:57,308-319) with all host cores.  Reported beside the D3Q19 port number because the
    reference cannot run the D3Q19 workload at all.  Returns None if the binary is not there."""
    import re
    import shutil
    binary = os.path.join(ROOT, "oracle", "_ref", "lbm_ref")
    if not os.path.exists(binary):
        return None
    cfg = {"dim": 2, "partitionLevel": level, "uniformLevel": level, "maxRfnmtLvl": level, "maxNoCells": 4 ** level * 2, "outputDir": "out",
           "gridFileName": "g", "output": {"format": "VTKB", "cellFilter": "leafCells", "type": "points", "outputValues": ["level"]},
           "geometry": {"cube": {"type": "box", "A": [0.0, 0.0], "B": [1.0, 1.0]}},
           "solver": {"type": "lbm", "method": "bgk", "relaxation": 0.6, "ma": 0.01, "maxSteps": steps, "info_interval": 10 ** 6,
                      "solution_interval": 10 ** 9, "output_dir": "out", "convergence": 0.0, "assumeAxisAligned": True,
                      "boundary": {"cube": {"+x": {"type": "periodic", "generateBndry": False, "connection": "cube_-x"},
                                            "-x": {"type": "periodic", "generateBndry": False, "connection": "cube_+x"},
                                            "+y": {"type": "wall", "model": "bounceback", "tangentialVelocity": LID_U},
                                            "-y": {"type": "wall", "model": "bounceback"}}}}}
    tmp = tempfile.mkdtemp(prefix="lbm_refbin_")
    try:
        json.dump(cfg, open(os.path.join(tmp, "case.json"), "w"))
        cores = os.cpu_count() or 1
        env = dict(os.environ, OMP_NUM_THREADS=str(cores))
        r = subprocess.run([binary, "case.json"], cwd=tmp, env=env, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            return {"error": f"reference binary exited {r.returncode}"}
        log = open(os.path.join(tmp, "lbm_log")).read()
        m = re.search(r"Computation\s+([0-9.eE+-]+) \[sec\]", log)
        if not m:
            return {"error": "no Computation timer in lbm_log"}
        sec = float(m.group(1))
        cells = 4 ** level
        return {"value": cells * steps / sec / 1e6, "unit": "MLUPS", "cores": cores, "kind": "reference",
                "sample": f"reference binary, D2Q9 BGK fp64 {2 ** level}^2 box ({cells} cells), {steps} steps, its own 'Computation' timer "
                          f"({sec:.3f} s); the reference has no 3D LBM"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    if args.workload == "box":
        # the configuration of the GPU arm itself (default 256^3): ~0.8 s per step on 16 cores, so --steps 20 --warmup 5 ends in well
        # under a minute; tables from plain numpy, the CUDA library is never loaded in this arm
        sample = args.size
        wl = workload(sample, args.lattice, native=False)
        args.collision = "bgk"
    else:
        sample = min(args.cpu_size, 64, args.size)
        wl = case_workload(args.workload, sample)
        args.lattice, args.collision = wl["lattice"], wl["collision"]
    o = oracle.Oracle(wl["ndim"], wl["ndist"], wl["nghbr"], OMEGA)
    n = wl["nghbr"].shape[0]
    del wl["nghbr"]
    if args.workload != "box":
        coll, om_minus, rates = collision_setup(args.collision, args.lattice)
        o.set_collision(coll, om_minus, rates)
    apply_bcs(o, wl)
    o.init()
    o.step(args.warmup)
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    mlups = n * args.steps / dt / 1e6
    unit = "MLUPS"
    line = {
        "impl": "reference", "metric": "MLUPS", "value": mlups, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, "CPU port of the reference time step (oracle/lbm_oracle.c), OpenMP like the reference"),
        "cpu_baseline": {"value": mlups, "unit": unit, "cores": oracle.threads(), "kind": "port",
                         "sample": f"{args.lattice} {sample}^{wl['ndim']} {args.workload} ({n} cells: the whole workload of the GPU arm), same "
                                   f"BCs/omega/collision, {args.steps} steps; the reference binary cannot run D3Q19 (SURVEY section 0)"},
        "e2e": {"value": mlups, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line["cpu_baseline"]["reference_binary_d2q9"] = reference_binary_d2q9() if args.workload == "box" else None
    emit(line)


HALO_TEXT = {"p2p": "peer-to-peer (CUDA IPC mailboxes over NVLink, copy engines, one launch per step)", "nccl": "ncclSend/ncclRecv"}


def config_dict(args, note):
    ndim, ndist = LATTICES[args.lattice]
    if args.impl == "reference" and args.workload != "box":
        args = argparse.Namespace(**{**vars(args), "size": min(args.cpu_size, 64, args.size)})   # the size that arm actually runs
    if args.workload != "box":
        what = {"sphere": "flow past a sphere (radius L/10 at the centre of a cube, pressure in-/outlet on -x/+x, bounce-back walls and "
                          "sphere; BASELINE.json configs[3], 3D form of test/sphere/sphere_ns.json)",
                "step": "channel with a step (block on the upper wall, pressure in-/outlet, bounce-back walls; BASELINE.json configs[4], "
                        "3D form of test/step/step_ns.json)"}[args.workload]
        coll_note = " (MRT rates: shear omega, bulk 1.19, ghost moments " + ("1.2/1.4/1.98 in turn" if os.environ.get("LBM_BENCH_MRT_GHOSTS") == "cycle" else "1.2") + ")" if args.collision == "mrt" else ""
        return {"workload": f"{args.workload}-3D {args.size}^3 {args.lattice} {args.collision.upper()} {getattr(args, 'precision', 'fp64')}: {what}, omega={OMEGA:.6f}{coll_note}",
                "lattice": args.lattice, "collision": args.collision, "arithmetic": args.arithmetic,
                "l2_policy": "inputs larger than L2 (no flush needed)" if args.size >= 128 else "SMALL CASE: fits L2, not a bandwidth measurement",
                "parallelism": (f"one cube of {args.size}^3 cells cut into {args.gpus} contiguous SFC ranges (fixed total size), "
                                f"{HALO_TEXT[getattr(args, 'halo', 'p2p')]} halo exchange of outgoing populations every step") if args.gpus > 1 else "single GPU", "note": note}
    return {"workload": f"bench-3D cube {args.size}^{ndim} {args.lattice} BGK {getattr(args, 'precision', 'fp64')} (BASELINE.json configs[2]; SURVEY 8d S3): "
                        f"periodic x, bounce-back walls, moving lid u={LID_U}, omega={OMEGA:.6f}",
            "cells_per_gpu": args.size ** ndim, "lattice": args.lattice, "collision": "bgk", "arithmetic": args.arithmetic,
            "l2_policy": "inputs larger than L2 (no flush needed)", "parallelism": (f"one box of {'x'.join(map(str, global_shape(args.size, ndim, args.gpus)))} cells cut into {args.gpus} contiguous "
                            f"SFC ranges, {HALO_TEXT[getattr(args, 'halo', 'p2p')]} halo exchange of outgoing populations every step")
            if args.gpus > 1 else "single GPU", "note": note}


def parity_check(rank, world, local, dist, torch, lbm_b200, halo):
    """Driver-visible parity of the (multi-rank) path: 5 STRICT steps of a D3Q19 box with 64^3 cells per rank, cut into `world`
    contiguous SFC ranges and exchanged over NCCL exactly like the timed run; rank 0 compares the owned m_f of all ranks, bit for bit,
    with the single-domain CPU oracle (the checker, never the thing measured) and reports the SHA-256 of both."""
    import hashlib
    per, steps = 64, 5
    wl = workload(per, "D3Q19", rank, world)
    s = lbm_b200.Solver(3, 19, wl["nghbr"], OMEGA, arithmetic=lbm_b200.STRICT, device=local, track_vars=0,
                        stream=torch.cuda.current_stream().cuda_stream)
    apply_bcs(s, wl)
    if world > 1:
        from lbm_b200.capi import comm_unique_id
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        wl["lp"].apply_halo(s)
        s.comm_init(uid[0], rank, world)
    s.init()
    if world > 1 and halo == "p2p":
        s.p2p_connect(dist, wl["lp"])
    s.step(steps)
    s.synchronize()
    mine = torch.from_numpy(np.ascontiguousarray(s.f[:wl["n_owned"]])).cuda()
    parts = [mine]
    if world > 1:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        torch.cuda.synchronize()
        dist.barrier()
    s.close()
    if rank != 0:
        return None
    from oracle import oracle
    oracle.build()
    shape = wl["shape"]
    ref = workload(per, "D3Q19", native=False, shape=shape)
    o = oracle.Oracle(3, 19, ref["nghbr"], OMEGA)
    apply_bcs(o, ref)
    o.init()
    o.step(steps)
    got = torch.cat(parts).cpu().numpy()
    want = np.ascontiguousarray(o.f)
    same = got.shape == want.shape and np.array_equal(got, want)
    out = {"result": "bit-identical" if same else "DIFFERENT", "what": f"owned m_f of {world} rank(s) after {steps} STRICT fp64 steps, D3Q19 box "
           f"{'x'.join(map(str, shape))} ({per}^3 per rank), vs the single-domain CPU oracle",
           "sha256_gpu": hashlib.sha256(got.tobytes()).hexdigest()[:16], "sha256_oracle": hashlib.sha256(want.tobytes()).hexdigest()[:16]}
    o.close()
    return out


def run_ours(args):
    import torch
    import lbm_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    arithmetic = lbm_b200.FAST if args.arithmetic == "fast" else lbm_b200.STRICT
    parity = None if args.no_parity else parity_check(rank, world, local, dist if world > 1 else None, torch, lbm_b200, args.halo)
    t_setup = time.perf_counter()
    if args.workload == "box":
        wl = workload(args.size, args.lattice, rank, world)
        args.collision = "bgk"
    else:
        wl = case_workload(args.workload, args.size, rank, world)
        args.lattice, args.collision = wl["lattice"], wl["collision"]
    ndim, ndist = LATTICES[args.lattice]
    n = wl["n_owned"]
    n_local = wl["nghbr"].shape[0]
    stream = torch.cuda.current_stream().cuda_stream
    coll, om_minus, rates = collision_setup(args.collision, args.lattice)
    precision = lbm_b200.FP32 if args.precision == "fp32" else lbm_b200.FP64
    s = lbm_b200.Solver(ndim, ndist, wl["nghbr"], OMEGA, arithmetic=arithmetic, device=local, track_vars=0, stream=stream,
                        collision=coll, omega_minus=om_minus, mrt_rates=rates, precision=precision)
    apply_bcs(s, wl)
    if world > 1:
        from lbm_b200.capi import comm_unique_id
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        wl["lp"].apply_halo(s)
        s.comm_init(uid[0], rank, world)
    s.init()
    if world > 1 and args.halo == "p2p" and not s.p2p_connect(dist, wl["lp"]):
        args.halo = "nccl"   # some rank cannot use the mailboxes (velocity halo of a pressure boundary across a cut)
    nghbr_keep = wl["nghbr"] if (args.conv_interval > 0 or not args.no_strict) else None   # the residual-mode / STRICT runs below set up more solvers
    del wl["nghbr"]
    t_setup = time.perf_counter() - t_setup
    st0 = s.stats()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    # the driver's default is 5 warm-up steps (~5 ms): too short for a device that has just been idle to reach its steady state.
    # Extra untimed steps of the same workload come first; the W warm-up steps and the K timed steps follow exactly as asked.
    s.step(args.prewarm)
    s.step(args.warmup)
    barrier()
    l0 = s.stats()["launches"]
    ms_total, ms_main = s.step_timed(args.steps)
    launches = s.stats()["launches"] - l0
    barrier()
    if world > 1:
        t = torch.tensor([ms_total, ms_main], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_main = float(t[0]), float(t[1])
    n_global = args.size ** ndim * world if args.workload == "box" else wl["n_global"]
    value = n_global * args.steps / (ms_total * 1e-3) / 1e6
    # ---- the same steps with the residual bookkeeping the reference does (convergenceCondition every conv_interval steps,
    # src/lbm/solver.cpp:233-263): m_vars / m_varsold are written by the fused kernel on the steps the residual needs, the reduction
    # and (partitioned) its all-reduce run inside the timed region
    with_residual = None
    if args.conv_interval > 0:
        s.close()
        s = lbm_b200.Solver(ndim, ndist, nghbr_keep, OMEGA, arithmetic=arithmetic, device=local, track_vars=args.conv_interval, stream=stream,
                            collision=coll, omega_minus=om_minus, mrt_rates=rates, precision=precision)
        apply_bcs(s, wl)
        if world > 1:
            uid = [comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            wl["lp"].apply_halo(s)
            s.comm_init(uid[0], rank, world)
        s.init()
        if world > 1 and args.halo == "p2p":
            s.p2p_connect(dist, wl["lp"])
        s.step(args.conv_interval * max(1, args.warmup // args.conv_interval))
        s.residual()   # warm-up of the reduction (and, partitioned, of NCCL's all-reduce channels)
        s.step(args.conv_interval)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done, res = 0, None
        while done < args.steps:
            k = min(args.conv_interval, args.steps - done)
            s.step(k)
            done += k
            if k == args.conv_interval:
                res = s.residual()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        with_residual = {"value": n_global * args.steps / (float(t[0]) * 1e-3) / 1e6, "unit": "MLUPS", "conv_interval": args.conv_interval,
                         "residual": None if res is None else [float(x) for x in res[0]], "diverged": None if res is None else bool(res[1])}

    # ---- the same steps in STRICT arithmetic (the ABI's default: bit-identical to the reference's CPU solver in fp64)
    strict = None
    if args.arithmetic == "fast" and not args.no_strict and world == 1 and nghbr_keep is not None and n <= 40_000_000:  # a second solver must fit
        s2 = lbm_b200.Solver(ndim, ndist, nghbr_keep, OMEGA, arithmetic=lbm_b200.STRICT, device=local, track_vars=0, stream=stream,
                             collision=coll, omega_minus=om_minus, mrt_rates=rates, precision=precision)
        apply_bcs(s2, wl)
        s2.init()
        s2.step(args.prewarm + args.warmup)
        barrier()
        ms_s, ms_sm = s2.step_timed(args.steps)
        s2.close()
        strict = {"value": n_global * args.steps / (ms_s * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_launch": ms_sm / args.steps,
                  "frac": st0["bytes_per_cell_alg"] * n * args.steps / (ms_sm * 1e-3) / 1e9 / measured_peak()[0],
                  "note": "same workload, LBM_B200_STRICT: the reference's operation order, never contracted, divisions kept (by constants: "
                          "correctly rounded through two fused multiply-adds) -- bit-identical to the reference in fp64"}
    clocks = sampler.stop() if sampler else None   # NVML queries contend with the driver: none while the host-side e2e path is timed

    # ---- e2e: state in pinned host buffers, through the C ABI: upload m_fold (the input of a time step; m_f is overwritten by the
    # collision before anything reads it, solver.cpp:601-613), K steps, download the fields
    e2e = None
    if not args.no_e2e:
        fold_host = torch.empty((n_local, ndist), dtype=torch.float64, pin_memory=True)
        mom_host = torch.empty((n_local, ndim + 1), dtype=torch.float64, pin_memory=True)
        import ctypes as C
        lib = s._lib
        lib.lbm_b200_get_populations(s._h, None, C.c_void_p(fold_host.data_ptr()))
        k = args.steps
        barrier()
        b0 = s.stats()
        t0 = time.perf_counter()
        rc = lib.lbm_b200_set_populations(s._h, None, fold_host.numpy())
        assert rc == 0, lib.lbm_b200_last_error()
        s.step(k)
        rc = lib.lbm_b200_get_moments(s._h, mom_host.numpy())
        assert rc == 0, lib.lbm_b200_last_error()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        b1 = s.stats()
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        e2e = {"value": n_global * k / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": (b1["h2d_bytes"] - b0["h2d_bytes"]) / k, "d2h_bytes_per_step": (b1["d2h_bytes"] - b0["d2h_bytes"]) / k,
               "region": f"lbm_b200_set_populations(pinned m_fold) + {k} x lbm_b200_step + lbm_b200_get_moments(pinned)",
               "finite": bool(torch.isfinite(mom_host[:n]).all())}

    # ---- output path, after everything that is timed for the headline: never allowed to cost the line
    output_path = None
    if world == 1 and not args.no_output and n <= (1 << 25):
        try:
            output_path = output_path_sample(s)
        except Exception as e:  # noqa: BLE001 -- an optional measurement
            output_path = {"error": f"{type(e).__name__}: {e}"[:300]}

    if world > 1:
        # orderly teardown: every rank destroys its NCCL communicator and leaves the process group together
        barrier()
        s.close()
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    b_alg = st0["bytes_per_cell_alg"]
    achieved = b_alg * n * args.steps / (ms_main * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(args.lattice, args.size) if args.workload == "box" else None, "peak_source": peak_src,
            "kernel": f"lbm::k_step_fast (persistent tile pipeline: cp.async pull -> shared memory -> BC + moments + {args.collision.upper()} collide in place -> 128-bit copy-out)",
            "bytes_per_cell_alg": b_alg, "cells_per_launch": n, "ms_per_launch": ms_main / args.steps}
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle
        oracle.build()
        csize = args.cpu_size if args.workload == "box" else min(args.cpu_size, 64, args.size)
        v, steps, nc, cores = cpu_sample(args.lattice, csize, args.cpu_budget, workload_name=args.workload)
        cpu = {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port",
               "sample": f"{args.lattice} {csize}^{ndim} {args.workload} ({nc} cells), same BCs/omega/collision, {steps} steps of oracle/lbm_oracle.c "
                         f"(reference algorithm; serial collision pass like src/lbm/solver.cpp:601)",
               "reference_binary_d2q9": reference_binary_d2q9() if args.workload == "box" else None}
    line = {
        "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak" if args.workload == "box" else "strong",
        "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
        "config": config_dict(args, f"{st0['cells_fast']} of {st0['ncells']} cells on the index-free chunk path; setup {t_setup:.1f} s"),
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "with_residual": with_residual, "strict": strict, "parity": parity, "prewarm_steps": args.prewarm, "output_path": output_path,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--workload", default="box", choices=["box", "sphere", "step"],
                    help="box: BASELINE.json configs[2] (default, the metric's configuration); sphere / step: configs[3] / configs[4] in 3D, "
                         "--size = cells per side of the whole cube")
    ap.add_argument("--lattice", default="D3Q19", choices=sorted(LATTICES))
    ap.add_argument("--arithmetic", default="fast", choices=["fast", "strict"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"],
                    help="fp64: the reference's arithmetic type and the metric's configuration (default); fp32: the opt-in of BASELINE.json's "
                         "north_star (populations in float, 2*Q*4 algorithmic bytes per cell; tolerance stated in tests/test_gpu_parity.py)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-size", type=int, default=128, dest="cpu_size")
    ap.add_argument("--cpu-budget", type=float, default=15.0, dest="cpu_budget")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--no-e2e", action="store_true", dest="no_e2e")
    ap.add_argument("--no-output", action="store_true", dest="no_output", help="skip the output-path measurement (output_path field)")
    ap.add_argument("--conv-interval", type=int, default=10, dest="conv_interval",
                    help="second measurement with the reference's residual bookkeeping every N steps inside the timed region (0: skip)")
    ap.add_argument("--no-parity", action="store_true", dest="no_parity")
    ap.add_argument("--no-strict", action="store_true", dest="no_strict", help="skip the extra measurement in STRICT arithmetic")
    ap.add_argument("--prewarm", type=int, default=40, help="extra untimed steps before the W warm-up steps (device steady state)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU halo exchange: p2p = CUDA IPC mailboxes + copy engines + flag words, one launch per step (default); "
                         "nccl = ncclSend / ncclRecv with an outer and an inner launch")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
