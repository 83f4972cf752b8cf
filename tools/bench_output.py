"""Output path measured: the fields of one solution file -- cell filter, 15-decimal rounding, base64 -- through
lbm_b200_encode_output (device; the text comes back in one copy) beside the host writer's route (lbm_b200_get_moments to the host, then
vtk_writer.hpp's rounding + base64 on all host cores).  Both produce the same bytes (tests/test_gpu_parity.py, tests/test_host_run_gpu.py).
Prints one JSON line.  usage: python tools/bench_output.py [--size 256] [--reps 5]"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import bench
    import lbm_b200
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device")
    wl = bench.workload(a.size, "D3Q19")
    s = bench.apply_bcs(lbm_b200.Solver(3, 19, wl["nghbr"], bench.OMEGA, device=0, track_vars=0), wl)
    s.init()
    s.step(20)
    n = s.n
    so = os.path.join(tempfile.mkdtemp(), "liboutput_harness.so")
    here = os.path.join(ROOT, "tests", "c")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-Wno-unknown-pragmas",
                           "-I", os.path.join(here, "fake_cuda"), os.path.join(here, "output_harness.cpp"), "-o", so])
    H = C.CDLL(so)
    H.oh_host_encode.restype = C.c_int64
    H.oh_host_encode.argtypes = [C.c_void_p, C.c_int64, C.c_int]
    total = 4 * int(s._lib.lbm_b200_output_chars(n))
    pinned_mem = lbm_b200.HostBuffer(total)                  # lbm_b200_host_alloc
    pinned = pinned_mem.array
    dev, dev_pin, host_fetch, host_enc = [], [], [], []
    chars = 0
    for _ in range(a.reps + 1):
        t0 = time.perf_counter()
        text = s.encode_output(None, raw=True)               # into fresh pageable memory, like the host route's moments array
        t1 = time.perf_counter()
        m = s.moments()
        t2 = time.perf_counter()
        chars = H.oh_host_encode(m.ctypes.data, n, m.shape[1])
        t3 = time.perf_counter()
        textp = s.encode_output(None, out=pinned, raw=True)  # into page-locked memory
        t4 = time.perf_counter()
        assert sum(len(t) for t in text) == chars and all(np.array_equal(x, y) for x, y in zip(text, textp))
        dev.append(t1 - t0), host_fetch.append(t2 - t1), host_enc.append(t3 - t2), dev_pin.append(t4 - t3)
        del text, m
    med = lambda v: float(np.median(v[1:]))     # first repetition = warm-up (allocations, page faults)
    d, dp, hf, he = med(dev), med(dev_pin), med(host_fetch), med(host_enc)
    print(json.dumps({"metric": "output fields of one solution file", "workload": f"D3Q19 box {a.size}^3, {n} cells, 4 fields, all cells kept",
                      "text_bytes": int(chars), "device_encode_s": round(d, 4), "device_encode_pinned_s": round(dp, 4),
                      "device_Mcells_per_s": round(n / d / 1e6, 1), "device_pinned_Mcells_per_s": round(n / dp / 1e6, 1),
                      "host_route_s": round(hf + he, 4), "host_fetch_moments_s": round(hf, 4), "host_round_base64_s": round(he, 4),
                      "host_cores": os.cpu_count(), "host_Mcells_per_s": round(n / (hf + he) / 1e6, 1), "reps": a.reps}))


if __name__ == "__main__":
    main()
