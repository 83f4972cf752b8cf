#!/usr/bin/env python3
"""Golden solution files: what the reference binary writes to out/<solution_filename>_<step>.vtp.

Runs oracle/_ref/lbm_ref on the same shortened configurations as make_golden.py (config_json of each fixture) and records,
per case, the name, size and SHA-256 of the final solution file; the two smallest files are kept whole (gzip) so that a
mismatch can be located.  The input of the writer -- the moments of the final m_fold -- is reproduced in the tests by the
oracle, which is bit-exact on these cases (tests/test_oracle_golden.py).

With --full the reference's UNMODIFIED test configurations (config_orig_json; the 15 Navier-Stokes cases of test/run.sh) are run to
their end and the digest of the final file goes to vtp/index_full.json -- what `lbm` must write after the same run on the GPU
(tests/test_host_run_gpu.py).

Needs /root/reference and oracle/_ref/lbm_ref (build container only).  usage: python tests/golden/make_vtp_golden.py [--full]
"""
import glob
import gzip
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
BIN = os.path.join(ROOT, "oracle", "_ref", "lbm_ref")
KEEP_WHOLE = ("couette", "poiseuille_bnd_pressure")


def main():
    full = "--full" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]  # case names: update just these entries of the index
    index_path = os.path.join(HERE, "vtp", "index_full.json" if full else "index.json")
    index = json.load(open(index_path)) if only and os.path.exists(index_path) else {}
    os.makedirs(os.path.join(HERE, "vtp"), exist_ok=True)
    for path in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        name = os.path.basename(path)[:-4]
        if only and name not in only:
            continue
        if full and "_ml_" in name:  # multi-level variants are not reference test cases; their short runs are pinned instead
            continue
        cfg = json.loads(str(np.load(path)["config_orig_json" if full else "config_json"]))
        if full:
            if name in ("step_ns", "sphere_ns"):  # not part of run.sh (no analytic solution, 10^4..10^5 steps)
                continue
            cfg["solver"]["solution_interval"] = 10 ** 9  # only the final, forced file
        tmp = tempfile.mkdtemp(prefix="lbm_vtp_")
        try:
            json.dump(cfg, open(os.path.join(tmp, "case.json"), "w"))
            r = subprocess.run([BIN, "case.json"], cwd=tmp, env=dict(os.environ, OMP_NUM_THREADS="1" if full or "_ml_" in name else "2"), capture_output=True,
                               text=True)
            if r.returncode != 0:
                raise RuntimeError(f"{name}: reference exited {r.returncode}\n{r.stderr[-2000:]}")
            stem = cfg["solver"].get("solution_filename", "solution")
            files = [f for f in glob.glob(os.path.join(tmp, cfg["solver"].get("output_dir", "out"), f"{stem}_*.vtp"))]
            assert len(files) == 1, (name, files)
            data = open(files[0], "rb").read()
            index[name] = {"file": os.path.basename(files[0]), "bytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
            line = os.path.join(tmp, "line.csv")  # postprocessing type "line" (postprocessing.h:104-114), written to the cwd
            if full and os.path.exists(line):
                index[name]["line_csv"] = open(line).read()
            if name in KEEP_WHOLE and not full:
                with gzip.GzipFile(os.path.join(HERE, "vtp", f"{name}.vtp.gz"), "wb", mtime=0) as f:
                    f.write(data)
            print(name, {k: v for k, v in index[name].items() if k != "line_csv"})
        finally:
            shutil.rmtree(tmp)
    json.dump(index, open(index_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
