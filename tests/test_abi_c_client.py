"""include/lbm_b200.h is a C header: a C99 client (tests/c/abi_client.c) compiles with -pedantic, links against liblbm_b200.so and drives
the GPU-free entry points (box tables, SFC key, native partition, inspection-only solver and its device plan)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_c99_client_builds_and_runs(tmp_path):
    exe = str(tmp_path / "abi_client")
    pkg = os.path.join(ROOT, "lbm_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(HERE, "c", "abi_client.c"), "-o", exe, "-L", pkg, "-llbm_b200", f"-Wl,-rpath,{pkg}"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "abi_client ok" in r.stdout, r.stdout + r.stderr
