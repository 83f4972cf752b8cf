#!/usr/bin/env python3
"""Prepare a scratch copy of the reference sources for the oracle build.

TEST INFRASTRUCTURE.  Nothing here is imported by the product.  The reference
(/root/reference, SFCMM/LBM v0.0.2) is compiled from its own sources; this script
copies `src/` to a scratch directory OUTSIDE the repository and applies three
non-arithmetic accommodations:

  1. `config.h` is generated from `src/config.h.in` (what CMake's configure_file does),
     with ENABLE_BACKTRACE off (its vendored boost subset needs system boost headers).
  2. `lbm/bnd/bnd_wall.h`: `&m_dirichletValue[index][0]` -> `m_dirichletValue[index].data()`.
     With libstdc++ >= 13 `std::array<double,0>::operator[]` is `__builtin_trap()`;
     the pointer is never dereferenced for no-slip walls, so arithmetic is unchanged.
  3. `lbm/solver.cpp`: a read-only dump hook (dump_hook.inc) that writes the neighbour
     table, property bits, centres, surfaces and the raw f/fold/vars arrays when
     SFCMM_DUMP is set.

usage: patch_reference.py <reference_root> <scratch_dir>
"""
import os
import re
import shutil
import sys


def main() -> None:
    ref, out = sys.argv[1], sys.argv[2]
    here = os.path.dirname(os.path.abspath(__file__))
    srcp = os.path.join(out, "srcp")
    if os.path.isdir(srcp):
        shutil.rmtree(srcp)
    os.makedirs(out, exist_ok=True)
    shutil.copytree(os.path.join(ref, "src"), srcp)

    # 1. config.h
    cfg = open(os.path.join(ref, "src", "config.h.in")).read()
    cfg = (cfg.replace("@PROJECT_VERSION@", "0.0.2")
              .replace("@CMAKE_CXX_COMPILER_VERSION@", "13")
              .replace("@CMAKE_CXX_COMPILER_ID@", "GNU")
              .replace("@CMAKE_BUILD_TYPE@", "Release"))
    cfg = re.sub(r"^#define ENABLE_BACKTRACE", "// #define ENABLE_BACKTRACE", cfg, flags=re.M)
    os.makedirs(os.path.join(out, "gen"), exist_ok=True)
    open(os.path.join(out, "gen", "config.h"), "w").write(cfg)

    # 2. libstdc++-13 trap in the no-slip wall BC
    p = os.path.join(srcp, "lbm", "bnd", "bnd_wall.h")
    s = open(p).read()
    n = s.count("&m_dirichletValue[index][0]")
    assert n >= 1, "bnd_wall.h: pattern not found"
    s = s.replace("&m_dirichletValue[index][0]", "m_dirichletValue[index].data()")
    open(p, "w").write(s)

    # 3. dump hook
    p = os.path.join(srcp, "lbm", "solver.cpp")
    s = open(p).read()
    hook = open(os.path.join(here, "dump_hook.inc")).read()
    s = s.replace("using namespace std;\n", "using namespace std;\n" + hook + "\n", 1)
    a = "  loadConfiguration();\n  allocateMemory();\n"
    assert s.count(a) == 1
    s = s.replace(a, "  loadConfiguration();\n  LBM_B200_DUMP_SETUP();\n  allocateMemory();\n")
    b = "    timeStep();\n\n"
    assert s.count(b) == 1
    s = s.replace(b, "    timeStep();\n    if(lbm_b200_dump::wantStep(m_timeStep + 1)) { LBM_B200_DUMP_STATE(\"_\" + std::to_string(m_timeStep + 1)); }\n\n")
    c = "  executePostprocess(pp::HOOK::ATEND);\n"
    assert s.count(c) == 1
    s = s.replace(c, c + "  LBM_B200_DUMP_STATE(\"\");\n")
    open(p, "w").write(s)
    print("patched copy in", srcp)


if __name__ == "__main__":
    main()
