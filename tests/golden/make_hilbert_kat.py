#!/usr/bin/env python3
"""Extract the reference's known-answer tests of hilbert::index (UnitTest/test_hilbert.cpp) into
tests/golden/hilbert_kat.json.  Needs /root/reference; the JSON it writes is committed."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
src = open("/root/reference/UnitTest/test_hilbert.cpp").read()
variables, kats = {}, []
for line in src.splitlines():
    m = re.search(r"VectorD<(\d)>\s+(\w+)\s*=\s*\{([^}]*)\}", line)
    if m:
        variables[m.group(2)] = (int(m.group(1)), [float(x) for x in m.group(3).split(",")])
    m = re.search(r"EXPECT_EQ\(hilbert::index<(\d)>\((\w+),\s*(\d+)\),\s*(\d+)\)", line)
    if m:
        nd, coords = variables[m.group(2)]
        kats.append({"ndim": nd, "x": coords, "level": int(m.group(3)), "index": int(m.group(4))})
json.dump({"source": "/root/reference/UnitTest/test_hilbert.cpp (EXPECT_EQ known answers of hilbert::index)", "kats": kats},
          open(os.path.join(HERE, "hilbert_kat.json"), "w"))
print(len(kats), "known answers")
