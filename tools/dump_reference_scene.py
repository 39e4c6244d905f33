#!/usr/bin/env python
"""Dumps the Scene / Camera PODs the reference's OWN loader builds from its shipped Cornell asset (bin/assets/cornellbox: 138k triangles, a
fog volume behind a Boundary mesh, sun + sky, conductor box) into tests/golden/ref_cornell_40.npz, so that the GPU parity test for that
asset runs on the GPU box (which has no /root/reference).  Run where the reference tree exists:  python tools/dump_reference_scene.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from etx_tracer_b200 import pod_io  # noqa: E402
from oracle import oracle_py  # noqa: E402

rs = oracle_py.ReferenceScene("assets/cornellbox/cornellbox.json")
rs.resize(40, 40, [0.0, 1.000000238418579, 3.819999933242798], [0.0, 1.000000238418579, -6.179999351501465], [0.0, 0.9999999403953552, -0.0], 39.597755335771296)
out = os.path.join(ROOT, "tests", "golden", "ref_cornell_40.npz")
pod_io.dump(out, rs)
print(out, os.path.getsize(out) / 1e6, "MB,", rs.triangle_count, "triangles")
