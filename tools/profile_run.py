#!/usr/bin/env python
"""A few iterations of a workload, for use under ncu (never a bench number).
usage: profile_run.py [C1|C2|C3|<cornell resolution>] [iterations] [resolution scale]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from etx_tracer_b200 import scenes
from etx_tracer_b200.api import GPUVCM
what = sys.argv[1] if len(sys.argv) > 1 else "C2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
if what.isdigit():
    sd = scenes.cornell_box(int(what), int(what), samples=256, spectral=True, sphere=True)
else:
    sd = scenes.config(what, scale)
g = GPUVCM(sd, flavor="fast")
g.render(iters)
print(g.status(), g.counters())
