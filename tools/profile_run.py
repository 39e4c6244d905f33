#!/usr/bin/env python
"""One or two iterations of a workload, for use under ncu (never a bench number)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from etx_tracer_b200 import scenes
from etx_tracer_b200.api import GPUVCM
res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sd = scenes.cornell_box(res, res, samples=256, spectral=True, sphere=True)
g = GPUVCM(sd, flavor="fast")
g.render(iters)
print(g.status(), g.counters())
