#!/usr/bin/env python
"""A few path-tracer iterations of a workload, for use under ncu (never a bench number).  usage: profile_run_pt.py [C1..C5] [iterations]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from etx_tracer_b200 import scenes
from etx_tracer_b200.api import GPUPathTracing
what = sys.argv[1] if len(sys.argv) > 1 else "C3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = GPUPathTracing(scenes.config(what), flavor="fast")
g.set_scene_settings(0.0, 0.0)
g.render(iters)
print(g.status(), g.counters())
