#!/usr/bin/env python
"""Experiment: L contexts on ONE GPU, each rendering its own whole-frame iterations (stride L) from its own host thread, so that the
latency-bound tail of one iteration's bounce loop overlaps the full-width head of another.  Prints Msamples/s for L = 1, 2, 3."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from etx_tracer_b200 import scenes
from etx_tracer_b200.api import GPUVCM

workload = sys.argv[1] if len(sys.argv) > 1 else "C2"
per_lane = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sd = scenes.config(workload)
for lanes in (1, 2, 3):
    gs = [GPUVCM(sd, flavor="fast") for _ in range(lanes)]
    for k, g in enumerate(gs):
        g.set_iteration_stride(lanes)
        g.run(k)
    def work(g, n):
        for _ in range(n):
            g.iterate()
        g.wait()
    for phase, n in (("warmup", 3), ("timed", per_lane)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(g, n)) for g in gs]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"lanes={lanes}: {lanes * per_lane} iterations in {dt*1e3:.1f} ms -> {sd.width * sd.height * lanes * per_lane / dt / 1e6:.3f} Msamples/s", flush=True)
    for g in gs:
        g.close()
