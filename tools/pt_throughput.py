#!/usr/bin/env python
"""Throughput of the device path tracer on a BASELINE workload, next to the reference's CPU path tracer (the oracle: run_path_iteration
compiled from the reference + the restated CPUPathTracing driver) on the box's host threads.  One JSON line.
usage: pt_throughput.py [C1..C5] [iterations] [cpu_iterations]   (metric: Msamples/s = W*H*iterations / device time of the iterations)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from etx_tracer_b200 import scenes, structs as S
from etx_tracer_b200.api import GPUPathTracing
what = sys.argv[1] if len(sys.argv) > 1 else "C3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cpu_iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sd = scenes.config(what)
g = GPUPathTracing(sd, flavor="fast", profile=True)
g.set_scene_settings(0.0, 0.0)  # every pixel stays active: a sample = one pixel-iteration
g.render(2)  # warm-up (iteration 0 uses the empty pixel filter; clocks, allocations)
st = g.render(iters, first_iteration=2)
img = g.film(S.FILM_CAMERA)[..., :3]
n = sd.width * sd.height
out = {"metric": "Msamples/s path tracer", "workload": what, "film": [sd.width, sd.height], "iterations": iters, "value": n * iters / st["total_time"] / 1e6,
       "ms_per_iteration": 1e3 * st["total_time"] / iters, "finite": bool(np.isfinite(img).all()), "mean": float(img.mean()),
       "kernel_ms_per_iteration": {k: round(v[0] / iters, 3) for k, v in sorted(g.kernel_times().items(), key=lambda kv: -kv[1][0]) if v[1]},
       "counters": {k: v for k, v in g.counters().items() if v}}
g.close()
try:  # several iterations in flight (etxb_group lanes, adaptive sampling off): the throughput form, like the VCM bench's `value`
    from etx_tracer_b200.api import GPUVCMGroup
    grp = GPUVCMGroup(sd, lanes=4, flavor="fast")
    grp.set_integrator(S.INTEGRATOR_PT)
    grp.render(4)
    grp.run(4)
    k = max(iters, 12)
    grp.enqueue(k)
    grp.wait()
    st4 = grp.status()
    img4 = grp.film(S.FILM_CAMERA)[..., :3]
    out["in_flight_4"] = {"value": n * k / st4["total_time"] / 1e6, "ms_per_iteration": 1e3 * st4["total_time"] / k, "iterations": k, "mean": float(img4.mean()),
                          "finite": bool(np.isfinite(img4).all())}
    grp.close()
except Exception as e:
    out["in_flight_4"] = {"error": f"{type(e).__name__}: {e}"[:300]}
if cpu_iters > 0:
    from oracle import oracle_py
    if oracle_py.available("native"):
        o = oracle_py.Oracle(sd, "native")
        o.set_integrator(S.INTEGRATOR_PT)
        o.pt_set_options(S.default_pt_options())
        o.set_scene_settings(0.0, 0.0)
        o.begin(2)
        threads = os.cpu_count() or 1
        t = o.run(cpu_iters, threads=threads)
        out["cpu_reference"] = {"value": n * cpu_iters / t / 1e6, "cores": threads, "kind": "reference", "sample": f"{cpu_iters} iteration(s) of the same scene and film, {t:.1f} s"}
        ref = o.film(S.FILM_CAMERA)[..., :3]
        out["cpu_reference"]["mean"] = float(ref.mean())
print(json.dumps(out))
