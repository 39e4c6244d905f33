#!/usr/bin/env python
"""Quick GPU-vs-oracle comparison used during bring-up (the pytest suite formalises these checks)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from etx_tracer_b200 import scenes, structs as S
from etx_tracer_b200.api import GPUVCM
from oracle import oracle_py

def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)

def compare(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}"); return False
    if a.dtype.kind == 'f':
        same = bits(a) == bits(b)
    else:
        same = a == b
    frac = same.mean() if same.size else 1.0
    extra = ""
    if a.dtype.kind == 'f' and same.size and frac < 1.0:
        d = np.abs(a.astype(np.float64) - b.astype(np.float64)); extra = f" max|d|={d.max():.3e} rel={d.sum()/max(np.abs(b).sum(),1e-30):.3e}"
    print(f"  {name}: identical {frac*100:.4f}% of {same.size}{extra}")
    return frac == 1.0

def run(flavor, sd, iters, label):
    print(f"== {label} [{flavor}] {sd.name} {sd.width}x{sd.height} tris={sd.triangle_count}")
    g = GPUVCM(sd, flavor=flavor)
    o = oracle_py.Oracle(sd, flavor="parity")
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2**32, 1000, dtype=np.uint64).astype(np.uint32); b = rng.integers(0, 1024, 1000).astype(np.uint32)
    gs, gv = g.debug_sampler(a, b, 8); os_, ov = oracle_py.sampler_kat(a, b, 8)
    compare("sampler seeds", gs, os_); compare("sampler values", gv, ov)
    x = rng.random(4096).astype(np.float32)
    for fn, nm, xs, ys in ((0,'sin',x*6.3,None),(1,'cos',x*6.3,None),(2,'exp',x*20-10,None),(3,'log',x+1e-3,None),(4,'pow',x,(x[::-1]*4).copy()),(5,'acos',x*2-1,None),
                           (6,'atan2',x-0.5,(x[::-1]-0.5).copy()),(7,'spectral_sample',x,None),(8,'sampling_pdf',390+x*440,None),(9,'to_rgb.x',390+x*440,None),(14,'bluenoise',np.floor(x*16384),np.floor(x[::-1]*256).copy())):
        compare(f"math {nm}", g.debug_math(fn, xs, ys), oracle_py.math_kat(fn, xs, ys))
    # rays: from camera position in random directions + random interior points
    n = 20000
    o3 = np.tile(np.array([0.0, 1.0, 3.5], np.float32), (n, 1)); o3[n//2:] = (rng.random((n - n//2, 3)) * np.array([1.8, 1.8, 1.8]) + np.array([-0.9, 0.1, -0.9])).astype(np.float32)
    d3 = rng.normal(size=(n, 3)).astype(np.float32); d3 /= np.linalg.norm(d3, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32); rays[:, 0:3] = o3; rays[:, 3] = 1e-4; rays[:, 4:7] = d3; rays[:, 7] = 3.0e38
    seeds = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    guvt, gtri, gseed = g.debug_trace(rays, seeds); ouvt, otri, oseed = o.trace(rays, seeds)
    compare("trace tri", gtri, otri); compare("trace uvt", guvt, ouvt); compare("trace seeds", gseed, oseed)
    t0 = time.time(); g.render(iters); tg = time.time() - t0
    o.begin(0); t0 = time.time(); o.run(iters, threads=1); to = time.time() - t0
    st = g.status(); print("  gpu status", st, f"wall {tg:.3f}s; oracle {to:.2f}s")
    ok = True
    for nm, bid, dt in (("light path count", S.BUF_LIGHT_PATH_COUNT, np.uint32), ("light path offset", S.BUF_LIGHT_PATH_OFFSET, np.uint32), ("wavelength", S.BUF_LIGHT_PATH_WAVELENGTH, np.float32),
                        ("light sampler end", S.BUF_LIGHT_SAMPLER, np.uint32), ("lv pos", S.BUF_LV_POS, np.float32), ("lv throughput", S.BUF_LV_THROUGHPUT, np.float32), ("lv mis", S.BUF_LV_MIS, np.float32),
                        ("camera sampler end", S.BUF_CAMERA_SAMPLER, np.uint32), ("camera value", S.BUF_CAMERA_GATHERED, np.float32)):
        ok &= compare(nm, g.buffer(bid, dt), o.buffer(bid, dt))
    for nm, layer in (("film camera", S.FILM_CAMERA), ("film light", S.FILM_LIGHT), ("film result", S.FILM_RESULT)):
        compare(nm, g.film(layer)[..., :3], o.film(layer)[..., :3])
    print("  counters gpu", g.counters()); print("  counters cpu", {k: int(v) for k, v in zip(o.counters().dtype.names, o.counters()[0])})
    np.save(f"gpurun_out/{label}_{flavor}_gpu.npy", g.film(0)); np.save(f"gpurun_out/{label}_{flavor}_cpu.npy", o.film(0))
    g.close(); o.close()

if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    for flavor in ("parity", "fast"):
        run(flavor, scenes.cornell_box(res, res, samples=16, spectral=False), 2, "c1")
        run(flavor, scenes.cornell_box(res, res, samples=256, spectral=True, sphere=True), 2, "c2")
