#!/usr/bin/env python
"""Generates tests/golden/*.npz from the oracle built out of the reference's own headers (run where /root/reference exists).

kat.npz         : known-answer vectors of the integer / scalar pieces (Sampler, spectral sampling, to_rgb, hash cell, offset_ray, blue noise)
oracle_c1_32.npz: a 32x32, 3-iteration render of config C1 by the parity oracle (film layers + per-path sampler states + light-vertex pool digest)
oracle_c2_32.npz: same for config C2 (spectral, dielectric sphere)
oracle_c3_24.npz, oracle_c4_24.npz, oracle_c5_24.npz, oracle_vmf_24.npz: the scenes of tests/golden_scenes.py (BASELINE configs 3-5 with their full geometry at a
                  small film, the vMF diffuse material box)
Existing files are kept (zip metadata would change their bytes); delete one to regenerate it.
oracle_pt_c2_32.npz: config C2 at 32x32, 3 iterations of the PATH TRACER oracle (camera image, normal / albedo layers, sampler end states)
trace_c2.npz    : 4096 rays against the C2 scene -> (tri,u,v,t, sampler state)
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from etx_tracer_b200 import scenes, structs as S
from oracle import oracle_py

OUT = os.path.join(ROOT, "tests", "golden")
rng = np.random.default_rng(20260923)

def kat():
    a = np.concatenate([[0, 1, 262143, 4294967295], rng.integers(0, 2**32, 252, dtype=np.uint64)]).astype(np.uint32)
    b = np.concatenate([[0, 2, 15, 1023], rng.integers(0, 4096, 252)]).astype(np.uint32)
    seeds, vals = oracle_py.sampler_kat(a, b, 16)
    x = rng.random(512).astype(np.float32)
    d = dict(sampler_a=a, sampler_b=b, sampler_seeds=seeds, sampler_values=vals, x=x)
    wl = (390 + x * 440).astype(np.float32)
    d["wl"] = wl
    d["spectral_sample"] = oracle_py.math_kat(7, x)
    d["sampling_pdf"] = oracle_py.math_kat(8, wl)
    for k, fn in (("to_rgb_x", 9), ("to_rgb_y", 10), ("to_rgb_z", 11)):
        d[k] = oracle_py.math_kat(fn, wl)
    px = np.floor(x * 16384).astype(np.float32); si = np.floor(x[::-1] * 256).astype(np.float32)
    d["bn_pixel"], d["bn_sample"] = px, si
    d["bn_dim0_x"] = oracle_py.math_kat(14, px, si); d["bn_dim4_y"] = oracle_py.math_kat(15, px, si)
    lib = oracle_py.load("parity")
    cells = rng.integers(-1000, 1000, (256, 3)).astype(np.int32); cells[0] = (1, 2, 3); cells[1] = (-1, 0, 7)
    d["cells"] = cells
    d["cell_index"] = np.array([lib.oracle_grid_cell_index((1 << 20) - 1, int(c[0]), int(c[1]), int(c[2])) for c in cells], np.uint32)
    p = (rng.random((256, 3)) * 4 - 2).astype(np.float32); n = rng.normal(size=(256, 3)).astype(np.float32); n /= np.linalg.norm(n, axis=1, keepdims=True)
    p[0] = (1, -2, 0.01); n[0] = (0, 1, 0)
    out = np.zeros_like(p)
    for i in range(256):
        lib.oracle_offset_ray(oracle_py._p(p[i]), oracle_py._p(n[i]), oracle_py._p(out[i]))
    d["offset_p"], d["offset_n"], d["offset_out"] = p, n, out
    for nm, fn, xs, ys in (("sin", 0, x * 12 - 6, None), ("cos", 1, x * 12 - 6, None), ("exp", 2, x * 40 - 20, None), ("log", 3, x * 10 + 1e-6, None), ("pow", 4, x * 3, x[::-1] * 8 - 2),
                           ("acos", 5, x * 2 - 1, None), ("atan2", 6, x - 0.5, x[::-1] - 0.5), ("atan", 12, x * 20 - 10, None), ("asin", 13, x * 2 - 1, None)):
        xs = np.ascontiguousarray(xs, np.float32); ys = None if ys is None else np.ascontiguousarray(ys, np.float32)
        d[f"pm_{nm}_x"] = xs
        if ys is not None: d[f"pm_{nm}_y"] = ys
        d[f"pm_{nm}"] = oracle_py.math_kat(fn, xs, ys)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **d)

def render(name, sd, iters, opts=None):
    if os.path.exists(os.path.join(OUT, name)):
        return None
    o = oracle_py.Oracle(sd)
    if opts is not None:
        o.set_options(opts)
    o.begin(0); o.run(iters, threads=1)
    lv = o.buffer(S.BUF_LV_POS, np.float32)
    d = dict(film_result=o.film(S.FILM_RESULT), film_camera=o.film(S.FILM_CAMERA), film_light=o.film(S.FILM_LIGHT),
             light_sampler=o.buffer(S.BUF_LIGHT_SAMPLER, np.uint32), camera_sampler=o.buffer(S.BUF_CAMERA_SAMPLER, np.uint32),
             light_path_count=o.buffer(S.BUF_LIGHT_PATH_COUNT, np.uint32), lv_pos=lv, iterations=np.array([iters]))
    np.savez_compressed(os.path.join(OUT, name), **d)
    return o

def trace(sd):
    if os.path.exists(os.path.join(OUT, "trace_c2.npz")):
        return
    o = oracle_py.Oracle(sd)
    n = 4096
    org = (rng.random((n, 3)) * np.array([1.8, 1.8, 4.6]) + np.array([-0.9, 0.1, -0.9])).astype(np.float32)
    d3 = rng.normal(size=(n, 3)).astype(np.float32); d3 /= np.linalg.norm(d3, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32); rays[:, 0:3] = org; rays[:, 3] = 2.28997145e-4; rays[:, 4:7] = d3; rays[:, 7] = 3.0e38
    seeds = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    uvt, tri, seeds_out = o.trace(rays, seeds)
    np.savez_compressed(os.path.join(OUT, "trace_c2.npz"), rays=rays, seeds=seeds, uvt=uvt, tri=tri, seeds_out=seeds_out)

def render_pt(name, sd, iters):
    """the path tracer: run_path_iteration compiled from the reference + the restated CPUPathTracing driver"""
    if os.path.exists(os.path.join(OUT, name)):
        return
    o = oracle_py.Oracle(sd)
    o.set_integrator(S.INTEGRATOR_PT)
    o.pt_set_options(S.default_pt_options())
    o.begin(0); o.run(iters, threads=1)
    np.savez_compressed(os.path.join(OUT, name), film_camera=o.film(S.FILM_CAMERA), film_normals=o.film(S.FILM_NORMALS), film_albedo=o.film(S.FILM_ALBEDO),
                        camera_sampler=o.buffer(S.BUF_CAMERA_SAMPLER, np.uint32), iterations=np.array([iters]))

if not os.path.exists(os.path.join(OUT, "kat.npz")):
    kat()
render_pt("oracle_pt_c2_32.npz", scenes.cornell_box(32, 32, samples=256, spectral=True, sphere=True), 3)
render("oracle_c1_32.npz", scenes.cornell_box(32, 32, samples=16, spectral=False), 3)
render("oracle_c2_32.npz", scenes.cornell_box(32, 32, samples=256, spectral=True, sphere=True), 3)
trace(scenes.cornell_box(32, 32, samples=256, spectral=True, sphere=True))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_scenes
for name, (factory, iters, opts) in golden_scenes.SCENES.items():
    render(name, factory(), iters, opts() if opts else None)
for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))
