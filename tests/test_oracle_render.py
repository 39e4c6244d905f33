"""The oracle's VCM driver against committed renders + properties of the restated driver (CPU only, small sizes)."""
import numpy as np
import pytest

from conftest import bit_equal, golden, rel_l2
from etx_tracer_b200 import scenes, structs as S


@pytest.mark.parametrize("name,kwargs", [
    ("oracle_c1_32.npz", dict(samples=16, spectral=False)),
    ("oracle_c2_32.npz", dict(samples=256, spectral=True, sphere=True)),
])
def test_oracle_render_is_pinned(oracle_mod, name, kwargs):
    g = golden(name)
    sd = scenes.cornell_box(32, 32, **kwargs)
    o = oracle_mod.Oracle(sd)
    o.begin(0)
    o.run(int(g["iterations"][0]), threads=1)
    assert bit_equal(o.buffer(S.BUF_LIGHT_SAMPLER, np.uint32), g["light_sampler"])
    assert bit_equal(o.buffer(S.BUF_CAMERA_SAMPLER, np.uint32), g["camera_sampler"])
    assert bit_equal(o.buffer(S.BUF_LIGHT_PATH_COUNT, np.uint32), g["light_path_count"])
    assert bit_equal(o.buffer(S.BUF_LV_POS, np.float32), g["lv_pos"])
    for layer, key in ((S.FILM_RESULT, "film_result"), (S.FILM_CAMERA, "film_camera"), (S.FILM_LIGHT, "film_light")):
        assert bit_equal(o.film(layer), g[key]), key
    img = o.film(S.FILM_RESULT)[..., :3]
    assert np.isfinite(img).all() and img.min() >= 0.0 and 0.05 < img.mean() < 1.0


@pytest.mark.parametrize("name", ["oracle_c3_24.npz", "oracle_c4_24.npz", "oracle_c5_24.npz", "oracle_vmf_24.npz"])
def test_oracle_renders_of_configs_3_to_5_are_pinned(oracle_mod, name):
    """The oracle (reference headers compiled in place) reproduces its committed renders of BASELINE configs 3-5 and the vMF diffuse box."""
    import golden_scenes
    factory, iters, opts = golden_scenes.SCENES[name]
    g = golden(name)
    o = oracle_mod.Oracle(factory())
    if opts:
        o.set_options(opts())
    o.begin(0)
    o.run(iters, threads=1)
    assert bit_equal(o.buffer(S.BUF_LIGHT_SAMPLER, np.uint32), g["light_sampler"])
    assert bit_equal(o.buffer(S.BUF_CAMERA_SAMPLER, np.uint32), g["camera_sampler"])
    assert bit_equal(o.buffer(S.BUF_LV_POS, np.float32), g["lv_pos"])
    for layer, key in ((S.FILM_RESULT, "film_result"), (S.FILM_CAMERA, "film_camera"), (S.FILM_LIGHT, "film_light")):
        assert bit_equal(o.film(layer), g[key]), key
    img = o.film(S.FILM_RESULT)[..., :3]
    assert np.isfinite(img).all() and img.min() >= 0.0 and img.mean() > 0.01


def test_oracle_threads_agree_on_the_camera_image(oracle_mod):
    # thread-local vertex vectors are appended in thread order -> the pool and every camera sample are independent of the thread count;
    # only the light image's float atomics may reorder
    sd = scenes.cornell_box(24, 24, samples=16, spectral=False)
    a, b = oracle_mod.Oracle(sd), oracle_mod.Oracle(sd)
    a.begin(0), b.begin(0)
    a.run(2, threads=1), b.run(2, threads=4)
    assert bit_equal(a.film(S.FILM_CAMERA), b.film(S.FILM_CAMERA))
    assert bit_equal(a.buffer(S.BUF_LV_POS, np.float32), b.buffer(S.BUF_LV_POS, np.float32))
    assert rel_l2(b.film(S.FILM_LIGHT), a.film(S.FILM_LIGHT)) < 1e-6


def test_oracle_connect_only_and_no_mis_options(oracle_mod):
    # VCMOptions::ConnectOnlyOptions (vcm_shared.hxx:33) is what BASELINE config 5 ("volumetric BDPT") uses
    sd = scenes.cornell_box(24, 24, samples=16, spectral=False)
    o = oracle_mod.Oracle(sd)
    full = S.default_vcm_options()
    full["initial_radius"] = 0.01  # merging is biased (consistent): keep the kernel small so both estimators agree at 24x24
    conn = S.default_vcm_options()
    conn["options"] = S.VCM_CONNECT_ONLY
    imgs = []
    for opts in (full, conn):
        o.set_options(opts)
        o.begin(0)
        o.run(6, threads=4)
        imgs.append(o.film(S.FILM_RESULT)[..., :3].astype(np.float64))
    # both estimators are unbiased for the same integral: means agree within Monte-Carlo noise
    assert abs(imgs[0].mean() - imgs[1].mean()) / imgs[0].mean() < 0.1


def test_native_oracle_agrees_statistically(oracle_mod):
    # liboracle_native.so = same sources, -O3 -march=x86-64-v3, glibc libm: used as the CPU baseline; different rounding -> tolerance
    if not oracle_mod.available("native"):
        pytest.skip("native oracle not built")
    try:
        oracle_mod.load("native")
    except OSError:
        pytest.skip("native oracle needs AVX2/FMA")
    sd = scenes.cornell_box(24, 24, samples=16, spectral=False)
    a, b = oracle_mod.Oracle(sd, "parity"), oracle_mod.Oracle(sd, "native")
    a.begin(0), b.begin(0)
    a.run(4, threads=2), b.run(4, threads=2)
    assert rel_l2(b.film(0)[..., :3], a.film(0)[..., :3]) < 0.05


def test_baseline_configs_4_and_5_render_on_the_oracle(oracle_mod):
    """The generators of BASELINE configs 4 (subsurface 'dragon' stand-in) and 5 (heterogeneous cloud, connect-only) produce what SURVEY.md
    8(d) names, and the reference's CPU VCM renders them (small film; the mesh / grid sizes are reduced only where the CPU suite needs it)."""
    c4 = scenes.sss_dragon(24, 24, target_triangles=60_000)
    assert 55_000 < c4.triangle_count < 65_000
    assert any(int(m["subsurface"]["cls"][0]) == 1 for m in c4.materials)  # random walk
    o = oracle_mod.Oracle(c4)
    o.begin(0)
    o.run(2, threads=4)
    img = o.film(S.FILM_RESULT)[..., :3]
    assert np.isfinite(img).all() and img.mean() > 0.01
    c5 = scenes.cloud_box(24, 24, grid=32)
    opts = S.default_vcm_options()
    opts["options"] = S.VCM_CONNECT_ONLY
    o = oracle_mod.Oracle(c5)
    o.set_options(opts)
    o.begin(0)
    o.run(4, threads=4)
    img = o.film(S.FILM_RESULT)[..., :3]
    assert np.isfinite(img).all() and img.mean() > 0.01
    assert int(o.counters()["merge_queries"][0]) == 0  # merging off = volumetric BDPT
    # the cloud attenuates what is behind it: the film's centre (through the cube) is darker than its top rows (open sky)
    assert img[8:16, 8:16].mean() < img[:4].mean()


def test_fbm_density_variants_are_normalised():
    for d in (scenes.fbm_density(16), scenes.fbm_density_fast(64)):
        assert d.dtype == np.float32 and float(d.max()) == 1.0 and float(d.min()) == 0.0
        assert 0.02 < float(d.mean()) < 0.3


def test_traversal_order_changes_the_image_only_statistically(oracle_mod):
    """SURVEY.md 8(c) tier C.  The reference draws one sampler value per CANDIDATE hit, so the candidate order of the ray caster (Embree there,
    the shared BVH here) decides which random numbers a path sees.  Rendering the same view of the same scene ROTATED by 33 degrees about the
    vertical axis (camera included: another axis-aligned BVH, other near/far decisions, another candidate order, other roundings) must change
    the image only like another set of samples does: the relative MSE between the two stays within 2x of the oracle's own run-to-run relative
    MSE at equal sample counts — and is not zero, i.e. the rotation really changed some sample paths."""
    import math
    ang = math.radians(33.0)
    rot = np.array([[math.cos(ang), 0.0, math.sin(ang)], [0.0, 1.0, 0.0], [-math.sin(ang), 0.0, math.cos(ang)]])

    def build(rotated):
        sd = scenes.cornell_box(32, 32, samples=256, spectral=True, sphere=True, sphere_segments=24, sphere_rings=13, finalize=False)
        origin, target, up = np.array([0.0, 1.0, 3.82]), np.array([0.0, 1.0, -6.18]), np.array([0.0, 1.0, 0.0])
        if rotated:
            for v in sd.vertices:
                for f in ("pos", "nrm", "tan", "btn"):
                    v[f] = (v[f].astype(np.float64) @ rot.T).astype(np.float32)
            for t in sd.triangles:
                t["geo_n"] = (t["geo_n"].astype(np.float64) @ rot.T).astype(np.float32)
            origin, target = rot @ origin, rot @ target
        sd.set_camera(origin, target, up, 32, 32, 39.597755335771296, clip_near=0.1, clip_far=100.0)
        return sd.finalize(samples=256, spectral=True)

    def render(sd, first):
        o = oracle_mod.Oracle(sd, "native")
        o.begin(first)
        o.run(24, threads=8)
        img = o.film(S.FILM_RESULT)[..., :3].astype(np.float64)
        o.close()
        return img

    a, b, a2 = render(build(False), 0), render(build(True), 0), render(build(False), 24)

    def rel_mse(x, y):
        m = 0.5 * (x + y)
        return float((((x - y) ** 2) / (m ** 2 + 1e-3)).mean())

    run_to_run, order_to_order = rel_mse(a, a2), rel_mse(a, b)
    # measured: 0.014 against 0.29 run to run — most rays meet ONE candidate (near-first traversal prunes the rest), so few paths change at all
    assert 1e-6 < order_to_order < 2.0 * run_to_run, (order_to_order, run_to_run)
    assert abs(a.mean() - b.mean()) / a.mean() < 0.05


def test_restated_ray_caster_finds_the_brute_force_closest_hit(oracle_mod):
    """The one piece of the oracle that is restated rather than compiled from the reference is the ray caster over the shared BVH (Embree is
    absent).  Independent check: for random rays the BVH traversal must return the closest triangle a float64 brute-force Moeller-Trumbore over
    ALL triangles finds (same triangle, or one at an indistinguishable distance), and miss exactly when brute force misses."""
    sd = scenes.cornell_box(16, 16, samples=16, spectral=False, sphere=True, sphere_segments=24, sphere_rings=13)
    o = oracle_mod.Oracle(sd)
    rng = np.random.default_rng(5)
    n = 3000
    org = (rng.random((n, 3)) * np.array([1.8, 1.8, 4.6]) + np.array([-0.9, 0.1, -0.9])).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    tmin = np.float32(2.28997145e-4)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = org, tmin, d, 3.0e38
    uvt, tri, _ = o.trace(rays, rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32))
    o.close()
    import ctypes as C
    nv, nt = int(sd.scene["vertices"]["count"][0]), int(sd.scene["triangles"]["count"][0])
    V = np.frombuffer((C.c_char * (nv * S.VERTEX.itemsize)).from_address(int(sd.scene["vertices"]["a"][0])), dtype=S.VERTEX)["pos"].astype(np.float64)
    T = np.frombuffer((C.c_char * (nt * S.TRIANGLE.itemsize)).from_address(int(sd.scene["triangles"]["a"][0])), dtype=S.TRIANGLE)["i"].astype(np.int64)
    a, b, c = V[T[:, 0]], V[T[:, 1]], V[T[:, 2]]
    e1, e2 = b - a, c - a
    best_t = np.full(n, np.inf)
    best_tri = np.full(n, -1, dtype=np.int64)
    for k in range(n):
        o64, d64 = org[k].astype(np.float64), d[k].astype(np.float64)
        p = np.cross(d64, e2)
        det = (e1 * p).sum(1)
        ok = np.abs(det) > 0
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        tv = o64 - a
        u = (tv * p).sum(1) * inv
        qv = np.cross(tv, e1)
        v = (qv @ d64) * inv
        t = (e2 * qv).sum(1) * inv
        hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > float(tmin))
        if hit.any():
            tt = np.where(hit, t, np.inf)
            best_tri[k], best_t[k] = int(np.argmin(tt)), float(tt.min())
    found = tri != 0xFFFFFFFF
    # rays that graze an edge can fall on either side in float32 vs float64: tolerate a handful, require exactness elsewhere
    miss_mismatch = int((found != np.isfinite(best_t)).sum())
    assert miss_mismatch <= 3, miss_mismatch
    both = found & np.isfinite(best_t)
    same_tri = tri[both].astype(np.int64) == best_tri[both]
    close_t = np.abs(uvt[both, 2].astype(np.float64) - best_t[both]) <= 1e-4 * np.maximum(1.0, best_t[both])
    assert close_t.mean() > 0.999, float(close_t.mean())
    assert (same_tri | close_t).mean() > 0.999 and same_tri.mean() > 0.99, (float(same_tri.mean()), float(close_t.mean()))
