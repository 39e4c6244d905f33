"""CPU checks of the path tracer's ORACLE: the reference's run_path_iteration compiled in place, driven by the restated
CPUPathTracing loop and Film members (oracle/oracle_vcm.cxx: run_pt_iteration, film_accumulate_pt, film_estimate_noise_levels).

The reference ships no tests or vectors for this path either, so the restated parts are pinned three ways: (1) against an independent
estimator of the same image — the VCM oracle with merging off (unbiased bidirectional connections), (2) against a numpy restatement of
Film::accumulate_camera_image / estimate_noise_levels in the reference's own scatter form, (3) by a committed golden render.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bit_equal, golden
from etx_tracer_b200 import scenes, structs as S


def _pt(oracle_mod, sd, iterations, flavor="native", threads=4, settings=None, first=0, options=None):
    if not oracle_mod.available(flavor):
        flavor = "parity"
    o = oracle_mod.Oracle(sd, flavor)
    o.set_integrator(S.INTEGRATOR_PT)
    opts = S.default_pt_options()
    for k, v in (options or {}).items():
        opts[k] = v
    o.pt_set_options(opts)
    if settings is not None:
        o.set_scene_settings(*settings)
    o.begin(first)
    o.run(iterations, threads=threads)
    return o


def test_path_tracer_agrees_with_connect_only_vcm(oracle_mod):
    """Two different estimators of the same integral (both unbiased): image means within 1.5 % at 96 spp on a 40x40 Cornell box."""
    sd = scenes.cornell_box(40, 40, samples=16, spectral=False)
    o = _pt(oracle_mod, sd, 96)
    pt = o.film(S.FILM_CAMERA)[..., :3].astype(np.float64)
    v = oracle_mod.Oracle(sd, "native" if oracle_mod.available("native") else "parity")
    opts = S.default_vcm_options()
    opts["options"] = S.VCM_CONNECT_ONLY
    v.set_options(opts)
    v.begin(0)
    v.run(96, threads=4)
    bd = v.film(S.FILM_RESULT)[..., :3].astype(np.float64)
    assert abs(pt.mean() - bd.mean()) / bd.mean() < 1.5e-2
    # BSDF sampling alone (light sampling off, and MIS off: the reference keeps weighting emitter hits while "mis" is on, path_tracing_shared.hxx:351)
    # estimates the same image, with more noise
    alt = _pt(oracle_mod, sd, 256, options=dict(nee=0, mis=0)).film(S.FILM_CAMERA)[..., :3].astype(np.float64)
    assert abs(alt.mean() - bd.mean()) / bd.mean() < 6e-2
    # with MIS off the reference adds light sampling and emitter hits at full weight (path_tracing_shared.hxx:318-319, 351-352): brighter, by design
    dbl = _pt(oracle_mod, sd, 32, options=dict(mis=0)).film(S.FILM_CAMERA)[..., :3].astype(np.float64)
    assert dbl.mean() > 1.3 * bd.mean()
    # direct hits off + light sampling on loses only the directly visible emitter
    nd = _pt(oracle_mod, sd, 8, options=dict(direct=0)).film(S.FILM_CAMERA)[..., :3]
    assert nd.max() < pt.max()


def test_first_hit_layers(oracle_mod):
    """view_normal / view_albedo are written at path_length == 1 only (path_tracing_shared.hxx:376-379); Film::layer shows normals as n/2 + 1/2."""
    sd = scenes.cornell_box(32, 32, samples=16, spectral=True, sphere=True)
    o = _pt(oracle_mod, sd, 4, flavor="parity", threads=1)
    nrm = o.film(S.FILM_NORMALS)[..., :3] * 2.0 - 1.0
    ln = np.linalg.norm(nrm, axis=-1)
    assert (ln < 1.0 + 1e-4).all() and (ln > 0.3).mean() > 0.95  # a mean of unit normals (pixel-filter jitter crosses edges)
    alb = o.film(S.FILM_ALBEDO)[..., :3]
    assert np.isfinite(alb).all() and alb.mean() > 0.05
    info = o.buffer(S.BUF_PIXEL_INFO, np.uint32)
    assert (info == 4).all()


def _numpy_noise_estimate(cam, var, conv, tmp, threshold):
    """film.cxx:233-330 in the reference's own (scatter) form."""
    h, w = conv.shape
    err = np.zeros((h, w), np.float32)
    conv, tmp = conv.copy(), tmp.copy()
    active = ~conv
    diff = np.abs(cam - var).astype(np.float32)
    e_diff = (diff[..., 0] + diff[..., 1]) + diff[..., 2]
    a = np.abs(cam).astype(np.float32)
    e_norm = (a[..., 0] + a[..., 1]) + a[..., 2]
    level = (e_diff / (np.where(e_norm < 1.0, np.sqrt(e_norm), e_norm) + np.float32(1e-6))).astype(np.float32)
    c = level < np.float32(threshold)
    err[active] = level[active]
    conv[active] = c[active]
    tmp[active] = c[active]
    for y, x in zip(*np.nonzero(~conv)):
        tmp[y, max(0, x - 5):min(w, x + 5)] = False
    for y, x in zip(*np.nonzero(~tmp)):
        conv[max(0, y - 5):min(h, y + 5), x] = False
    return err, conv, tmp


def test_adaptive_sampling_follows_the_reference_passes(oracle_mod):
    """The oracle's history after 33 iterations (one estimate, after iteration 32) against the numpy restatement applied to the oracle's own layers
    after 33 iterations without an estimate; then 41 iterations: converged pixels stop being sampled."""
    sd = scenes.cornell_box(40, 36, samples=16, spectral=False)
    thr = 0.25
    plain = _pt(oracle_mod, sd, 33, flavor="parity", threads=1, settings=(0.0, 0.0))
    cam, var = plain.film(S.FILM_CAMERA)[..., :3], plain.film(S.FILM_CAMERA_ADAPTIVE)[..., :3]
    z = np.zeros((36, 40), bool)
    err, conv, tmp = _numpy_noise_estimate(cam, var, z, z, thr)
    o = _pt(oracle_mod, sd, 33, flavor="parity", threads=1, settings=(thr, 0.0))
    assert bit_equal(o.film(S.FILM_CAMERA), plain.film(S.FILM_CAMERA))  # the estimate after iteration 32 changes nothing rendered so far
    info = o.buffer(S.BUF_PIXEL_INFO, np.uint32).reshape(36, 40)
    assert ((info & S.PIXEL_COUNT_MASK) == 33).all()
    assert np.array_equal((info & S.PIXEL_CONVERGED) != 0, conv) and np.array_equal((info & S.PIXEL_TMP) != 0, tmp)
    assert 0 < conv.sum() < conv.size
    np.testing.assert_allclose(o.buffer(S.BUF_PIXEL_ERROR, np.float32).reshape(36, 40), err, rtol=2e-5, atol=1e-7)  # numpy rounds the three-term sums and the division in float32 steps of its own
    assert o.pt_status()["active_pixels"] == int((err < thr).sum())
    o41 = _pt(oracle_mod, sd, 41, flavor="parity", threads=1, settings=(thr, 0.0))
    info41 = o41.buffer(S.BUF_PIXEL_INFO, np.uint32).reshape(36, 40)
    counts = info41 & S.PIXEL_COUNT_MASK
    assert set(np.unique(counts)) <= {33, 35, 37, 39, 41} and counts.min() == 33 and counts.max() == 41
    # a pixel that converged at the first estimate and stayed converged kept its 33 samples and its colour
    stayed = (counts == 33)
    assert stayed.any() and bit_equal(o41.film(S.FILM_CAMERA)[stayed], o.film(S.FILM_CAMERA)[stayed])
    # threads do not change the result (pixels are independent; the passes only clear flags)
    o41t = _pt(oracle_mod, sd, 41, flavor="parity", threads=4, settings=(thr, 0.0))
    assert bit_equal(o41t.film(S.FILM_CAMERA), o41.film(S.FILM_CAMERA)) and np.array_equal(o41t.buffer(S.BUF_PIXEL_INFO, np.uint32), info41.ravel())


def test_running_means(oracle_mod):
    """Film::accumulate_camera_image: colour mean over all samples, adaptive mean over the even-indexed ones (film.cxx:199-222)."""
    sd = scenes.cornell_box(24, 24, samples=16, spectral=False)
    singles = []
    for k in range(4):
        singles.append(_pt(oracle_mod, sd, 1, flavor="parity", threads=1, first=k).film(S.FILM_CAMERA)[..., :3].astype(np.float64))
    o = _pt(oracle_mod, sd, 4, flavor="parity", threads=1)
    # NOTE iteration 0 of a run uses the empty pixel filter; singles[k>0] started a run at k, so their filter differs from iteration k of `o`:
    # compare only what is filter-independent — the structure of the two means — on iteration 0 and the film after one iteration
    one = _pt(oracle_mod, sd, 1, flavor="parity", threads=1)
    assert bit_equal(one.film(S.FILM_CAMERA), one.film(S.FILM_CAMERA_ADAPTIVE))
    np.testing.assert_allclose(one.film(S.FILM_CAMERA)[..., :3], singles[0], rtol=0, atol=0)
    cam, var = o.film(S.FILM_CAMERA)[..., :3].astype(np.float64), o.film(S.FILM_CAMERA_ADAPTIVE)[..., :3].astype(np.float64)
    assert np.abs(cam - var).max() > 0.0 and abs(cam.mean() - var.mean()) / cam.mean() < 0.5


def test_golden_render_matches_live_oracle(oracle_mod):
    ref = golden("oracle_pt_c2_32.npz")
    sd = scenes.cornell_box(32, 32, samples=256, spectral=True, sphere=True)
    o = _pt(oracle_mod, sd, int(ref["iterations"][0]), flavor="parity", threads=1)
    assert bit_equal(o.film(S.FILM_CAMERA), ref["film_camera"]) and bit_equal(o.buffer(S.BUF_CAMERA_SAMPLER, np.uint32), ref["camera_sampler"])
    assert bit_equal(o.film(S.FILM_NORMALS), ref["film_normals"]) and bit_equal(o.film(S.FILM_ALBEDO), ref["film_albedo"])


def test_gather_form_of_the_dilation_passes_equals_the_reference_scatter_form():
    """The device runs Film::estimate_noise_levels' two scatter passes (film.cxx:283-320) as GATHERS (kernels_pt.cuh: k_film_noise_rows / _columns: a pixel
    loses `tmp` when a pixel x in [p - 4, p + 5] of its row has not converged; it loses `converged` when a pixel y in [p - 4, p + 5] of its column has no `tmp`).
    A numpy model of the gathers against the scatter form of _numpy_noise_estimate, on random masks including the film borders."""
    rng = np.random.default_rng(3)
    for (h, w, density) in ((1, 1, 0.5), (7, 5, 0.3), (36, 40, 0.05), (36, 40, 0.6), (64, 3, 0.2)):
        conv = rng.random((h, w)) > density
        tmp = conv.copy()
        # scatter form (the reference)
        t_ref, c_ref = tmp.copy(), conv.copy()
        for y, x in zip(*np.nonzero(~conv)):
            t_ref[y, max(0, x - 5):min(w, x + 5)] = False
        for y, x in zip(*np.nonzero(~t_ref)):
            c_ref[max(0, y - 5):min(h, y + 5), x] = False
        # gather form (the kernels)
        t_dev = tmp.copy()
        for y in range(h):
            for x in range(w):
                if tmp[y, x] and (~conv[y, max(0, x - 4):min(w - 1, x + 5) + 1]).any():
                    t_dev[y, x] = False
        c_dev = conv.copy()
        for y in range(h):
            for x in range(w):
                if conv[y, x] and (~t_dev[max(0, y - 4):min(h - 1, y + 5) + 1, x]).any():
                    c_dev[y, x] = False
        assert np.array_equal(t_dev, t_ref) and np.array_equal(c_dev, c_ref), (h, w, density)
