"""The oracle against known answers: SURVEY.md §4's table (produced from the reference headers with g++ -O1 and glibc's libm)
and the committed golden vectors (tools/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import bit_equal, golden
from etx_tracer_b200 import structs as S


def test_sampler_matches_survey_table(oracle_mod):
    # Sampler::random_seed / next (render/shared/sampler.hxx:54,66) — integer arithmetic, must be exact
    a = np.array([0, 1, 262143, 4294967295], np.uint32)
    b = np.array([0, 2, 15, 1023], np.uint32)
    seeds, vals = oracle_mod.sampler_kat(a, b, 4)
    assert list(seeds[:, 0]) == [1947998333, 212161807, 2995896715, 1313741542]
    assert list(seeds[1, 1:]) == [0xB2DFB132, 0x41CF080E, 0xCA99235C, 0x2D3B5351]
    np.testing.assert_array_equal(vals[1], np.array([0.6987257, 0.257065296, 0.791399121, 0.176686406], np.float32))


def test_spectral_sampling_matches_survey_table(oracle_mod):
    # SpectralQuery::spectral_sample / sampling_pdf (spectrum.hxx:234,219); the survey used glibc's atanh/cosh, the oracle the
    # portable ones -> agree to float rounding
    rnd = np.array([0.0, 0.25, 0.5, 0.75, 0.999], np.float32)
    wl = oracle_mod.math_kat(7, rnd)
    np.testing.assert_allclose(wl, [390.0, 487.528534, 550.729492, 620.156738, 825.905090], rtol=2e-6)
    pdf = oracle_mod.math_kat(8, wl)
    np.testing.assert_allclose(pdf, [1.49466214e-3, 3.46212741e-3, 3.90689401e-3, 2.82895798e-3, 2.41757501e-4], rtol=2e-5)


def test_to_rgb_hash_offset_ray_match_survey_table(oracle_mod):
    lib = oracle_mod.load("parity")
    wl = np.array([550.0], np.float32)
    rgb = [oracle_mod.math_kat(fn, wl)[0] for fn in (9, 10, 11)]
    np.testing.assert_allclose(rgb, [0.00183429325, 0.0112881949, -0.00143251161], rtol=1e-6)
    assert lib.oracle_grid_cell_index((1 << 20) - 1, 1, 2, 3) == 363078
    assert lib.oracle_grid_cell_index((1 << 20) - 1, -1, 0, 7) == 849314
    p = np.array([1, -2, 0.01], np.float32)
    n = np.array([0, 1, 0], np.float32)
    out = np.zeros(3, np.float32)
    lib.oracle_offset_ray(oracle_mod._p(p), oracle_mod._p(n), oracle_mod._p(out))
    np.testing.assert_array_equal(out, np.array([1.0, -1.99996948, 0.00999999978], np.float32))


def test_struct_sizes_match_survey_table(oracle_mod):
    lib = oracle_mod.load("parity")
    expected = {0: 528, 1: 176, 2: 56, 3: 32, 4: 200, 5: 48, 6: 32, 7: 3552, 8: 112, 9: 80, 10: 32, 11: 176, 12: 112}
    for k, v in expected.items():
        assert lib.oracle_sizeof(k) == v
    for name, (dt, size) in S.EXPECTED_SIZES.items():
        assert dt.itemsize == size, name


def test_oracle_reproduces_golden_kat(oracle_mod):
    g = golden("kat.npz")
    seeds, vals = oracle_mod.sampler_kat(g["sampler_a"], g["sampler_b"], 16)
    assert bit_equal(seeds, g["sampler_seeds"]) and bit_equal(vals, g["sampler_values"])
    assert bit_equal(oracle_mod.math_kat(7, g["x"]), g["spectral_sample"])
    assert bit_equal(oracle_mod.math_kat(8, g["wl"]), g["sampling_pdf"])
    assert bit_equal(oracle_mod.math_kat(9, g["wl"]), g["to_rgb_x"])
    assert bit_equal(oracle_mod.math_kat(14, g["bn_pixel"], g["bn_sample"]), g["bn_dim0_x"])


@pytest.mark.parametrize("name,fn,ref", [
    ("sin", 0, np.sin), ("cos", 1, np.cos), ("exp", 2, np.exp), ("log", 3, np.log), ("acos", 5, np.arccos), ("atan", 12, np.arctan), ("asin", 13, np.arcsin)])
def test_portable_math_is_within_one_ulp(oracle_mod, name, fn, ref):
    # etx_tracer_b200/csrc/portable_math.h through the oracle's libm override
    g = golden("kat.npz")
    x = g[f"pm_{name}_x"]
    got = oracle_mod.math_kat(fn, x)
    assert bit_equal(got, g[f"pm_{name}"])
    want = ref(x.astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(got.astype(np.float64) - want) <= 1.0 * ulp + 1e-45)


def test_portable_pow_atan2_are_within_one_ulp(oracle_mod):
    g = golden("kat.npz")
    got = oracle_mod.math_kat(4, g["pm_pow_x"], g["pm_pow_y"])
    want = np.power(g["pm_pow_x"].astype(np.float64), g["pm_pow_y"].astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(got.astype(np.float64) - want) <= ulp)
    got = oracle_mod.math_kat(6, g["pm_atan2_x"], g["pm_atan2_y"])
    want = np.arctan2(g["pm_atan2_x"].astype(np.float64), g["pm_atan2_y"].astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(got.astype(np.float64) - want) <= ulp)
    # C99 special cases the microfacet walk hits (bsdf_external.hxx:70,93): pow(x, +-inf), pow(0, y), pow(x, 0)
    x = np.array([0.5, 2.0, 1.0, 0.0, 0.0, 0.3, 0.0], np.float32)
    y = np.array([np.inf, np.inf, np.inf, 2.0, -1.0, 0.0, 0.0], np.float32)
    np.testing.assert_array_equal(oracle_mod.math_kat(4, x, y), np.array([0.0, np.inf, 1.0, 0.0, np.inf, 1.0, 1.0], np.float32))
