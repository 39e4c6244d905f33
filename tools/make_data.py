#!/usr/bin/env python
"""Generates etx_tracer_b200/data/*.npz from the reference (run in the dev container; needs /root/reference).

Outputs (committed, small):
  color_tables.npz : xyz_441x3 (CIE 2006 table, render/shared/spectrum.hxx:28), rgb_response_391x3 (render/host/spectrum.cxx:399),
                     y_integral (spectrum::kYIntegral)
  bluenoise.npz    : sobol (256x256 u8), scrambling_<spp>, ranking_<spp> (128*128*8 u8) for spp in 1..256 (thirdparty/bluenoise/*.hpp)
  spectra.npz      : named IOR spectra from bin/spectrum/**/*.spd resampled by the reference loader (SpectralDistribution::load_from_file)
"""
import os, re, sys, glob
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from etx_tracer_b200 import structs as S
from oracle import oracle_py

REF = os.environ.get("ETX_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "etx_tracer_b200", "data")
os.makedirs(OUT, exist_ok=True)
lib = oracle_py.load("parity")
P = oracle_py._p

xyz = np.zeros((441, 3), np.float32); rgbr = np.zeros((391, 3), np.float32); yint = np.zeros(1, np.float32)
lib.oracle_color_tables(P(xyz), P(rgbr), P(yint))
np.savez_compressed(os.path.join(OUT, "color_tables.npz"), xyz_441x3=xyz, rgb_response_391x3=rgbr, y_integral=yint)

def parse_table(text, name):
    m = re.search(name + r"\[[^\]]*\]\s*=\s*\{([^}]*)\}", text)
    return np.array([int(v) for v in m.group(1).split(",") if v.strip()], dtype=np.uint8)

bn = {}
bn["sobol"] = parse_table(open(os.path.join(REF, "thirdparty/bluenoise/bluenoise_shared.hpp")).read(), "sobol_256spp_256d")
assert bn["sobol"].size == 256 * 256
for spp in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    t = open(os.path.join(REF, f"thirdparty/bluenoise/samplerBlueNoiseErrorDistribution_128x128_OptimizedFor_2d2d2d2d_{spp}spp.hpp")).read()
    bn[f"scrambling_{spp}"] = parse_table(t, "scramblingTile")
    bn[f"ranking_{spp}"] = parse_table(t, "rankingTile")
    assert bn[f"scrambling_{spp}"].size == 128 * 128 * 8 and bn[f"ranking_{spp}"].size == 128 * 128 * 8
np.savez_compressed(os.path.join(OUT, "bluenoise.npz"), **bn)

spectra = {}
for path in sorted(glob.glob(os.path.join(REF, "bin/spectrum/*/*.spd"))):
    kind = os.path.basename(os.path.dirname(path))
    if kind == "emission":
        continue
    name = os.path.splitext(os.path.basename(path))[0]
    eta = np.zeros(1, S.SPECTRUM); k = np.zeros(1, S.SPECTRUM)
    cls = lib.oracle_spectrum_load_ior(path.encode(), P(eta), P(k))
    spectra[f"{name}.eta_power"] = eta["entries"]["power"][0].copy()
    spectra[f"{name}.eta_rgb"] = eta["integrated"][0].copy()
    spectra[f"{name}.k_power"] = k["entries"]["power"][0].copy()
    spectra[f"{name}.k_rgb"] = k["integrated"][0].copy()
    spectra[f"{name}.cls"] = np.array([cls], np.uint32)
    assert eta["entry_count"][0] == 441 and np.all(eta["entries"]["wavelength"][0] == np.arange(390, 831, dtype=np.float32))
# a few blackbody emitters used by the synthetic scenes (scene files say e.g. "nblackbody 2700 scale 5", cornellbox.mtl)
for (t, scale) in ((2700, 5.0), (5800, 1.0), (6500, 1.0), (12000, 0.1)):
    s = np.zeros(1, S.SPECTRUM)
    lib.oracle_spectrum_blackbody(t, scale, 1, P(s))
    spectra[f"nblackbody_{t}_{scale}.power"] = s["entries"]["power"][0].copy()
    spectra[f"nblackbody_{t}_{scale}.rgb"] = s["integrated"][0].copy()
# the three atmosphere spectra every scene's spectrum pool starts with (scattering::init -> init_default_values, scene_representation.cxx:206-213):
# entries 2, 3, 4 of the pool of any scene the reference's own loader has read
rs = oracle_py.ReferenceScene("assets/cornellbox/cornellbox.json")
pool = np.frombuffer((__import__("ctypes").c_char * (5 * S.SPECTRUM.itemsize)).from_address(int(rs.scene["spectrums"]["a"][0])), dtype=S.SPECTRUM)
for idx, key in ((2, "rayleigh"), (3, "mie"), (4, "ozone")):
    assert int(rs.scene[key + "_spectrum"][0]) == idx and int(pool[idx]["entry_count"]) == 441
    spectra[f"atmosphere_{key}.power"] = pool[idx]["entries"]["power"].copy()
    spectra[f"atmosphere_{key}.rgb"] = pool[idx]["integrated"].copy()
rs.close()
np.savez_compressed(os.path.join(OUT, "spectra.npz"), **spectra)
for f in os.listdir(OUT):
    print(f, os.path.getsize(os.path.join(OUT, f)))
