"""The reference's OWN scene loader (compiled in place into oracle/_ref/libreference_loader.so) against this repo's scene generators and the
oracle.  CPU only; skipped where the reference tree is absent (the GPU box): assets and IOR tables are never copied into the repo.

What it pins: (1) the Scene / Camera PODs the reference builds from its shipped Cornell asset are the bytes `etxb_upload_scene` and the
oracle consume (sizes, pointer graph, a render through the oracle); (2) the conventions `etx_tracer_b200/scenes.py` uses for materials,
emission, media and the camera are the loader's — records of same-named materials are compared field by field."""
import numpy as np
import pytest

from conftest import bit_equal
from etx_tracer_b200 import scenes, structs as S

CAMERA = dict(origin=[0.0, 1.000000238418579, 3.819999933242798], target=[0.0, 1.000000238418579, -6.179999351501465], up=[0.0, 0.9999999403953552, -0.0],
              fov=39.597755335771296)


@pytest.fixture(scope="module")
def cornell(oracle_mod):
    if not oracle_mod.ReferenceScene.available():
        pytest.skip("reference tree or oracle/_ref/libreference_loader.so not present")
    rs = oracle_mod.ReferenceScene("assets/cornellbox/cornellbox.json")
    yield rs
    rs.close()


def _view(array_view, dtype):
    """numpy view over an ArrayView {pointer, count} of a Scene POD."""
    import ctypes as C
    n = int(array_view["count"][0])
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_char * (n * dtype.itemsize)).from_address(int(array_view["a"][0])), dtype=dtype)


def test_loader_reads_the_shipped_cornell_asset(cornell):
    sc = cornell.scene
    assert cornell.triangle_count == 138318 and (cornell.width, cornell.height) == (640, 640)
    assert int(sc["samples"][0]) == 32 and int(sc["max_path_length"][0]) == 1023 and int(sc["random_path_termination"][0]) == 6
    mats = _view(sc["materials"], S.MATERIAL)
    assert len(mats) == 11
    assert int(mats[cornell.material_index("fog")]["cls"]) == S.MAT_BOUNDARY and int(mats[cornell.material_index("fog")]["int_medium"]) == 0
    assert int(mats[cornell.material_index("tallbox")]["cls"]) == S.MAT_CONDUCTOR
    assert int(sc["mediums"]["count"][0]) == 1 and int(sc["emitter_instances"]["count"][0]) == 12


def test_generator_conventions_match_the_loader(cornell):
    """scenes.py builds its materials the way the loader does: same record for the same directives."""
    sc = cornell.scene
    mats = _view(sc["materials"], S.MATERIAL)
    spectra = _view(sc["spectrums"], S.SPECTRUM)
    sd = scenes.cornell_box(16, 16, samples=32, spectral=False, finalize=False)
    mine = {n: sd.materials[i] for n, i in sd.material_names.items()}
    for ref_name, my_name in (("leftwall", "leftWall"), ("rightwall", "rightWall"), ("floor", "floor")):
        if my_name not in mine:
            continue
        r, m = mats[cornell.material_index(ref_name)], mine[my_name][0]
        assert int(r["cls"]) == int(m["cls"]) == S.MAT_DIFFUSE
        assert int(r["two_sided"]) == int(m["two_sided"]) == 1
        assert float(r["opacity"]) == float(m["opacity"]) == 1.0 and float(r["normal_scale"]) == float(m["normal_scale"])
        assert np.array_equal(r["roughness"]["value"], m["roughness"]["value"])
        # Kd -> SpectralDistribution::rgb_reflectance: the loader's spectrum and the generator's are the same bytes
        rs_, ms_ = spectra[int(r["scattering"]["spectrum_index"])], sd.spectra[int(m["scattering"]["spectrum_index"])]
        assert bytes(rs_.tobytes()) == bytes(np.asarray(ms_).tobytes()), ref_name


def test_emission_spectrum_convention_matches_the_loader(oracle_mod):
    """`emitter color r g b` -> SpectralDistribution::rgb_luminance: the saved-scene variant of the asset (cornellbox.etx.materials) carries the
    light's colour as RGB; the generator's helper produces the loader's spectrum for the same numbers."""
    if not oracle_mod.ReferenceScene.available():
        pytest.skip("reference tree or oracle/_ref/libreference_loader.so not present")
    rs = oracle_mod.ReferenceScene("assets/cornellbox/cornellbox.etx.json")
    mats = _view(rs.scene["materials"], S.MATERIAL)
    spectra = _view(rs.scene["spectrums"], S.SPECTRUM)
    light = mats[rs.material_index("light")]
    ref = spectra[int(light["emission"]["spectrum_index"])]
    mine = np.asarray(scenes.spd_rgb_luminance([10.018112, 3.918244, 0.932069]))
    assert bytes(ref.tobytes()) == bytes(mine.tobytes())
    assert int(light["two_sided"]) == 1 and int(light["ext_medium"]) == 0  # `ext_medium fog__vol`
    rs.close()


def test_generator_camera_is_the_loaders_camera(cornell):
    sd = scenes.SceneData()
    sd.set_camera(CAMERA["origin"], CAMERA["target"], CAMERA["up"], 640, 640, CAMERA["fov"], clip_near=0.10000000149011612, clip_far=100.0)
    for field in ("view_proj", "position", "side", "up", "direction", "tan_half_fov", "aspect", "area", "image_plane", "film_size", "clip_near", "clip_far", "lens_radius",
                  "focal_distance"):
        a, b = np.asarray(sd.camera[field]), np.asarray(cornell.camera[field])
        # `position` comes out of inverse(view) in the loader: its x is -0.0 where the generator stores the origin's +0.0 (equal as values)
        assert (np.array_equal(a, b) if field == "position" else bit_equal(a, b)), field


def test_oracle_renders_the_reference_loaded_scene(cornell, oracle_mod):
    """The loader's PODs go straight into the oracle (the bytes etxb_upload_scene takes): fog volume, Boundary mesh of 137k triangles, sun +
    sky emitters, conductor box — two iterations at a 40x40 film."""
    cornell.resize(40, 40, CAMERA["origin"], CAMERA["target"], CAMERA["up"], CAMERA["fov"])
    try:
        o = oracle_mod.Oracle(cornell, "native")
        o.begin(0)
        o.run(2, threads=4)
        img = o.film(S.FILM_RESULT)[..., :3]
        assert img.shape == (40, 40, 3) and np.isfinite(img).all() and img.min() >= 0.0 and img.mean() > 0.01
        c = o.counters()
        assert int(c["rays_closest"][0]) > 3000 and int(c["light_vertices"][0]) > 0
        o.close()
    finally:
        cornell.resize(640, 640, CAMERA["origin"], CAMERA["target"], CAMERA["up"], CAMERA["fov"])


def test_named_ior_convention_matches_the_loader(cornell):
    """`int_ior silver` -> the IOR database's eta / k spectra: the generator's tables (tools/make_data.py) hold the loader's bytes."""
    mats = _view(cornell.scene["materials"], S.MATERIAL)
    spectra = _view(cornell.scene["spectrums"], S.SPECTRUM)
    ref = mats[cornell.material_index("tallbox")]
    sd = scenes.SceneData()
    mine = sd.materials[sd.add_material("metal", cls=S.MAT_CONDUCTOR, ks=[1.0, 1.0, 1.0], roughness=0.0, int_ior="silver")][0]
    assert int(ref["cls"]) == int(mine["cls"]) == S.MAT_CONDUCTOR
    assert int(ref["int_ior"]["cls"]) == int(mine["int_ior"]["cls"])
    for part in ("eta_index", "k_index"):
        a, b = spectra[int(ref["int_ior"][part])], np.asarray(sd.spectra[int(mine["int_ior"][part])])
        assert bytes(a.tobytes()) == bytes(b.tobytes()), part


def test_committed_pod_dump_is_what_the_loader_builds(oracle_mod):
    """tests/golden/ref_cornell_40.npz (tools/dump_reference_scene.py) is the loader's Scene / Camera byte for byte: the oracle renders the
    dump and the live PODs to the same bits (film, sampler states).  The GPU box has only the dump."""
    import os
    from conftest import GOLDEN
    from etx_tracer_b200 import pod_io
    if not oracle_mod.ReferenceScene.available():
        pytest.skip("reference tree or oracle/_ref/libreference_loader.so not present")
    rs = oracle_mod.ReferenceScene("assets/cornellbox/cornellbox.json")
    rs.resize(40, 40, CAMERA["origin"], CAMERA["target"], CAMERA["up"], CAMERA["fov"])
    sd = pod_io.load(os.path.join(GOLDEN, "ref_cornell_40.npz"))
    assert bytes(sd.camera.tobytes()) == bytes(rs.camera.tobytes())
    out = []
    for scene in (sd, rs):
        o = oracle_mod.Oracle(scene)
        o.begin(0)
        o.run(1, threads=1)
        out.append((o.film(S.FILM_CAMERA).copy(), o.film(S.FILM_LIGHT).copy(), o.buffer(S.BUF_CAMERA_SAMPLER, np.uint32), o.buffer(S.BUF_LIGHT_SAMPLER, np.uint32)))
        o.close()
    for x, y in zip(*out):
        assert bit_equal(x, y)
    rs.close()
