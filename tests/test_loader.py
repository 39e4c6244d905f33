"""The product-side scene loader (etx_tracer_b200/loader.py) against the reference's OWN loader (scene_representation.cxx and
friends compiled in place into oracle/_ref/libreference_loader.so — test infrastructure): the same scene FILES read by both, the Scene / Camera
PODs compared array by array.  CPU only; skipped where the reference tree is absent.

Byte-identical: triangles (indices, material, geometric normal), vertex positions / normals / texture coordinates, tangent frames of meshes without
texture coordinates, every Material record, emitter profiles / instances / the emitter distribution, media, images (pixels, options, sampling
tables), the scene scalars, the camera (up to the sign of a zero in `position`).  Stated differences: tangent frames of meshes WITH texture
coordinates (per-triangle UV derivatives here, MikkTSpace there), black-body spectra to 1e-6 relative (glibc expf against numpy's float32 exp),
padding bytes the reference leaves uninitialised."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from etx_tracer_b200 import api, loader, structs as S


@pytest.fixture(scope="module")
def ref(oracle_mod):
    if not oracle_mod.ReferenceScene.available():
        pytest.skip("reference tree or oracle/_ref/libreference_loader.so not present")
    return oracle_mod.ReferenceScene


def _view(av, dt):
    n = int(np.asarray(av["count"]).reshape(-1)[0])
    return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(int(np.asarray(av["a"]).reshape(-1)[0])), dtype=dt) if n else np.zeros(0, dt)


def _diff(name, a, b, skip=(), out=None):
    """field-by-field byte comparison of two structured arrays; returns the list of differing leaf fields"""
    out = [] if out is None else out
    if a.shape != b.shape:
        out.append(f"{name}: shape {a.shape} vs {b.shape}")
        return out
    for f in a.dtype.names:
        if f in skip:
            continue
        x, y = a[f], b[f]
        if x.dtype.names:
            _diff(f"{name}.{f}", x, y, skip, out)
        elif not np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8)):
            out.append(f"{name}.{f}")
    return out


def compare_scenes(rs, sd, uv_tangents_exact=True):
    rsc, msc = rs.scene, sd.scene
    problems = []
    va, vb = _view(rsc["vertices"], S.VERTEX), _view(msc["vertices"], S.VERTEX)
    problems += _diff("vertices", va, vb, skip=() if uv_tangents_exact else ("tan", "btn"))
    if not uv_tangents_exact and va.shape == vb.shape:
        # both frames are orthonormal around the same normal
        for arr in (va, vb):
            assert np.abs((arr["tan"] * arr["nrm"]).sum(axis=1)).max() < 1e-3 and np.abs(np.linalg.norm(arr["tan"], axis=1) - 1.0).max() < 1e-3
    problems += _diff("triangles", _view(rsc["triangles"], S.TRIANGLE), _view(msc["triangles"], S.TRIANGLE))
    problems += _diff("materials", _view(rsc["materials"], S.MATERIAL), _view(msc["materials"], S.MATERIAL))
    problems += _diff("emitter_profiles", _view(rsc["emitter_profiles"], S.EMITTER_PROFILE), _view(msc["emitter_profiles"], S.EMITTER_PROFILE), skip=("pad",))
    problems += _diff("emitter_instances", _view(rsc["emitter_instances"], S.EMITTER), _view(msc["emitter_instances"], S.EMITTER))
    problems += _diff("mediums", _view(rsc["mediums"], S.MEDIUM), _view(msc["mediums"], S.MEDIUM), skip=("density",))
    if not np.array_equal(_view(rsc["triangle_to_emitter"], np.dtype(np.uint32)), _view(msc["triangle_to_emitter"], np.dtype(np.uint32))):
        problems.append("triangle_to_emitter")
    sa, sb = _view(rsc["spectrums"], S.SPECTRUM), _view(msc["spectrums"], S.SPECTRUM)
    if sa.shape != sb.shape:
        problems.append(f"spectrums: {sa.shape} vs {sb.shape}")
    else:
        pa, pb = sa["entries"]["power"], sb["entries"]["power"]
        rel = np.abs(pa - pb).max(axis=1) / np.maximum(np.abs(pa).max(axis=1), 1e-30)
        reli = np.abs(sa["integrated"] - sb["integrated"]).max(axis=1) / np.maximum(np.abs(sa["integrated"]).max(axis=1), 1e-30)
        if rel.max() > 2e-6 or reli.max() > 5e-6:
            problems.append(f"spectrums: power {rel.max():.2e} (index {int(rel.argmax())}), integrated {reli.max():.2e} (index {int(reli.argmax())})")
    ea, eb = _view(rsc["emitters_distribution"]["values"], S.DIST_ENTRY), _view(msc["emitters_distribution"]["values"], S.DIST_ENTRY)
    n = int(rsc["emitter_instances"]["count"][0]) + 1
    if ea.shape[0] < n - 1 or eb.shape[0] < n - 1 or not np.allclose(ea["pdf"][:n - 1], eb["pdf"][:n - 1], rtol=2e-6) or not np.allclose(ea["cdf"][:n - 1], eb["cdf"][:n - 1], rtol=2e-6, atol=1e-7):
        problems.append("emitters_distribution")
    for f in S.SCENE.names:
        if rsc[f].dtype.names or f == "pad":
            continue
        if f == "bounding_sphere_radius" or f == "bounding_sphere_center":
            ok = np.allclose(rsc[f], msc[f], rtol=1e-6)
        else:
            ok = np.array_equal(np.ascontiguousarray(rsc[f]).view(np.uint8), np.ascontiguousarray(msc[f]).view(np.uint8))
        if not ok:
            problems.append(f"scene.{f}: {rsc[f]} vs {msc[f]}")
    for f in S.CAMERA.names:
        a, b = np.asarray(rs.camera[f]), np.asarray(sd.camera[f])
        # `position` comes out of inverse(view) in the reference (build_camera :582-585): the origin up to a rounding step and the sign of a zero
        if not (np.allclose(a, b, rtol=2e-7, atol=1e-7) if f == "position" else np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))):
            problems.append(f"camera.{f}: {a} vs {b}")
    ia, ib = _view(rsc["images"], S.IMAGE), _view(msc["images"], S.IMAGE)
    if ia.shape != ib.shape:
        problems.append(f"images: {ia.shape} vs {ib.shape}")
    else:
        for k in range(len(ia)):
            for f in ("fsize", "isize", "offset", "scale", "options", "format", "data_size"):
                if not np.array_equal(ia[k][f], ib[k][f]):
                    problems.append(f"image {k}.{f}: {ia[k][f]} vs {ib[k][f]}")
            nb = int(ia[k]["data_size"])
            if nb == int(ib[k]["data_size"]):
                pa = bytes((C.c_char * nb).from_address(int(ia[k]["pixels"]["a"])))
                pb = bytes((C.c_char * nb).from_address(int(ib[k]["pixels"]["a"])))
                if pa != pb:
                    # 8-bit images byte for byte; float images (the Blackman-Harris pixel filter is COMPUTED by both sides: cosf against numpy's cos) to 1e-6
                    fa, fb = np.frombuffer(pa, np.float32), np.frombuffer(pb, np.float32)
                    if int(ia[k]["format"]) != 1 or not np.allclose(fa, fb, rtol=2e-6, atol=2e-7):
                        problems.append(f"image {k}: pixels differ")
            if not np.isclose(float(ia[k]["normalization"]), float(ib[k]["normalization"]), rtol=1e-5):
                problems.append(f"image {k}.normalization: {ia[k]['normalization']} vs {ib[k]['normalization']}")
            if int(ia[k]["options"]) & 1:  # sampling table
                ya, yb = _view(ia[k]["y_distribution"]["values"], S.DIST_ENTRY), _view(ib[k]["y_distribution"]["values"], S.DIST_ENTRY)
                h = int(ia[k]["isize"][1])
                if not np.allclose(ya["pdf"][:h], yb["pdf"][:h], rtol=1e-5, atol=1e-9) or not np.allclose(ya["cdf"][:h], yb["cdf"][:h], rtol=1e-5, atol=1e-7):
                    problems.append(f"image {k}: y distribution")
    return problems


def test_shipped_cornell_asset_matches_the_reference_loader(ref):
    """bin/assets/cornellbox/cornellbox.json: 138 318 triangles (quads split along the shorter diagonal like tinyobjloader does), fog medium behind a
    Boundary mesh with its running bounding box, constant environment + sun, black-body emitters, `int_ior silver`, camera from focal-length."""
    rs = ref("assets/cornellbox/cornellbox.json")
    sd = loader.load_scene(os.path.join(os.environ.get("ETX_REFERENCE", "/root/reference"), "bin", "assets", "cornellbox", "cornellbox.json"))
    assert sd.triangle_count == 138318 and (sd.width, sd.height) == (640, 640)
    problems = compare_scenes(rs, sd, uv_tangents_exact=False)  # the asset has texture coordinates: MikkTSpace frames in the reference
    assert not problems, problems
    rs.close()


def test_saved_scene_variant_matches_the_reference_loader(ref):
    """cornellbox.etx.json + cornellbox.etx.materials: what the application writes back (RGB emitter colours, et::camera block, medium ids)."""
    rs = ref("assets/cornellbox/cornellbox.etx.json")
    sd = loader.load_scene(os.path.join(os.environ.get("ETX_REFERENCE", "/root/reference"), "bin", "assets", "cornellbox", "cornellbox.etx.json"))
    problems = compare_scenes(rs, sd, uv_tangents_exact=False)
    assert not problems, problems
    rs.close()


OBJ = """# a room without texture coordinates (tangent frames: orthonormal_basis, exact), quads, a polygon-free mix, a degenerate triangle
mtllib ignored.mtl
v -1 0 -1
v 1 0 -1
v 1 0 1
v -1 0 1
v -1 2 -1
v 1 2 -1
v 1 2 1
v -1 2 1
v -0.3 1.99 -0.3
v 0.3 1.99 -0.3
v 0.3 1.99 0.3
v -0.3 1.99 0.3
v -0.5 0 -0.2
v 0.1 0 -0.5
v 0.4 0 0.3
v -0.2 0.9 0.0
vn 0 1 0
o floor
usemtl Floor
f 1//1 4//1 3//1 2//1
o walls
usemtl LeftWall
f 1 5 8 4
usemtl RightWall
f 2 3 7 6
usemtl Back
f 1 2 6 5
usemtl Ceiling
f 5 6 7 8
o lamp
usemtl Lamp
f 9 10 11 12
g props
usemtl Gold
f 13 14 16
f 14 15 16
usemtl Glass
f 15 13 16
usemtl Film
f 13 15 14
usemtl Missing
f 1 2 3
usemtl Gold
f 13 13 14
o volume
usemtl FogShell
f -16 -15 -10
f -14 -13 -12
"""

MTL = """newmtl et::spectrum
id warm
illuminant
rgb 1.0 0.6 0.2
scale 3.0

newmtl et::spectrum
id tint
rgb 0.2 0.5 0.9

newmtl et::spectrum
id lamp_bb
blackbody 3200
scale 0.00001

newmtl et::spectrum
id measured
samples 400 0.1 500 0.8 600 0.4 700 0.2
normalize luminance

newmtl et::medium
id haze
absorption 0.02 0.03 0.05
scattering 0.4
g 0.3

newmtl et::medium
id sealed
scattering 0.1 0.2 0.3
enclosed

newmtl et::dir
direction 0.2 1.0 0.3
color nblackbody 5800 scale 2.0
angular_diameter 1.5

newmtl et::env
color 0.1 0.2 0.4
rotation 45
scale 2.0

newmtl et::camera
id second
viewport 96 64
origin 0 1 3.5
target 0 1 0
fov 42
lens-radius 0.05
focal-distance 3.0
clip-near 0.05

newmtl et::camera
id main
active 1
viewport 80 60
origin 0.2 1.1 3.2
target 0 0.9 0
up 0 1 0
focal-length 35
ext_medium haze

newmtl Floor
material class diffuse
Kd 0.8 0.8 0.8
diffuse 1
Pr 0.4

newmtl LeftWall
Kd tint
two_sided true

newmtl RightWall
material class plastic
Kd 0.1 0.7 0.2
Ks 1 1 1
Pr 0.3 0.1
int_ior 1.49
two_sided 1

newmtl Back
base rightwall
material class velvet
Kd 0.6 0.1 0.1

newmtl Ceiling
material class translucent
Kt 0.5 0.5 0.5
opacity 0.75

newmtl Lamp
Kd 0 0 0
emitter color 4 3 2 scale 2.5 collimated 0.4 twosided

newmtl Gold
material class conductor
int_ior gold
Ks 1 0.9 0.8
Pr 0.2
thinfilm range 200 600 ior 1.33

newmtl Glass
material class dielectric
int_ior glass
ext_ior 1.0003
Kt 1 1 1
int_medium sealed
subsurface path refracted distances 0.5 0.3 0.1 scale 0.2 class approximate

newmtl Film
material class thinfilm
thinfilm range 300 700 ior water
Ke warm
metalness 0.3
transmission 0.6

newmtl FogShell
material class boundary
int_medium haze
"""


def _write_scene(tmp_path, obj=OBJ, mtl=MTL, js=None):
    (tmp_path / "room.obj").write_text(obj)
    (tmp_path / "room.mtl").write_text(mtl)
    d = {"geometry": "room.obj", "materials": "room.mtl", "samples": 24, "max-path-length": 12, "min-path-length": 2, "random-termination-start": 4, "spectral": True,
         "force-tangents": False}
    d.update(js or {})
    (tmp_path / "room.json").write_text(json.dumps(d))
    return str(tmp_path / "room.json")


def test_directive_coverage_scene_matches_the_reference_loader(ref, tmp_path):
    """A generated scene that walks through the dialect: et::spectrum (rgb / illuminant / blackbody / samples + normalize), et::medium (absorption,
    scalar scattering, g, enclosed), et::dir with a disk, et::env colour + rotation + scale, two et::camera blocks (the active one wins; focal-length,
    ext_medium), Kd by spectrum name, base inheritance, two-valued Pr, numeric / named / default IORs, thin film, subsurface, emitter keywords, Ke,
    metalness / transmission, opacity, an undeclared material (faces dropped), a degenerate triangle, negative indices, medium bounds per shape."""
    path = _write_scene(tmp_path)
    rs = ref(path)
    sd = loader.load_scene(path)
    assert sd.triangle_count == int(rs.scene["triangles"]["count"][0]) == 19
    problems = compare_scenes(rs, sd, uv_tangents_exact=True)
    assert not problems, problems
    assert (sd.width, sd.height) == (80, 60) and int(sd.scene["max_path_length"][0]) == 12 and int(sd.scene["flags"][0]) & S.SCENE_SPECTRAL
    rs.close()


def test_json_camera_and_obj_only_entry_points(ref, tmp_path):
    """The camera block of the .json (no et::camera in the materials), and an .obj given directly (mtllib, default settings)."""
    mtl = "\n".join(b for b in MTL.split("\n\n") if not b.startswith("newmtl et::camera")) + "\n"
    cam = {"class": "perspective", "viewport": [72, 48], "origin": [0.0, 1.0, 3.0], "target": [0.0, 1.0, 0.0], "up": [0.0, 1.0, 0.0], "fov": 50.0, "lens-radius": 0.02,
           "focal-distance": 2.5, "clip-near": 0.1, "clip-far": 50.0}
    path = _write_scene(tmp_path, mtl=mtl, js={"camera": cam})
    rs = ref(path)
    sd = loader.load_scene(path)
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()
    (tmp_path / "direct.obj").write_text(OBJ.replace("mtllib ignored.mtl", "mtllib room.mtl"))
    rs = ref(str(tmp_path / "direct.obj"))
    sd = loader.load_scene(str(tmp_path / "direct.obj"))
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()


def test_textures_match_the_reference_loader(ref, tmp_path):
    """8-bit PNG textures stay RGBA8 with the sRGB curve removed and re-quantised (normal maps skip the conversion), a float EXR environment map gets
    its importance-sampling table, an emission image too; files written with the module's own writers, read there by stb_image / tinyexr."""
    rng = np.random.default_rng(5)
    albedo = rng.integers(0, 256, (8, 16, 4), dtype=np.uint8)
    albedo[..., 3] = 255
    albedo[2, 3, 3] = 100  # an alpha hole: HasAlphaChannel
    api.write_png(str(tmp_path / "albedo.png"), albedo)
    nm = rng.integers(100, 156, (4, 4, 4), dtype=np.uint8)
    nm[..., 2] = 250
    nm[..., 3] = 255
    api.write_png(str(tmp_path / "normal.png"), nm)
    env = (rng.random((16, 32, 4)) * 2.0).astype(np.float32)
    env[..., 3] = 1.0
    env[3, 7, :3] = 40.0
    api.write_exr(str(tmp_path / "sky.exr"), env)
    glow = rng.integers(0, 256, (4, 8, 4), dtype=np.uint8)
    glow[..., 3] = 255
    api.write_png(str(tmp_path / "glow.png"), glow)
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    mtl = MTL.replace("newmtl et::env\ncolor 0.1 0.2 0.4", "newmtl et::env\nimage sky.exr\ncolor 0.1 0.2 0.4")
    mtl = mtl.replace("newmtl Floor\n", "newmtl Floor\nmap_Kd albedo.png\nnormalmap image normal.png scale 0.5\n")
    mtl = mtl.replace("collimated 0.4 twosided", "collimated 0.4 twosided image glow.png")  # `image` last: the reference loses what follows it on the line
    path = _write_scene(tmp_path, obj=obj, mtl=mtl)
    rs = ref(path)
    sd = loader.load_scene(path)
    assert int(sd.scene["images"]["count"][0]) == int(rs.scene["images"]["count"][0]) == 5  # sky, albedo, normal map, glow, pixel filter
    problems = compare_scenes(rs, sd, uv_tangents_exact=False)
    assert not problems, problems
    rs.close()


def _write_hdr(path, img, rle):
    """float RGB -> Radiance RGBE (the classic float2rgbe), flat or with every scan line run-length framed (literal runs and one repeat run per plane)"""
    h, w = img.shape[:2]
    v = img[..., :3].max(axis=-1)
    m, e = np.frexp(v)
    scale = np.where(v < 1e-32, 0.0, m * 256.0 / np.maximum(v, 1e-38))
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[..., :3] = (img[..., :3] * scale[..., None]).astype(np.uint8)
    rgbe[..., 3] = np.where(v < 1e-32, 0, e + 128).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode())
        if not rle:
            f.write(rgbe.tobytes())
            return
        for y in range(h):
            f.write(bytes([2, 2, w >> 8, w & 255]))
            for c in range(4):
                row = rgbe[y, :, c]
                if (row == row[0]).all():
                    x = 0
                    while x < w:
                        n = min(127, w - x)
                        f.write(bytes([128 + n, int(row[0])]))
                        x += n
                else:
                    x = 0
                    while x < w:
                        n = min(128, w - x)
                        f.write(bytes([n]) + row[x:x + n].tobytes())
                        x += n


@pytest.mark.parametrize("rle", [False, True])
def test_hdr_pfm_images_and_parametric_media_match_the_reference_loader(ref, tmp_path, rle):
    """A Radiance .hdr environment map (flat and run-length encoded scan lines, read there by stb_image), a .pfm texture in the reference's own header
    variant, and an et::medium given as `parametric color … distances … scale …` (subsurface::remap)."""
    rng = np.random.default_rng(11)
    env = (rng.random((12, 24, 3)) * 3.0).astype(np.float32)
    env[5, 5] = (30.0, 20.0, 10.0)
    env[:, 20:] = 0.5  # constant columns: one repeat run per plane when run-length encoded
    _write_hdr(str(tmp_path / "sky.hdr"), env, rle)
    tex = rng.random((6, 10, 3)).astype(np.float32)
    with open(tmp_path / "paint.pfm", "wb") as f:
        f.write(b"PF\n10\n6\n-1.0\n" + tex.tobytes())
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    mtl = MTL.replace("newmtl et::env\ncolor 0.1 0.2 0.4", "newmtl et::env\nimage sky.hdr\ncolor 0.1 0.2 0.4")
    mtl = mtl.replace("newmtl Floor\n", "newmtl Floor\nmap_Kd paint.pfm\n")
    mtl = mtl.replace("id sealed\nscattering 0.1 0.2 0.3", "id sealed\nparametric color 0.8 0.5 0.3 distances 0.4 0.2 0.1 scale 0.5")
    path = _write_scene(tmp_path, obj=obj, mtl=mtl)
    rs = ref(path)
    sd = loader.load_scene(path)
    problems = compare_scenes(rs, sd, uv_tangents_exact=False)
    assert not problems, problems
    rs.close()


def test_loader_refuses_what_it_does_not_read(tmp_path):
    path = _write_scene(tmp_path, mtl=MTL + "\nnewmtl et::atmosphere\nquality 0.1\n")
    with pytest.raises(loader.LoaderError):
        loader.load_scene(path)
    path = _write_scene(tmp_path, mtl=MTL.replace("scattering 0.4", "scattering 0.4\nvolume cloud.nvdb"))
    with pytest.raises(loader.LoaderError):
        loader.load_scene(path)
    (tmp_path / "x.json").write_text(json.dumps({"geometry": "scene.gltf"}))
    with pytest.raises(loader.LoaderError):
        loader.load_scene(str(tmp_path / "x.json"))


def test_loaded_scene_renders_through_the_oracle(ref, oracle_mod, tmp_path):
    """The loader's PODs are what etxb_upload_scene and the oracle consume: two VCM iterations of the generated scene, finite and lit."""
    sd = loader.load_scene(_write_scene(tmp_path))
    o = oracle_mod.Oracle(sd, "native" if oracle_mod.available("native") else "parity")
    o.begin(0)
    o.run(2, threads=2)
    img = o.film(S.FILM_RESULT)[..., :3]
    assert np.isfinite(img).all() and img.mean() > 1e-3
    o.close()


@pytest.mark.gpu
def test_loaded_scene_files_render_bit_exact_on_the_device(oracle_mod, tmp_path):
    """Scene FILE -> loader -> etxb_upload_scene: the parity build against the oracle on the loader's PODs, both integrators (the generated
    directive-coverage scene: media, Boundary shell, thin film, subsurface glass, distant emitters with a disk, thin-lens camera inside a medium)."""
    from conftest import bit_equal
    sd = loader.load_scene(_write_scene(tmp_path))
    o = oracle_mod.Oracle(sd)
    o.begin(0)
    o.run(2, threads=1)
    g = api.GPUVCM(sd, flavor="parity")
    g.render(2)
    for bid, dt in ((S.BUF_LIGHT_SAMPLER, np.uint32), (S.BUF_LV_POS, np.float32), (S.BUF_CAMERA_SAMPLER, np.uint32), (S.BUF_CAMERA_GATHERED, np.float32)):
        assert bit_equal(g.buffer(bid, dt), o.buffer(bid, dt)), f"buffer {bid}"
    assert bit_equal(g.film(S.FILM_CAMERA)[..., :3], o.film(S.FILM_CAMERA)[..., :3])
    g.close()
    o.set_integrator(S.INTEGRATOR_PT)
    o.pt_set_options(S.default_pt_options())
    o.begin(0)
    o.run(3, threads=1)
    p = api.GPUPathTracing(sd, flavor="parity")
    p.render(3)
    assert bit_equal(p.buffer(S.BUF_CAMERA_SAMPLER, np.uint32), o.buffer(S.BUF_CAMERA_SAMPLER, np.uint32))
    for layer in (S.FILM_CAMERA, S.FILM_NORMALS, S.FILM_ALBEDO):
        assert bit_equal(p.film(layer)[..., :3], o.film(layer)[..., :3]), layer
    p.close()
    o.close()


@pytest.mark.gpu
def test_headless_render_command(tmp_path):
    """python -m etx_tracer_b200.render: scene file in, OpenEXR / PNG out (loader + integrator pump + film export), both integrators."""
    from etx_tracer_b200 import render
    path = _write_scene(tmp_path)
    out = str(tmp_path / "vcm.exr")
    assert render.main([path, "-o", out, "--spp", "4", "--option", "vcm-merging=0"]) == 0
    d = open(out, "rb").read()
    assert d[:4] == b"\x76\x2f\x31\x01" and len(d) > 80 * 60 * 16
    out = str(tmp_path / "pt.png")
    assert render.main([path, "-o", out, "--integrator", "pt", "--spp", "6", "--layer", "camera", "--exposure", "2.0", "--option", "bn=0"]) == 0
    assert open(out, "rb").read()[:8] == b"\x89PNG\r\n\x1a\n"
