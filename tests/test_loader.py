"""The product-side scene loaders (the module's C++ loader csrc/scene_loader.cpp behind etxb_scene_file_load, and its Python twin
etx_tracer_b200/loader.py) against the reference's OWN loader (scene_representation.cxx and
friends compiled in place into oracle/_ref/libreference_loader.so — test infrastructure): the same scene FILES read by both, the Scene / Camera
PODs compared array by array.  CPU only; skipped where the reference tree is absent.

Byte-identical: triangles (indices, material, geometric normal), vertex positions / normals / texture coordinates, tangent frames (incl. the 414 954
vertices of the Cornell asset), every Material record, emitter profiles / instances / the emitter distribution, media, images (pixels, options, sampling
tables), the scene scalars, the camera (up to the sign of a zero in `position`).  Stated differences: the Python twin's black-body spectra to 1e-6
relative (glibc expf against numpy's float32 exp); padding bytes the reference leaves uninitialised."""
import ctypes as C
import json
import os

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")  # the PIZ test writes its file with OpenCV's OpenEXR encoder

import numpy as np
import pytest

from etx_tracer_b200 import api, loader, structs as S


@pytest.fixture(scope="module")
def ref(oracle_mod):
    if not oracle_mod.ReferenceScene.available():
        pytest.skip("reference tree or oracle/_ref/libreference_loader.so not present")
    return oracle_mod.ReferenceScene


def _load_cpp(path):
    return api.SceneFile(path, flavor="parity")


LOADERS = {"cpp": (_load_cpp, api.EtxbError), "python": (loader.load_scene, loader.LoaderError)}


@pytest.fixture(params=["cpp", "python"])
def load(request):
    """the loader under test: the C ABI's etxb_scene_file_load (host C++ of the module) and the Python twin, held to the same comparisons"""
    fn = LOADERS[request.param][0]

    def run(path):
        return fn(path)
    # meshes with texture coordinates: both loaders run the module's restatement of the tangent-space generator the reference calls (bit-identical frames)
    run.exact_tangents = True
    return run


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


def compare_scenes(rs, sd, uv_tangents_exact=True, generated_sky=False):
    """generated_sky: the scene carries a procedural sky image, whose `average of the upper hemisphere` term the reference itself sums in a run-dependent
    order (atomics, scattering.cxx:287-322) — float images, their tables and the emitter weights derived from them are then held to 1e-4"""
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
    ea_, eb_ = _view(rsc["emitter_instances"], S.EMITTER), _view(msc["emitter_instances"], S.EMITTER)
    problems += _diff("emitter_instances", ea_, eb_, skip=("spectrum_weight", "additional_weight") if generated_sky else ())
    if generated_sky and ea_.shape == eb_.shape:
        for f in ("spectrum_weight", "additional_weight"):
            if not np.allclose(ea_[f], eb_[f], rtol=1e-5):
                problems.append(f"emitter_instances.{f}")
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
                    if int(ia[k]["format"]) != 1 or not np.allclose(fa, fb, rtol=5e-5 if generated_sky else 2e-6, atol=2e-7):
                        problems.append(f"image {k}: pixels differ")
            if not np.isclose(float(ia[k]["normalization"]), float(ib[k]["normalization"]), rtol=1e-4 if generated_sky else 1e-5):
                problems.append(f"image {k}.normalization: {ia[k]['normalization']} vs {ib[k]['normalization']}")
            if int(ia[k]["options"]) & 1:  # sampling table
                ya, yb = _view(ia[k]["y_distribution"]["values"], S.DIST_ENTRY), _view(ib[k]["y_distribution"]["values"], S.DIST_ENTRY)
                h = int(ia[k]["isize"][1])
                tol = 1e-4 if generated_sky else 1e-5
                if not np.allclose(ya["pdf"][:h], yb["pdf"][:h], rtol=tol, atol=1e-9) or not np.allclose(ya["cdf"][:h], yb["cdf"][:h], rtol=tol, atol=1e-7):
                    problems.append(f"image {k}: y distribution")
    return problems


def test_shipped_cornell_asset_matches_the_reference_loader(ref, load):
    """bin/assets/cornellbox/cornellbox.json: 138 318 triangles (quads split along the shorter diagonal like tinyobjloader does), fog medium behind a
    Boundary mesh with its running bounding box, constant environment + sun, black-body emitters, `int_ior silver`, camera from focal-length."""
    rs = ref("assets/cornellbox/cornellbox.json")
    sd = load(os.path.join(os.environ.get("ETX_REFERENCE", "/root/reference"), "bin", "assets", "cornellbox", "cornellbox.json"))
    assert sd.triangle_count == 138318 and (sd.width, sd.height) == (640, 640)
    problems = compare_scenes(rs, sd, uv_tangents_exact=load.exact_tangents)  # the asset has texture coordinates: MikkTSpace frames in the reference
    assert not problems, problems
    rs.close()


def test_saved_scene_variant_matches_the_reference_loader(ref, load):
    """cornellbox.etx.json + cornellbox.etx.materials: what the application writes back (RGB emitter colours, et::camera block, medium ids)."""
    rs = ref("assets/cornellbox/cornellbox.etx.json")
    sd = load(os.path.join(os.environ.get("ETX_REFERENCE", "/root/reference"), "bin", "assets", "cornellbox", "cornellbox.etx.json"))
    problems = compare_scenes(rs, sd, uv_tangents_exact=load.exact_tangents)
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


def test_directive_coverage_scene_matches_the_reference_loader(ref, tmp_path, load):
    """A generated scene that walks through the dialect: et::spectrum (rgb / illuminant / blackbody / samples + normalize), et::medium (absorption,
    scalar scattering, g, enclosed), et::dir with a disk, et::env colour + rotation + scale, two et::camera blocks (the active one wins; focal-length,
    ext_medium), Kd by spectrum name, base inheritance, two-valued Pr, numeric / named / default IORs, thin film, subsurface, emitter keywords, Ke,
    metalness / transmission, opacity, an undeclared material (faces dropped), a degenerate triangle, negative indices, medium bounds per shape."""
    path = _write_scene(tmp_path)
    rs = ref(path)
    sd = load(path)
    assert sd.triangle_count == int(rs.scene["triangles"]["count"][0]) == 19
    problems = compare_scenes(rs, sd, uv_tangents_exact=True)
    assert not problems, problems
    assert (sd.width, sd.height) == (80, 60) and int(sd.scene["max_path_length"][0]) == 12 and int(sd.scene["flags"][0]) & S.SCENE_SPECTRAL
    rs.close()


def test_json_camera_and_obj_only_entry_points(ref, tmp_path, load):
    """The camera block of the .json (no et::camera in the materials), and an .obj given directly (mtllib, default settings)."""
    mtl = "\n".join(b for b in MTL.split("\n\n") if not b.startswith("newmtl et::camera")) + "\n"
    cam = {"class": "perspective", "viewport": [72, 48], "origin": [0.0, 1.0, 3.0], "target": [0.0, 1.0, 0.0], "up": [0.0, 1.0, 0.0], "fov": 50.0, "lens-radius": 0.02,
           "focal-distance": 2.5, "clip-near": 0.1, "clip-far": 50.0}
    path = _write_scene(tmp_path, mtl=mtl, js={"camera": cam})
    rs = ref(path)
    sd = load(path)
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()
    (tmp_path / "direct.obj").write_text(OBJ.replace("mtllib ignored.mtl", "mtllib room.mtl"))
    rs = ref(str(tmp_path / "direct.obj"))
    sd = load(str(tmp_path / "direct.obj"))
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()


def test_textures_match_the_reference_loader(ref, tmp_path, load):
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
    sd = load(path)
    assert int(sd.scene["images"]["count"][0]) == int(rs.scene["images"]["count"][0]) == 5  # sky, albedo, normal map, glow, pixel filter
    problems = compare_scenes(rs, sd, uv_tangents_exact=load.exact_tangents)
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
def test_hdr_pfm_images_and_parametric_media_match_the_reference_loader(ref, tmp_path, rle, load):
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
    sd = load(path)
    problems = compare_scenes(rs, sd, uv_tangents_exact=load.exact_tangents)
    assert not problems, problems
    rs.close()


@pytest.mark.parametrize("which", ["cpp", "python"])
def test_loader_refuses_what_it_does_not_read(tmp_path, which):
    fn, error = LOADERS[which]
    (tmp_path / "x.json").write_text(json.dumps({"geometry": "scene.gltf"}))
    with pytest.raises(error, match="glTF"):
        fn(str(tmp_path / "x.json"))
    with pytest.raises(error):
        fn(str(tmp_path / "missing.json"))


def test_cpp_loader_matches_the_python_loader(tmp_path):
    """Runs where the reference tree is absent too: the C++ loader and the Python loader read the generated directive-coverage scene (with an 8-bit
    texture, a float environment map and an emission image) into the same PODs."""
    rng = np.random.default_rng(7)
    albedo = rng.integers(0, 256, (8, 16, 4), dtype=np.uint8)
    albedo[..., 3] = 255
    api.write_png(str(tmp_path / "albedo.png"), albedo)
    env = (rng.random((16, 32, 4)) * 2.0).astype(np.float32)
    env[..., 3] = 1.0
    api.write_exr(str(tmp_path / "sky.exr"), env)
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    mtl = MTL.replace("newmtl et::env\ncolor 0.1 0.2 0.4", "newmtl et::env\nimage sky.exr\ncolor 0.1 0.2 0.4")
    mtl = mtl.replace("newmtl Floor\n", "newmtl Floor\nmap_Kd albedo.png\n")
    mtl = mtl.replace("collimated 0.4 twosided", "collimated 0.4 twosided image albedo.png")
    path = _write_scene(tmp_path, obj=obj, mtl=mtl)
    a, b = loader.load_scene(path), _load_cpp(path)
    problems = compare_scenes(a, b, uv_tangents_exact=True)
    assert not problems, problems
    assert sorted(b.material_names) == sorted(a.material_names)
    b.close()


def test_loaded_scene_renders_through_the_oracle(ref, oracle_mod, tmp_path, load):
    """The loader's PODs are what etxb_upload_scene and the oracle consume: two VCM iterations of the generated scene, finite and lit."""
    sd = load(_write_scene(tmp_path))
    o = oracle_mod.Oracle(sd, "native" if oracle_mod.available("native") else "parity")
    o.begin(0)
    o.run(2, threads=2)
    img = o.film(S.FILM_RESULT)[..., :3]
    assert np.isfinite(img).all() and img.mean() > 1e-3
    o.close()


@pytest.mark.gpu
def test_loaded_scene_files_render_bit_exact_on_the_device(oracle_mod, tmp_path, load):
    """Scene FILE -> loader -> etxb_upload_scene: the parity build against the oracle on the loader's PODs, both integrators (the generated
    directive-coverage scene: media, Boundary shell, thin film, subsurface glass, distant emitters with a disk, thin-lens camera inside a medium)."""
    from conftest import bit_equal
    sd = load(_write_scene(tmp_path))
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


def _png_bytes(img, ctype, filters, level, palette=None, idat_split=0):
    """A PNG the way an ordinary encoder writes one: zlib-compressed IDAT (dynamic / fixed Huffman blocks), per-row filter types 0-4"""
    import struct
    import zlib
    h, w = img.shape[:2]
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    rows = img.reshape(h, w * ch).astype(np.int32)
    raw = bytearray()
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        ft = filters[y % len(filters)]
        cur = rows[y]
        a = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        c = np.concatenate([np.zeros(ch, np.int32), prev[:-ch]])
        if ft == 0:
            pred = np.zeros_like(cur)
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (a + prev) >> 1
        else:
            pa, pb, pc = np.abs(prev - c), np.abs(a - c), np.abs(a + prev - 2 * c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        raw += bytes([ft]) + ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(typ, body):
        return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body) & 0xffffffff)
    z = zlib.compress(bytes(raw), level)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0))
    if palette is not None:
        out += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    if idat_split:
        out += chunk(b"IDAT", z[:idat_split]) + chunk(b"tEXt", b"k\0v") + chunk(b"IDAT", z[idat_split:])
    else:
        out += chunk(b"IDAT", z)
    return out + chunk(b"IEND", b"")


def _exr_bytes(img, comp, types):
    """A scan-line OpenEXR file with `none` (0), ZIPS (2) or ZIP (3) blocks; channels A, B, G, R stored as half (1) or float (2)"""
    import struct
    import zlib
    h, w = img.shape[:2]
    names = ["A", "B", "G", "R"]
    chan = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", types[n], 0, 0, 0, 0, 1, 1) for n in names) + b"\0"

    def attr(name, typ, body):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<I", len(body)) + body
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = struct.pack("<II", 20000630, 2) + attr("channels", "chlist", chan) + attr("compression", "compression", bytes([comp])) + attr("dataWindow", "box2i", box) + \
        attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + \
        attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    lines = {0: 1, 1: 1, 2: 1, 3: 16}[comp]

    def run_lengths(data):
        out, i, n = bytearray(), 0, len(data)
        while i < n:
            j = i
            while j + 1 < n and data[j + 1] == data[i] and j - i < 126:
                j += 1
            if j - i >= 2:
                out += bytes([j - i, data[i]])
                i = j + 1
            else:
                k = i
                while k < n and k - i < 127 and not (k + 2 < n and data[k] == data[k + 1] == data[k + 2]):
                    k += 1
                out += bytes([(256 - (k - i)) & 255]) + bytes(data[i:k])
                i = k
        return bytes(out)
    blocks = []
    for y0 in range(0, h, lines):
        body = b""
        for y in range(y0, min(h, y0 + lines)):
            for n in names:
                v = img[y, :, "RGBA".index(n)]
                body += (v.astype(np.float16) if types[n] == 1 else v.astype(np.float32)).tobytes()
        if comp:
            b = np.frombuffer(body, np.uint8)
            re = np.concatenate([b[0::2], b[1::2]]).astype(np.int32)
            d = re.copy()
            d[1:] = (re[1:] - re[:-1] + 128 + 256) & 255
            packed = run_lengths(d.astype(np.uint8).tobytes()) if comp == 1 else zlib.compress(d.astype(np.uint8).tobytes(), 6)
            if len(packed) < len(body):
                body = packed
        blocks.append(struct.pack("<iI", y0, len(body)) + body)
    table_at = len(header)
    off, offsets = table_at + 8 * len(blocks), []
    for b in blocks:
        offsets.append(off)
        off += len(b)
    return header + struct.pack(f"<{len(blocks)}Q", *offsets) + b"".join(blocks)


@pytest.mark.parametrize("case", ["rgba_filters", "rgb_fixed_huffman", "gray", "gray_alpha", "palette_split_idat"])
def test_compressed_png_files_decode_like_the_reference(ref, tmp_path, load, case):
    """PNG files as ordinary encoders write them — zlib streams with dynamic and fixed Huffman blocks, the five row filters, grey / grey + alpha /
    RGB / RGBA / palette colour types, IDAT split over chunks — through the module's own inflate + PNG reader, the Python twin (zlib) and the
    reference (stb_image): same RGBA8 pixels after the sRGB step."""
    rng = np.random.default_rng(hash(case) & 0xffff)
    h, w = 24, 40
    smooth = (np.add.outer(np.arange(h) * 5, np.arange(w) * 3)[..., None] + np.arange(4) * 17) & 255
    noisy = rng.integers(0, 256, (h, w, 4))
    img = np.where(rng.random((h, w, 1)) < 0.5, smooth, noisy).astype(np.uint8)
    img[..., 3] = 255
    if case == "rgba_filters":
        data = _png_bytes(img, 6, [0, 1, 2, 3, 4], 9)
    elif case == "rgb_fixed_huffman":
        data = _png_bytes(np.ascontiguousarray(img[:3, :5, :3]), 2, [4, 1], 9)  # a stream this short is written with the fixed code
    elif case == "gray":
        data = _png_bytes(np.ascontiguousarray(img[..., :1]), 0, [3, 4, 0], 6)
    elif case == "gray_alpha":
        img[..., 1] = rng.integers(0, 256, (h, w))
        data = _png_bytes(np.ascontiguousarray(img[..., :2]), 4, [1, 2], 1)
    else:
        pal = rng.integers(0, 256, (37, 3))
        idx = rng.integers(0, 37, (h, w, 1)).astype(np.uint8)
        data = _png_bytes(idx, 3, [0, 2], 9, palette=pal, idat_split=50)
    (tmp_path / "albedo.png").write_bytes(data)
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    mtl = MTL.replace("newmtl Floor\n", "newmtl Floor\nmap_Kd albedo.png\n")
    path = _write_scene(tmp_path, obj=obj, mtl=mtl)
    rs = ref(path)
    sd = load(path)
    assert int(sd.scene["images"]["count"][0]) == int(rs.scene["images"]["count"][0]) == 3  # the placeholder environment, the texture, the pixel filter
    problems = compare_scenes(rs, sd, uv_tangents_exact=load.exact_tangents)
    assert not problems, problems
    rs.close()


@pytest.mark.parametrize("comp,half", [(2, False), (3, False), (3, True), (0, True), (1, True), (1, False)])
def test_compressed_exr_files_decode_like_the_reference(ref, tmp_path, load, comp, half):
    """OpenEXR scan-line files with ZIPS / ZIP blocks (the predictor + byte interleave around a zlib stream) and half-float channels, through the
    module's own reader, the Python twin and the reference (tinyexr): the same float pixels and the same importance-sampling table."""
    rng = np.random.default_rng(comp * 2 + int(half))
    h, w = 37, 48  # 37: a ZIP file ends with a partial 16-line block
    env = (np.add.outer(np.arange(h), np.arange(w))[..., None] / 64.0 + rng.random((h, w, 4)) * 0.25).astype(np.float32)
    env[..., 3] = 1.0
    env[7, 9, :3] = 25.0
    types = {"A": 2, "B": 1 if half else 2, "G": 1 if half else 2, "R": 1 if half else 2}
    (tmp_path / "sky.exr").write_bytes(_exr_bytes(env, comp, types))
    mtl = MTL.replace("newmtl et::env\ncolor 0.1 0.2 0.4", "newmtl et::env\nimage sky.exr\ncolor 0.1 0.2 0.4")
    path = _write_scene(tmp_path, mtl=mtl)
    rs = ref(path)
    sd = load(path)
    problems = compare_scenes(rs, sd, uv_tangents_exact=True)
    assert not problems, problems
    rs.close()


ATMOSPHERE_OBJ = "mtllib s.mtl\nv -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nv -0.3 1.9 -0.3\nv 0.3 1.9 -0.3\nv 0.3 1.9 0.3\nv -0.3 1.9 0.3\nusemtl floor\nf 1 4 3 2\nusemtl lamp\nf 5 6 7 8\n"
ATMOSPHERE_MTL = "newmtl floor\nKd 0.6 0.5 0.4\n\nnewmtl lamp\nKd 0 0 0\nKe 9 9 9\n\nnewmtl et::camera\nviewport 40 30\norigin 0 1 3.5\ntarget 0 0.8 0\nfov 45\n"


@pytest.fixture(scope="module")
def atmosphere_case(ref, tmp_path_factory):
    """One scene with an explicit et::atmosphere block (every parameter off its default, a small sky), loaded ONCE by the reference: its optical-length
    table alone takes the reference half a minute on eight cores."""
    d = tmp_path_factory.mktemp("atmosphere")
    (d / "s.obj").write_text(ATMOSPHERE_OBJ)
    (d / "s.mtl").write_text(ATMOSPHERE_MTL + "\nnewmtl et::atmosphere\ndirection 0.4 0.25 -0.7\nquality 0.0625\nangular_diameter 1.5\nanisotropy 0.7\naltitude 2500\nscale 0.8\n"
                                              "sky_scale 1.25\nsun_scale 0.5\nrayleigh 1.5\nmie 0.75\nozone 0.5\n")
    (d / "s.json").write_text('{"geometry": "s.obj", "materials": "s.mtl", "samples": 4}')
    rs = ref(str(d / "s.json"))
    yield str(d / "s.json"), rs
    rs.close()


def test_atmosphere_block_matches_the_reference_loader(atmosphere_case, load):
    """`newmtl et::atmosphere` (parse_atmosphere_light + render/host/scattering.cxx): the sun profile with its limb-darkened extinction image, the sky
    profile with the single-scattering dome image and its importance table, both black-body spectra — generated by the module's host code."""
    path, rs = atmosphere_case
    sd = load(path)
    assert int(sd.scene["images"]["count"][0]) == int(rs.scene["images"]["count"][0]) == 3  # sun 128 x 128, sky 128 x 64, pixel filter
    problems = compare_scenes(rs, sd, generated_sky=True)
    assert not problems, problems
    ia, ib = _view(rs.scene["images"], S.IMAGE), _view(sd.scene["images"], S.IMAGE)
    assert tuple(ia[1]["isize"]) == tuple(ib[1]["isize"]) == (128, 64) and int(ib[0]["options"]) == 0 and int(ib[1]["options"]) == 1
    sun_a = np.frombuffer(bytes((C.c_char * int(ia[0]["data_size"])).from_address(int(ia[0]["pixels"]["a"]))), np.float32)
    sun_b = np.frombuffer(bytes((C.c_char * int(ib[0]["data_size"])).from_address(int(ib[0]["pixels"]["a"]))), np.float32)
    assert np.array_equal(sun_a, sun_b), "the sun image has no order-dependent term: bit-exact"


def test_a_scene_without_distant_emitters_gets_the_default_atmosphere(tmp_path, load):
    """load_from_file :805-820: no et::dir / et::env / et::atmosphere block in the material file -> sun + sky with the default parameters (direction
    (0, 2, 1), 0.5422 degrees, quality 1 / 8: a 256 x 128 sky), BEFORE the area emitters; a file that declares a distant emitter gets none."""
    (tmp_path / "s.obj").write_text(ATMOSPHERE_OBJ)
    (tmp_path / "s.mtl").write_text(ATMOSPHERE_MTL)
    (tmp_path / "s.json").write_text('{"geometry": "s.obj", "materials": "s.mtl", "samples": 4}')
    sd = load(str(tmp_path / "s.json"))
    prof, inst = _view(sd.scene["emitter_profiles"], S.EMITTER_PROFILE), _view(sd.scene["emitter_instances"], S.EMITTER)
    assert [int(c) for c in prof["cls"]] == [S.EMITTER_DIRECTIONAL, S.EMITTER_ENVIRONMENT, S.EMITTER_AREA]
    assert [int(c) for c in inst["cls"]] == [S.EMITTER_DIRECTIONAL, S.EMITTER_ENVIRONMENT, S.EMITTER_AREA, S.EMITTER_AREA]
    assert int(sd.scene["environment_emitter_count"][0]) == 2
    d = np.array([0.0, 2.0, 1.0], np.float32)
    assert np.allclose(prof["direction"][0], d / np.linalg.norm(d), atol=1e-7) and np.allclose(prof["direction"][1], prof["direction"][0])
    assert np.isclose(float(prof["angular_size"][0]), np.radians(0.5422), rtol=1e-6) and float(prof["angular_size"][1]) == 0.0
    img = _view(sd.scene["images"], S.IMAGE)
    assert [tuple(i["isize"]) for i in img] == [(128, 128), (256, 128), (128, 128)] and [int(i["options"]) for i in img] == [0, 1, 33]
    sky = np.frombuffer(bytes((C.c_char * int(img[1]["data_size"])).from_address(int(img[1]["pixels"]["a"]))), np.float32).reshape(128, 256, 4)
    assert np.isfinite(sky).all() and sky[:64, :, 2].mean() > sky[:64, :, 0].mean() > 0.0, "a blue sky above the horizon"
    (tmp_path / "s.mtl").write_text(ATMOSPHERE_MTL + "\nnewmtl et::dir\ncolor 2 2 2\ndirection 0 1 0\n")
    sd2 = load(str(tmp_path / "s.json"))
    assert [int(c) for c in _view(sd2.scene["emitter_profiles"], S.EMITTER_PROFILE)["cls"]] == [S.EMITTER_DIRECTIONAL, S.EMITTER_AREA]


def _tangent_stress_obj():
    """A smooth UV sphere whose right half mirrors its texture coordinates (orientation flip along a seam), a fin that makes one edge non-manifold, a
    triangle without texture area (it groups with whatever claims it first) and a flat-shaded strip (split normals along shared positions)."""
    rng = np.random.default_rng(21)
    nu, nv = 20, 10
    lines = ["mtllib room.mtl"]
    for j in range(nv + 1):
        for i in range(nu + 1):
            phi, theta = 2 * np.pi * i / nu, np.pi * j / nv
            p = np.array([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)])
            lines.append("v %.6f %.6f %.6f" % tuple(0.7 * p + [0.0, 1.0, 0.0]))
            lines.append("vn %.6f %.6f %.6f" % tuple(p))
            u = i / nu
            lines.append("vt %.6f %.6f" % ((1.0 - u) if i > nu // 2 else u, 1.0 - j / nv + 0.01 * rng.random()))
    lines.append("usemtl Floor")
    for j in range(nv):
        for i in range(nu):
            a, b, c, d = (j * (nu + 1) + i + 1, j * (nu + 1) + i + 2, (j + 1) * (nu + 1) + i + 2, (j + 1) * (nu + 1) + i + 1)
            if j > 0:
                lines.append(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}")
            if j < nv - 1:
                lines.append(f"f {a}/{a}/{a} {c}/{c}/{c} {d}/{d}/{d}")
    n = (nu + 1) * (nv + 1)
    # the fin: a third triangle on the edge between the sphere's vertices (5, 2) and (5, 3)
    a, b = 5 * (nu + 1) + 3, 5 * (nu + 1) + 4
    lines += ["v 0.2 1.1 1.4", "vn 0 0 1", "vt 0.3 0.9", f"f {a}/{a}/{a} {b}/{b}/{b} {n + 1}/{n + 1}/{n + 1}"]
    # no texture area
    lines += ["v 2 0 0", "v 2 1 0", "v 2 0 1", "vt 0.5 0.5", f"f {n + 2}/{n + 2}/{n + 1} {n + 3}/{n + 2}/{n + 1} {n + 4}/{n + 2}/{n + 1}"]
    # flat strip without normals (validate_normals gives every corner its triangle's normal), texture coordinates shared along the strip
    base_v, base_t = n + 5, n + 3
    for k in range(6):
        lines += ["v %.3f %.3f -2" % (0.4 * k, 0.3 * (k % 2)), "v %.3f %.3f -1.5" % (0.4 * k, 0.3 * (k % 2)), "vt %.3f 0" % (k / 5), "vt %.3f 1" % (k / 5)]
    for k in range(5):
        v0, t0 = base_v + 2 * k, base_t + 2 * k
        lines += [f"f {v0}/{t0} {v0 + 2}/{t0 + 2} {v0 + 3}/{t0 + 3}", f"f {v0}/{t0} {v0 + 3}/{t0 + 3} {v0 + 1}/{t0 + 1}"]
    return "\n".join(lines) + "\n"


def test_tangent_frames_match_the_reference_generator(ref, tmp_path):
    """csrc/scene_loader_tangents.inl against the generator the reference calls (thirdparty/mikktspace through build_tangents): mirrored mapping, a
    non-manifold edge, a triangle without texture area, hard edges.  Bit-identical frames; the generator's unsorted final edge run (see the header of
    the .inl) may cost a handful of corners at the end of the mesh, never more."""
    path = _write_scene(tmp_path, obj=_tangent_stress_obj())
    rs = ref(path)
    sd = _load_cpp(path)
    va, vb = _view(rs.scene["vertices"], S.VERTEX), _view(sd.scene["vertices"], S.VERTEX)
    assert va.shape == vb.shape and len(va) > 1000
    assert not compare_scenes(rs, sd, uv_tangents_exact=False)
    same = (va["tan"].view(np.uint32) == vb["tan"].view(np.uint32)).all(axis=1) & (va["btn"].view(np.uint32) == vb["btn"].view(np.uint32)).all(axis=1)
    assert (~same).sum() <= 6, f"{int((~same).sum())} of {len(same)} corners differ"
    flips = np.einsum("ij,ij->i", np.cross(vb["nrm"], vb["tan"]), vb["btn"])
    assert (flips > 0.5).any() and (flips < -0.5).any(), "both orientations occur in the mirrored mapping"
    rs.close()
    sd.close()


def test_corrupt_files_are_refused_or_read_never_worse(tmp_path):
    """A scene's files with random truncations, overwrites, insertions, deletions and bit flips (textures included: PNG, ZIP-compressed EXR, run-length HDR): the
    C++ loader either reads the scene (a texture it cannot decode becomes the 1 x 1 placeholder, like in the reference) or fails with a message.  The same
    mutations were run under AddressSanitizer + UBSan while developing (tools/fuzz_loader.md)."""
    import random
    import shutil
    rng = np.random.default_rng(5)
    base = tmp_path / "base"
    base.mkdir()
    albedo = rng.integers(0, 256, (8, 16, 4), dtype=np.uint8)
    albedo[..., 3] = 255
    (base / "albedo.png").write_bytes(_png_bytes(albedo, 6, [0, 1, 2, 3, 4], 9))
    env = (rng.random((16, 32, 4)) * 2).astype(np.float32)
    env[..., 3] = 1
    (base / "sky.exr").write_bytes(_exr_bytes(env, 3, {"A": 2, "B": 1, "G": 1, "R": 1}))
    _write_hdr(str(base / "bump.hdr"), env[..., :3], True)
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    mtl = MTL.replace("newmtl et::env\ncolor 0.1 0.2 0.4", "newmtl et::env\nimage sky.exr\ncolor 0.1 0.2 0.4").replace("newmtl Floor\n", "newmtl Floor\nmap_Kd albedo.png\nnormalmap image bump.hdr\n")
    _write_scene(base, obj=obj, mtl=mtl)
    files = sorted(p.name for p in base.iterdir())
    r = random.Random(77)
    read = refused = 0
    for it in range(80):
        work = tmp_path / "work"
        if work.exists():
            shutil.rmtree(work)
        shutil.copytree(base, work)
        victim = r.choice(files)
        d = bytearray((work / victim).read_bytes())
        mode = r.randrange(5)
        if mode == 0:
            d = d[:r.randrange(len(d))]
        elif mode == 1:
            for _ in range(r.randrange(1, 8)):
                d[r.randrange(len(d))] = r.randrange(256)
        elif mode == 2:
            i = r.randrange(len(d))
            d[i:i] = bytes(r.randrange(256) for _ in range(r.randrange(1, 16)))
        elif mode == 3:
            i = r.randrange(len(d))
            del d[i:min(len(d), i + r.randrange(1, 32))]
        else:
            for _ in range(r.randrange(1, 4)):
                d[r.randrange(len(d))] ^= 1 << r.randrange(8)
        (work / victim).write_bytes(bytes(d))
        try:
            sf = _load_cpp(str(work / "room.json"))
            assert sf.triangle_count >= 0
            sf.close()
            read += 1
        except api.EtxbError as e:
            assert len(str(e)) > 20, "a refusal carries its reason"
            refused += 1
    assert read + refused == 80 and read > 0 and refused > 0


NVDB_MAKE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "nvdb_make")


def _medium_density(scene):
    out = []
    for m in _view(scene["mediums"], S.MEDIUM):
        n = int(m["density"]["count"])
        out.append(np.frombuffer(bytes((C.c_char * (n * 4)).from_address(int(m["density"]["a"]))), np.float32) if n else np.zeros(0, np.float32))
    return out


@pytest.mark.parametrize("kind", ["sphere", "blobs", "empty", "vec3"])
def test_nanovdb_volume_matches_the_reference_loader(ref, tmp_path, kind, load):
    """`et::medium … volume cloud.nvdb`: a file written with the reference's NanoVDB headers (oracle/nvdb_make.cxx, test infrastructure) read by the module's
    own NanoVDB reader and by the reference (nanovdb::io::readGrid + an accessor sweep, medium_pool.cxx:102-159): same class, dimensions and dense
    grid — a fog sphere; two blobs across node boundaries plus a far block (negative coordinates, a sparse tree); an all-negative grid and a Vec3f grid
    (the medium stays homogeneous in both)."""
    import subprocess
    if not os.path.exists(NVDB_MAKE):
        pytest.skip("oracle/_ref/nvdb_make not built")
    subprocess.check_call([NVDB_MAKE, kind, str(tmp_path / "cloud.nvdb")])
    path = _write_scene(tmp_path, mtl=MTL.replace("scattering 0.4", "scattering 0.4\nvolume cloud.nvdb"))
    rs = ref(path)
    sd = load(path)
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    ma, mb = _view(rs.scene["mediums"], S.MEDIUM), _view(sd.scene["mediums"], S.MEDIUM)
    heterogeneous = [int(c) for c in mb["cls"]]
    assert heterogeneous.count(1) == (1 if kind in ("sphere", "blobs") else 0)
    for a, b in zip(_medium_density(rs.scene), _medium_density(sd.scene)):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if kind == "blobs":
        k = heterogeneous.index(1)
        assert tuple(int(x) for x in mb[k]["dimensions"]) == tuple(int(x) for x in ma[k]["dimensions"]) and float(_medium_density(sd.scene)[k].max()) == 1.0
    rs.close()


def test_zip_compressed_nanovdb_file_and_refusals(tmp_path):
    """The ZIP codec of NanoVDB files (a 64-bit size + one zlib stream, util/IO.h) through the module's own inflate: the same grid as the uncompressed file;
    BLOSC files, other versions and truncated files are refused with a message."""
    import struct
    import subprocess
    import zlib
    if not os.path.exists(NVDB_MAKE):
        pytest.skip("oracle/_ref/nvdb_make not built")
    subprocess.check_call([NVDB_MAKE, "blobs", str(tmp_path / "cloud.nvdb")])
    path = _write_scene(tmp_path, mtl=MTL.replace("scattering 0.4", "scattering 0.4\nvolume cloud.nvdb"))
    plain = _load_cpp(path)
    want = [d.copy() for d in _medium_density(plain.scene)]
    plain.close()
    raw = bytearray((tmp_path / "cloud.nvdb").read_bytes())
    name_size, = struct.unpack_from("<I", raw, 16 + 136)
    body_at = 16 + 176 + name_size
    grid_size, = struct.unpack_from("<Q", raw, 16)
    assert body_at + grid_size == len(raw)
    packed = zlib.compress(bytes(raw[body_at:]), 6)
    head = bytearray(raw[:body_at])
    struct.pack_into("<H", head, 14, 1)            # Header::codec
    struct.pack_into("<Q", head, 16 + 8, 8 + len(packed))  # MetaData::fileSize
    struct.pack_into("<H", head, 16 + 168, 1)      # MetaData::codec
    (tmp_path / "cloud.nvdb").write_bytes(bytes(head) + struct.pack("<Q", len(packed)) + packed)
    zipped = _load_cpp(path)
    got = _medium_density(zipped.scene)
    assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want)) and sum(len(a) for a in got) > 10000
    zipped.close()
    struct.pack_into("<H", head, 16 + 168, 2)
    (tmp_path / "cloud.nvdb").write_bytes(bytes(head) + struct.pack("<Q", len(packed)) + packed)
    with pytest.raises(api.EtxbError, match="BLOSC"):
        _load_cpp(path)
    (tmp_path / "cloud.nvdb").write_bytes(bytes(raw[:len(raw) // 2]))
    with pytest.raises(api.EtxbError, match="outside the file"):
        _load_cpp(path)
    older = bytearray(raw)
    struct.pack_into("<I", older, 8, 30 << 21)
    (tmp_path / "cloud.nvdb").write_bytes(bytes(older))
    with pytest.raises(api.EtxbError, match="30.x"):
        _load_cpp(path)
    path = _write_scene(tmp_path, mtl=MTL.replace("scattering 0.4", "scattering 0.4\nvolume cloud.vdb"))
    with pytest.raises(api.EtxbError, match="nvdb"):
        _load_cpp(path)


FUZZ_OBJ = """mtllib room.mtl
v -1 0 -1
v 1 0 -1
v 1 0 1
v -1 0 1
v -0.3 1.9 -0.3
v 0.3 1.9 -0.3
v 0.3 1.9 0.3
v -0.3 1.9 0.3
v -1 2 -1
v 1 2 -1
v 1 2 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
o floor
usemtl m0
f 1/1 4/4 3/3 2/2
o lamp
usemtl m1
f 5 6 7 8
o top
usemtl m2
f 9/1 10/2 11/3
usemtl m3
f 1 2 10 9
"""


def _random_material_file(rng):
    """A random walk through the material dialect: mostly well-formed values, now and then a doubled space (the reference's tokenizer keeps the empty piece),
    a missing number, an unknown name, a missing texture file.  Cases whose meaning in the reference is undefined behaviour or a NaN camera are left out."""
    def num():
        return rng.choice(["0", "1", "0.5", "0.25", "2.5", "1e-3", "10", "0.8", "0.05", "3", "1.5", "0.33", "7e-2"]) if rng.random() < 0.95 else rng.choice(["-1", "abc", ""])

    def rgb():
        return (" " if rng.random() < 0.9 else "  ").join(num() for _ in range(rng.choice([3, 3, 3, 3, 3, 1, 2, 4])))
    classes = ["diffuse", "translucent", "plastic", "conductor", "dielectric", "thinfilm", "mirror", "boundary", "velvet", "principled", "void", "bogus"]
    iors = ["gold", "silver", "water", "glass", "copper", "1.33", "1.5 0.2", "0.2 3.9", "unobtainium", "diamond", "plastic", "chrome"]

    def material(name, others):
        o = []
        if rng.random() < 0.7: o.append("Kd " + rgb())
        if rng.random() < 0.4: o.append("Ks " + rgb())
        if rng.random() < 0.2: o.append("Kt " + rgb())
        if rng.random() < 0.3: o.append("Ke " + rng.choice([rgb(), "blackbody 3000", "nblackbody 5000 scale 2", "sunlike", "5", "warm"]))
        if rng.random() < 0.6: o.append("material class " + rng.choice(classes))
        if rng.random() < 0.4: o.append("Pr " + rng.choice([num(), num() + " " + num()]))
        if rng.random() < 0.3: o.append("int_ior " + rng.choice(iors))
        if rng.random() < 0.2: o.append("ext_ior " + rng.choice(iors))
        if rng.random() < 0.2: o.append("two_sided " + rng.choice(["1", "0", "true", "on", "off", "yes", "true 1"]))
        if rng.random() < 0.2: o.append("opacity " + num())
        if rng.random() < 0.2: o.append("metalness " + num())
        if rng.random() < 0.2: o.append("transmission " + num())
        if rng.random() < 0.2: o.append("diffuse " + rng.choice(["0", "1", "2", "x"]))
        if rng.random() < 0.25: o.append("thinfilm " + rng.choice(["range 100 500 ior 1.3", "ior water range 50 60", "range 10", "ior 1.4"]))
        if rng.random() < 0.25: o.append("subsurface " + rng.choice(["path refracted distances 1 0.5 0.2 scale 0.1", "class approximate scale 2", "distances 0.1 0.2 0.3", "path diffuse"]))
        if rng.random() < 0.3: o.append("emitter " + rng.choice(["color 3 2 1 scale 2", "blackbody 4500 scale 0.5 twosided", "nblackbody 6500 collimated 0.5", "scale 3", "collimated 2 color 1 1", "twosided"]))
        if rng.random() < 0.25 and others: o.append("base " + rng.choice(others))
        if rng.random() < 0.2: o.append("int_medium " + rng.choice(["fog", "haze", "none"]))
        if rng.random() < 0.1: o.append("ext_medium " + rng.choice(["fog", "haze"]))
        if rng.random() < 0.15: o.append("normalmap scale " + num())
        if rng.random() < 0.2: o.append("map_Ml tex.png" + rng.choice(["", " channel 2", " channel -3", "  channel 1"]))
        if rng.random() < 0.2: o.append("map_Tm " + rng.choice(["tex.png channel 1", "missing.png", "tex.png"]))
        if rng.random() < 0.4:
            o.append(rng.choice(["map_Kd", "map_Ks", "map_Ke", "map_Kt", "map_kd", "MAP_KD", "map_Ka", "map_Bump", "bump", "map_d", "disp", "norm", "refl -type sphere"]) + " " +
                     rng.choice(["tex.png", "-s 1 1 1 tex.png", "-bm 0.5 tex.png", "-clamp on tex.png", "-o 0.1 0.2 tex.png -mm 0 1", "tex.png tex.png", "my tex.png", "-blendu off -blendv on tex.png",
                                 "  tex.png", "-imfchan r tex.png", "-texres 512 tex.png", "-boost 2 tex.png", "-colorspace linear tex.png"]))
        if rng.random() < 0.3: o.append(rng.choice(["Ns 10", "Ni 1.45", "d 0.5", "Tr 0.2", "illum 2", "Ka 0.1 0.1 0.1", "Tf 1 1 1", "Pm 0.5", "Ps 0.1", "Pc 0.2", "Pcr 0.3", "aniso 0.1", "anisor 0.2", "Ke", "Kd", "d"]))
        if rng.random() < 0.1: o.append("map_Pr tex.png channel 1")
        rng.shuffle(o)
        return "\n".join(["newmtl " + name] + o) + "\n"
    parts = ["newmtl et::spectrum\nid warm\n" + rng.choice(["blackbody 2800", "rgb 1 0.6 0.3\nilluminant", "samples 400 0.2 500 0.5 700 1.0\nnormalize " + rng.choice(["luminance", "max"]),
                                                           "nblackbody 4000 scale 3", "rgb 0.5 0.5", "samples 400 1 500"]) + "\n" + (("scale " + num() + "\n") if rng.random() < 0.5 else ""),
             "newmtl et::medium\nid fog\n" + rng.choice(["absorption 0.1 0.2 0.3\nscattering 0.5", "scattering 0.2 0.3 0.4\ng 0.5", "rayleigh scale 0.01", "mie scale 2\nanisotropy 0.7",
                                                        "parametric color 0.9 0.5 0.3 distance 0.5", "absorbtion 0.3", "parametric distances 1 2 3 scale 0.2"]) + "\n" + ("enclosed\n" if rng.random() < 0.3 else "")]
    if rng.random() < 0.9:
        parts.append("newmtl et::dir\ncolor " + rng.choice([rgb(), "warm", "blackbody 5500 scale 2", "7"]) + "\ndirection " + rgb() + "\n" + (("angular_diameter " + num() + "\n") if rng.random() < 0.5 else ""))
    else:
        parts.append("newmtl et::env\ncolor " + rgb() + "\n" + (("rotation " + num() + "\n") if rng.random() < 0.5 else "") + (("scale " + num() + "\n") if rng.random() < 0.5 else ""))
    if rng.random() < 0.5:
        parts.append("newmtl et::env\ncolor " + rng.choice([rgb(), "warm"]) + "\n")
    if rng.random() < 0.7:
        keys = ["target 0 1 0", "up 0 1 0", "fov " + num(), "focal-length 35", "lens-radius 0.01", "focal-distance 3", "clip-near 0.1", "clip-far 100", "class eq", "ext_medium fog", "active 1"]
        parts.append("newmtl et::camera\n" + "\n".join(rng.sample(keys, rng.randrange(1, 8))) + "\nviewport 40 30\norigin 0.5 1 4\n")
    names = []
    for k in range(4):
        if rng.random() < 0.1:
            continue
        parts.append(material("m%d" % k, names))
        names.append("m%d" % k)
    return "\n".join(parts)


def test_random_material_files_match_the_reference_loader(ref, tmp_path):
    """Differential test of the material dialect: 60 random material files (seeded) read by the C++ loader and by the reference's loader, PODs compared like
    everywhere else in this file.  This is the test that found: kEpsilon is FLT_EPSILON, the reference's tokenizer keeps the empty piece between two spaces,
    `two_sided` compares its whole value, texture paths are not checked for existence (a missing file is the white placeholder), `map_Ml` / `map_Tm`, and that
    the scene description's values are type-checked (`"spectral": 1` is ignored)."""
    import random
    rng = random.Random(2024)
    tex = np.random.default_rng(3).integers(0, 256, (4, 6, 4), dtype=np.uint8)
    tex[..., 3] = 255
    for it in range(60):
        d = tmp_path / f"s{it}"
        d.mkdir()
        (d / "room.obj").write_text(FUZZ_OBJ)
        (d / "tex.png").write_bytes(_png_bytes(tex, 6, [0, 1], 6))
        (d / "room.mtl").write_text(_random_material_file(rng))
        js = {"geometry": "room.obj", "materials": "room.mtl", "samples": rng.choice([1, 16, 300, 0]), "spectral": rng.choice([True, False, 1, 0])}  # 1 / 0 are not booleans: ignored
        if rng.random() < 0.5:
            options = (("class", rng.choice(["perspective", "eq"])), ("fov", rng.choice([40, 65.5, 120])), ("focal-length", rng.choice([24, 50.0])), ("lens-radius", 0.02),
                       ("focal-distance", 2.5), ("clip-near", 0.05), ("clip-far", 200.0), ("origin", [0.2, 1, 3]), ("target", [0, 1, 0]), ("up", [0, 1, 0.1]),
                       ("viewport", rng.choice([[32, 24], [17, 9]])))
            js["camera"] = {k: v for k, v in options if rng.random() < 0.5}
        for key, choices in (("max-path-length", [0, 1, 7, 100000]), ("min-path-length", [0, 1, 3]), ("random-termination-start", [0, 1, 9]), ("force-tangents", [True, False])):
            if rng.random() < 0.3:
                js[key] = rng.choice(choices)
        (d / "room.json").write_text(json.dumps(js))
        rs = ref(str(d / "room.json"))
        sd = _load_cpp(str(d / "room.json"))
        problems = compare_scenes(rs, sd)
        assert not problems, (it, problems, (d / "room.mtl").read_text())
        rs.close()
        sd.close()


def _random_obj_file(rng):
    """Random geometry in the .obj dialect: polygons with 3-6 corners (quads split along the shorter diagonal, larger ones by ear clipping), the four index
    styles, negative indices, doubled spaces, `o` / `g` / `s` lines, material switches incl. an undeclared one (the index-cursor quirk), a degenerate face."""
    def f():
        return "%.3f" % rng.uniform(-2, 2)
    lines = ["mtllib room.mtl"]
    nv, nn, nt = rng.randrange(6, 30), rng.choice([0, 0, rng.randrange(1, 10)]), rng.choice([0, rng.randrange(1, 10)])
    lines += ["v %s %s %s" % (f(), f(), f()) + (" 1.0" if rng.random() < 0.1 else "") for _ in range(nv)]
    lines += ["vn %s %s %s" % (f(), f(), f()) for _ in range(nn)]
    lines += ["vt %s %s" % (f(), f()) + (" 0" if rng.random() < 0.2 else "") for _ in range(nt)]
    for face in range(rng.randrange(3, 25)):
        r = rng.random()
        if r < 0.25:
            lines.append("usemtl " + rng.choice(["a", "b", "lamp", "A", "nope", "b"]))
        elif r < 0.35:
            lines.append(rng.choice(["o", "g"]) + " part%d" % face)
        elif r < 0.4:
            lines.append("s " + rng.choice(["1", "off", "0"]))
        elif r < 0.43:
            lines.append("# comment")
        corners = rng.sample(range(1, nv + 1), rng.choice([3, 3, 3, 4, 4, 5, 6]))
        style, negative, tokens = rng.choice(["v", "v/t", "v//n", "v/t/n"]), rng.random() < 0.15, []
        for i in corners:
            vi = (i - nv - 1) if negative else i
            t, n = (rng.randrange(1, nt + 1) if nt else None), (rng.randrange(1, nn + 1) if nn else None)
            if style == "v" or (style == "v/t" and t is None) or (style == "v//n" and n is None) or (style == "v/t/n" and (t is None or n is None)):
                tokens.append(str(vi))
            elif style == "v/t":
                tokens.append(f"{vi}/{t}")
            elif style == "v//n":
                tokens.append(f"{vi}//{n}")
            else:
                tokens.append(f"{vi}/{t}/{n}")
        lines.append("f " + (" " if rng.random() < 0.9 else "  ").join(tokens))
    if rng.random() < 0.2:
        lines.append("f 1 1 2")
    return "\n".join(lines) + "\n"


def test_random_obj_files_match_the_reference_loader(ref, tmp_path):
    """Differential test of the geometry side: 60 random .obj files (seeded) through the C++ loader and the reference's (tinyobjloader + load_from_obj +
    validate_normals + the tangent-space generator + commit).  It found the ear clipping of polygons with more than four corners."""
    import random
    rng = random.Random(77)
    mtl = ("newmtl et::dir\ncolor 2 2 2\ndirection 0.2 1 0.3\n\nnewmtl a\nKd 0.5 0.5 0.5\n\nnewmtl b\nKd 0.2 0.6 0.2\nmaterial class plastic\n\nnewmtl lamp\nKe 5 5 5\n\n"
           "newmtl et::camera\nviewport 32 24\norigin 0.3 1 6\ntarget 0 0.5 0\n")
    for it in range(60):
        d = tmp_path / f"o{it}"
        d.mkdir()
        (d / "room.obj").write_text(_random_obj_file(rng))
        (d / "room.mtl").write_text(mtl)
        (d / "room.json").write_text(json.dumps({"geometry": "room.obj", "materials": "room.mtl", "samples": 4, "force-tangents": rng.random() < 0.2}))
        rs = ref(str(d / "room.json"))
        sd = _load_cpp(str(d / "room.json"))
        problems = compare_scenes(rs, sd)
        assert not problems, (it, problems, (d / "room.obj").read_text())
        rs.close()
        sd.close()


def _png_any(samples, ctype, depth, interlace=False, trns=None, palette=None, level=6):
    """A PNG of any form: `samples` (h, w, channels) integers below 2**depth; bit depths 1 / 2 / 4 / 8 / 16; Adam7 interlacing; a tRNS chunk; a palette.
    Row filters cycle through the five types."""
    import struct
    import zlib
    h, w, ch = samples.shape

    def pack_rows(block):
        out = bytearray()
        ph, pw = block.shape[:2]
        prev = np.zeros(0, np.int32)
        bpp = max(1, ch * depth // 8)
        for y in range(ph):
            row = block[y].reshape(-1).astype(np.uint32)
            if depth == 16:
                data = np.stack([row >> 8, row & 255], axis=1).reshape(-1).astype(np.int32)
            elif depth == 8:
                data = row.astype(np.int32)
            else:
                bits = np.zeros(((len(row) * depth + 7) // 8) * 8, np.uint8)
                for k in range(depth):
                    bits[np.arange(len(row)) * depth + k] = (row >> (depth - 1 - k)) & 1
                data = np.packbits(bits).astype(np.int32)
            if len(prev) != len(data):
                prev = np.zeros(len(data), np.int32)
            ft = y % 5
            a = np.concatenate([np.zeros(bpp, np.int32), data[:-bpp]]) if len(data) > bpp else np.zeros(len(data), np.int32)
            c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if len(data) > bpp else np.zeros(len(data), np.int32)
            if ft == 0:
                pred = np.zeros_like(data)
            elif ft == 1:
                pred = a
            elif ft == 2:
                pred = prev
            elif ft == 3:
                pred = (a + prev) >> 1
            else:
                pa, pb, pc = np.abs(prev - c), np.abs(a - c), np.abs(a + prev - 2 * c)
                pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
            out += bytes([ft]) + ((data - pred) & 255).astype(np.uint8).tobytes()
            prev = data
        return bytes(out)
    if interlace:
        raw = b""
        for ox, oy, sx, sy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            block = samples[oy::sy, ox::sx]
            if block.shape[0] and block.shape[1]:
                raw += pack_rows(block)
    else:
        raw = pack_rows(samples)

    def chunk(typ, body):
        return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body) & 0xffffffff)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        out += chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None:
        out += chunk(b"tRNS", bytes(trns))
    return out + chunk(b"IDAT", zlib.compress(raw, level)) + chunk(b"IEND", b"")


PNG_FORMS = ["rgb16", "rgba16_interlaced", "gray1", "gray2_interlaced", "gray4", "gray16", "palette2", "palette4_trns_interlaced", "palette8_trns", "rgb8_key", "rgb16_key", "gray8_key",
             "rgba8_interlaced", "gray_alpha16"]


@pytest.mark.parametrize("form", PNG_FORMS)
def test_every_png_form_decodes_like_the_reference(ref, tmp_path, form, load):
    """Bit depths 1 / 2 / 4 / 16, Adam7 interlacing, colour keys and palette alpha: whatever stb_image decodes, reduced to 8 bits its way (16-bit samples keep the
    high byte, low-depth grey is scaled) — and grey + alpha forms (incl. grey with a colour key) zero-filled like in the reference's texture pool."""
    import struct
    rng = np.random.default_rng(PNG_FORMS.index(form))
    h, w = 11, 13  # odd sizes: partial bytes at low depths, short Adam7 passes
    kw = {}
    if form.startswith("rgb16"):
        s, ct, dp = rng.integers(0, 65536, (h, w, 3)), 2, 16
        if form.endswith("key"):
            s[3, 4] = (1000, 2000, 3000)
            kw["trns"] = struct.pack(">HHH", 1000, 2000, 3000)
    elif form.startswith("rgba16"):
        s, ct, dp = rng.integers(0, 65536, (h, w, 4)), 6, 16
    elif form.startswith("rgba8"):
        s, ct, dp = rng.integers(0, 256, (h, w, 4)), 6, 8
    elif form == "gray_alpha16":
        s, ct, dp = rng.integers(0, 65536, (h, w, 2)), 4, 16
    elif form.startswith("gray") and not form.startswith("gray_"):
        dp = int("".join(c for c in form.split("_")[0] if c.isdigit()))
        s, ct = rng.integers(0, 2 ** dp, (h, w, 1)), 0
        if form.endswith("key"):
            kw["trns"] = struct.pack(">H", int(s[2, 2, 0]))
    elif form.startswith("palette"):
        dp = int("".join(c for c in form.split("_")[0] if c.isdigit()))
        n = min(2 ** dp, 200)
        s, ct = rng.integers(0, n, (h, w, 1)), 3
        kw["palette"] = rng.integers(0, 256, (n, 3))
        if "trns" in form:
            kw["trns"] = bytes(rng.integers(0, 256, n // 2).tolist())
    else:
        s, ct, dp = rng.integers(0, 256, (h, w, 3)), 2, 8
        s[5, 6] = (9, 8, 7)
        kw["trns"] = struct.pack(">HHH", 9, 8, 7)
    (tmp_path / "albedo.png").write_bytes(_png_any(s, ct, dp, interlace="interlaced" in form, **kw))
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    path = _write_scene(tmp_path, obj=obj, mtl=MTL.replace("newmtl Floor\n", "newmtl Floor\nmap_Kd albedo.png\n"))
    rs = ref(path)
    sd = load(path)
    ia = _view(rs.scene["images"], S.IMAGE)
    assert tuple(ia[1]["isize"]) == (w, h) and int(ia[1]["format"]) == 2, "the reference decoded the file (not its 1 x 1 placeholder)"
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()


JPEG_FORMS = ["baseline_444", "baseline_420_odd", "progressive_422", "progressive_420_q20", "grey", "grey_progressive", "restart_optimized", "adobe_rgb", "tiny"]


@pytest.mark.parametrize("form", JPEG_FORMS)
def test_jpeg_textures_decode_like_the_reference(ref, tmp_path, form, load):
    """.jpg textures: the reference decodes them with stb_image; the module's JPEG reader (csrc/scene_loader_jpeg.inl) restates that decoder's inverse DCT,
    chroma upsampling and colour conversion, so the RGBA8 texels are the same bytes — baseline and progressive, 4:4:4 / 4:2:2 / 4:2:0, grey, restart intervals
    with optimised Huffman tables, an RGB (Adobe transform 0) file, odd sizes.  140 further files were compared while developing."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(JPEG_FORMS.index(form))
    h, w = (9, 11) if form == "tiny" else ((33, 47) if "odd" in form else (40, 56))
    y, x = np.mgrid[0:h, 0:w]
    a = (np.stack([(x * 5 + y * 3) % 256, (x * 2 + y * 7) % 256, (x * y) % 256], -1) + rng.integers(0, 60, (h, w, 3))).clip(0, 255).astype(np.uint8)
    file = tmp_path / "albedo.jpg"
    if form.startswith("grey"):
        Image.fromarray(a[..., 0]).save(file, "JPEG", quality=80, progressive="progressive" in form)
    elif form == "restart_optimized":
        cv2 = pytest.importorskip("cv2")
        cv2.imwrite(str(file), a, [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_RST_INTERVAL, 3, cv2.IMWRITE_JPEG_OPTIMIZE, 1])
    elif form == "adobe_rgb":
        Image.fromarray(a).save(file, "JPEG", quality=85, keep_rgb=True)
    else:
        sub = {"444": 0, "422": 1, "420": 2}[[t for t in form.split("_") if t in ("444", "422", "420")][0]] if form != "tiny" else 2
        Image.fromarray(a).save(file, "JPEG", quality=20 if "q20" in form else 88, subsampling=sub, progressive="progressive" in form)
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    path = _write_scene(tmp_path, obj=obj, mtl=MTL.replace("newmtl Floor\n", "newmtl Floor\nmap_Kd albedo.jpg\n"))
    rs = ref(path)
    sd = load(path)
    ia = _view(rs.scene["images"], S.IMAGE)
    assert tuple(ia[1]["isize"]) == (w, h) and int(ia[1]["format"]) == 2, "the reference decoded the file (not its 1 x 1 placeholder)"
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()


TGA_FORMS = ["rgb", "rgba_rle", "grey_top_down", "palette_rle", "rgb16", "grey16"]


@pytest.mark.parametrize("form", TGA_FORMS)
def test_tga_textures_decode_like_the_reference(ref, tmp_path, form, load):
    """.tga textures the way stb_image reads them: true colour (blue first) and grey, raw and run-length encoded, bottom-up and top-down, colour-mapped, x555
    16-bit colour scaled by 255 / 31, and 16-bit grey + alpha (zero-filled by the reference's texture pool like every two-channel image)."""
    import struct
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(TGA_FORMS.index(form))
    h, w = 7, 9
    a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    a[:, 3:6] = a[:, 3:4]
    file = str(tmp_path / "albedo.tga")
    if form == "rgb":
        Image.fromarray(a[..., :3]).save(file)
    elif form == "rgba_rle":
        Image.fromarray(a).save(file, compression="tga_rle")
    elif form == "grey_top_down":
        Image.fromarray(a[..., 0]).save(file, orientation=1)
    elif form == "palette_rle":
        Image.fromarray(a[..., :3]).quantize(16).save(file, compression="tga_rle")
    else:
        body = rng.integers(0, 65536, h * w, dtype=np.uint16).tobytes()
        open(file, "wb").write(struct.pack("<BBBHHBHHHHBB", 0, 0, 2 if form == "rgb16" else 3, 0, 0, 0, 0, 0, w, h, 16, 0) + body)
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    path = _write_scene(tmp_path, obj=obj, mtl=MTL.replace("newmtl Floor\n", "newmtl Floor\nmap_Kd albedo.tga\n"))
    rs = ref(path)
    sd = load(path)
    ia = _view(rs.scene["images"], S.IMAGE)
    assert tuple(ia[1]["isize"]) == (w, h) and int(ia[1]["format"]) == 2, "the reference decoded the file (not its 1 x 1 placeholder)"
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()


BMP_FORMS = ["rgb24", "rgba32", "palette8", "one_bit", "x555", "rgb565_top_down", "rgb32_zero_alpha"]


@pytest.mark.parametrize("form", BMP_FORMS)
def test_bmp_textures_decode_like_the_reference(ref, tmp_path, form, load):
    """.bmp textures the way stb_image reads them: 24 / 32 bits, palettes of 8 and 1 bit, 16 bits through the default x555 masks and through BI_BITFIELDS 565
    (channels widened by bit replication), top-down rows, and a 32-bit file whose alpha bytes are all zero (opaque)."""
    import struct
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(BMP_FORMS.index(form))
    h, w = 7, 9
    a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    file = str(tmp_path / "albedo.bmp")

    def write(bpp, rows, compress=0, masks=b"", top_down=False):
        body = b"".join(r + b"\0" * ((-len(r)) & 3) for r in rows)
        off = 14 + 40 + len(masks)
        open(file, "wb").write(b"BM" + struct.pack("<IHHI", off + len(body), 0, 0, off) + struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, bpp, compress, len(body), 2835, 2835, 0, 0) +
                               masks + body)
    if form == "rgb24":
        Image.fromarray(a[..., :3]).save(file)
    elif form == "rgba32":
        Image.fromarray(a).save(file)
    elif form == "palette8":
        Image.fromarray(a[..., :3]).quantize(16).save(file)
    elif form == "one_bit":
        Image.fromarray(a[..., 0] > 127).convert("1").save(file)
    elif form == "x555":
        write(16, [rng.integers(0, 65536, w, dtype=np.uint16).tobytes() for _ in range(h)])
    elif form == "rgb565_top_down":
        write(16, [rng.integers(0, 65536, w, dtype=np.uint16).tobytes() for _ in range(h)], compress=3, masks=struct.pack("<III", 0xF800, 0x07E0, 0x001F), top_down=True)
    else:
        write(32, [(rng.integers(0, 2 ** 32, w, dtype=np.uint32) & 0x00ffffff).astype(np.uint32).tobytes() for _ in range(h)])
    obj = OBJ.replace("vn 0 1 0\n", "vn 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n").replace("f 1//1 4//1 3//1 2//1", "f 1/1/1 4/4/1 3/3/1 2/2/1")
    path = _write_scene(tmp_path, obj=obj, mtl=MTL.replace("newmtl Floor\n", "newmtl Floor\nmap_Kd albedo.bmp\n"))
    rs = ref(path)
    sd = load(path)
    ia = _view(rs.scene["images"], S.IMAGE)
    assert tuple(ia[1]["isize"]) == (w, h) and int(ia[1]["format"]) == 2, "the reference decoded the file (not its 1 x 1 placeholder)"
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()


@pytest.mark.parametrize("half", [True, False])
def test_piz_compressed_exr_decodes_like_the_reference(ref, tmp_path, half, load):
    """PIZ — OpenEXR's default codec (value bitmap + lookup table, wavelet transform, Huffman coding with run lengths) — written by the OpenEXR library itself
    (through OpenCV), read by the module's own decoder and by the reference (tinyexr): the same environment map and importance table.  48 further files
    (noise, gradients, flat, sparse; sizes 1 x 1 ... 300 x 5; half and float) were compared with OpenEXR's own reader while developing."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(int(half))
    h, w = 37, 50
    y, x = np.mgrid[0:h, 0:w]
    env = (np.stack([x / 40.0, y / 30.0, (x + y) / 60.0], -1) + rng.random((h, w, 3)) * 0.5).astype(np.float32)
    env[7, 9] = 30.0
    file = str(tmp_path / "sky.exr")
    if not cv2.imwrite(file, env[..., ::-1], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF if half else cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ]):
        pytest.skip("this OpenCV build does not write OpenEXR files")
    mtl = MTL.replace("newmtl et::env\ncolor 0.1 0.2 0.4", "newmtl et::env\nimage sky.exr\ncolor 0.1 0.2 0.4")
    path = _write_scene(tmp_path, mtl=mtl)
    rs = ref(path)
    sd = load(path)
    ia = _view(rs.scene["images"], S.IMAGE)
    assert tuple(ia[0]["isize"]) == (w, h), "the reference decoded the file (not its 1 x 1 placeholder)"
    problems = compare_scenes(rs, sd)
    assert not problems, problems
    rs.close()
