"""Scene files of the reference: the `.json` scene description + Wavefront `.obj` geometry + the `.mtl` dialect with the
`et::camera / et::medium / et::dir / et::env / et::spectrum` blocks and the per-material directives of
sources/etx/render/host/scene_representation.cxx (load_from_file :679-838, load_from_obj :964-1052, parse_camera :1054-1159,
parse_medium :1161-1306, parse_directional_light :1308-1343, parse_env_light :1345-1378, parse_spectrum :1496-1611,
load_reflectance_spectrum / load_illuminant_spectrum :1613-1680, parse_material :1682-2079, validate_materials :262-302,
validate_normals / validate_tangents :304-418, commit :420-455) read into the Scene / Camera PODs `etxb_upload_scene` takes.

Everything is built on `scenes.SceneData` (the same record builders the synthetic generators use); `tests/test_loader.py` compares the
result with the reference's own loader (compiled in place as test infrastructure) array by array on the shipped Cornell asset and on
generated scene files.  This file is the Python twin of csrc/scene_loader.cpp (what the C ABI ships) and shares three pieces of host code with it through
the C ABI: the tangent-space generator of meshes with texture coordinates (etxb_mesh_tangents), the NanoVDB reader (etxb_nvdb_density) and the two images of an
atmosphere block (etxb_atmosphere_images).  Remaining difference to the reference: spectra in the last bits (numpy's float32 arithmetic against the C++ builders; 2e-6 in the tests).
glTF geometry is refused.  Image files: PNG (8-bit, non-interlaced), OpenEXR (float, scan lines, none / ZIP), Radiance HDR and the reference's PFM variant.
The .mtl reader follows the reference's patched tinyobjloader (thirdparty/tinyobjloader/tiny_obj_loader.hxx:1900-2190): names are lower-cased,
`Kd / Ks / Kt / Ke` and every non-standard key land in the material's parameter list, the standard texture keys are consumed.
"""
import ctypes as C
import json
import math
import os
import struct
import zlib

import numpy as np

from . import scenes
from . import structs as S
from .scenes import f32

# keys the patched tinyobjloader consumes itself (they never reach get_param)
_TINYOBJ_TEXTURES = {"map_ka": "ambient", "map_kd": "diffuse", "map_ks": "specular", "map_kt": "transmittance", "map_ns": "specular_highlight", "map_bump": "bump",
                     "map_d": "alpha", "disp": "displacement", "refl": "reflection", "map_pr": "roughness", "map_pm": "metallic", "map_ps": "sheen", "map_ke": "emissive",
                     "norm": "normal"}
_TINYOBJ_SCALARS = ("Ni", "illum", "d", "Tr", "Pm", "Ps", "Pc", "Pcr", "aniso", "anisor")
MATERIAL_CLASSES = {"diffuse": S.MAT_DIFFUSE, "translucent": S.MAT_TRANSLUCENT, "plastic": S.MAT_PLASTIC, "conductor": S.MAT_CONDUCTOR, "dielectric": S.MAT_DIELECTRIC,
                    "thinfilm": S.MAT_THINFILM, "mirror": S.MAT_MIRROR, "boundary": S.MAT_BOUNDARY, "velvet": S.MAT_VELVET, "principled": S.MAT_PRINCIPLED, "void": S.MAT_VOID}


# Image option bits (render/shared/image.hxx:15-25)
IMG_BUILD_TABLE, IMG_REPEAT_U, IMG_REPEAT_V, IMG_SKIP_SRGB, IMG_HAS_ALPHA, IMG_UNIFORM_TABLE, IMG_PERFORM_LOADING = 1, 2, 4, 8, 16, 32, 64
IMG_REPEAT = IMG_REPEAT_U | IMG_REPEAT_V


class LoaderError(RuntimeError):
    pass


class MtlBlock:
    def __init__(self, name):
        self.name = name
        self.params = []   # (key, value) in file order
        self.textures = {}

    def get(self, key):
        """get_param (scene_representation.cxx:102-112): first parameter whose key matches, case-insensitively."""
        k = key.lower()
        for pk, pv in self.params:
            if pk.lower() == k:
                return pv
        return None


_TEXTURE_OPTIONS = {"-blendu": 1, "-blendv": 1, "-clamp": 1, "-boost": 1, "-bm": 1, "-o": 3, "-s": 3, "-t": 3, "-type": 1, "-texres": 1, "-imfchan": 1, "-mm": 2, "-colorspace": 1}


def _texture_name(text):
    """ParseTextureNameAndOption (tiny_obj_loader.hxx:1242-1321): `-option args` are skipped by their argument COUNT, the first word that is not an option
    starts the file name, which runs to the end of the line — spaces included."""
    at, n = 0, len(text)
    while at < n:
        while at < n and text[at] in " \t":
            at += 1
        if at >= n:
            break
        for key, args in _TEXTURE_OPTIONS.items():
            k = len(key)
            if text[at:at + k].lower() == key and at + k < n and text[at + k] in " \t":
                at += k
                for _ in range(args):
                    while at < n and text[at] in " \t":
                        at += 1
                    while at < n and text[at] not in " \t\r":
                        at += 1
                break
        else:
            return text[at:]
    return None


def parse_mtl(path):
    blocks, cur = [], None
    with open(path, "r", errors="replace") as f:
        for raw in f:
            line = raw.rstrip("\n").rstrip("\r").rstrip(" \t")
            line = line.lstrip(" \t")
            if not line or line[0] == "#":
                continue
            low = line.lower()
            if low.startswith("newmtl") and len(line) > 6 and line[6] in " \t":
                cur = MtlBlock(line[7:].lower())
                blocks.append(cur)
                continue
            if cur is None:
                continue
            consumed = False
            for key, slot in _TINYOBJ_TEXTURES.items():
                if low.startswith(key) and len(line) > len(key) and line[len(key)] in " \t":
                    name = _texture_name(line[len(key) + 1:])
                    if name is not None:  # without a name the earlier value stays
                        cur.textures[slot] = name
                    consumed = True
                    break
            if consumed:
                continue
            for key in _TINYOBJ_SCALARS:
                if (low.startswith(key.lower()) if key in ("illum", "Pcr", "aniso", "anisor") else line.startswith(key)) and len(line) > len(key) and line[len(key)] in " \t":
                    consumed = True
                    break
            if consumed:
                continue
            sp = line.find(" ")
            if sp < 0:
                sp = line.find("\t")
            cur.params.append((line, "") if sp < 0 else (line[:sp], line[sp + 1:]))
    return blocks


def _floats(text, n=None):
    """sscanf("%f %f ...") semantics: as many leading floats as parse."""
    out = []
    for tok in text.split():
        try:
            out.append(float(tok))
        except ValueError:
            break
        if n is not None and len(out) == n:
            break
    return out


def _atof(tok):
    """C atof: the longest leading prefix that parses, else 0."""
    for end in range(len(tok), 0, -1):
        try:
            return float(tok[:end])
        except ValueError:
            continue
    return 0.0


def gamma_to_linear(v):
    """math.hxx gamma_to_linear on float32 values."""
    v = np.asarray(v, dtype=f32)
    lo = (v / f32(12.92)).astype(f32)
    hi = np.power(((v + f32(0.055)) / f32(1.055)).astype(f32), f32(2.4)).astype(f32)
    return np.where(v <= f32(0.04045), lo, hi).astype(f32)


# ---- spectra (render/host/spectrum.cxx) ---------------------------------------------------------------------------------------------
def _xyz_to_rgb(xyz):
    """spectrum::xyz_to_rgb (spectrum.hxx:142-148)."""
    x, y, z = (f32(v) for v in xyz)
    return np.array([f32(f32(f32(3.24045420) * x) - f32(f32(1.5371385) * y)) - f32(f32(0.4985314) * z),
                     f32(f32(f32(-0.9692660) * x) + f32(f32(1.8760108) * y)) + f32(f32(0.0415560) * z),
                     f32(f32(f32(0.05564340) * x) - f32(f32(0.2040259) * y)) + f32(f32(1.0572252) * z)], dtype=f32)


def integrate_to_xyz(power441):
    """SpectralDistribution::integrate_to_xyz (spectrum.cxx:348-377) on the 1 nm grid: per nanometre (v0 + (v1 - v0) / 2) of the CIE-weighted values,
    summed in float32 in the reference's order."""
    power = np.asarray(power441, dtype=f32)
    t = scenes.tables("color_tables")
    k = f32(1.0) / f32(t["y_integral"][0])
    v = (t["xyz_441x3"].astype(f32) * (power * k).astype(f32)[:, None]).astype(f32)
    v[power == 0] = 0.0
    seg = (v[:-1] + (f32(0.5) * (v[1:] - v[:-1]).astype(f32)).astype(f32)).astype(f32)
    xyz = np.zeros(3, dtype=f32)
    for row in seg:
        xyz = (xyz + row).astype(f32)
    return xyz


def spd_from_power(power441, integrated=None):
    """A distribution on the fixed 390..830 nm grid; `integrated` defaults to xyz_to_rgb(integrate_to_xyz) (spectrum.cxx:92-93)."""
    power = np.asarray(power441, dtype=f32)
    if integrated is None:
        integrated = _xyz_to_rgb(integrate_to_xyz(power))
    return scenes._spd(power, integrated)


_libm = None


def _expf(x):
    """the C library's expf, element by element: numpy's float32 exp differs from it in the last bit for some arguments, and the reference's black-body
    spectra (hence emitter weights and max_sigma of media built from them) are made with expf"""
    global _libm
    if _libm is None:
        import ctypes.util
        _libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _libm.expf.restype = C.c_float
        _libm.expf.argtypes = [C.c_float]
    x = np.asarray(x, dtype=f32)
    return np.array([_libm.expf(float(v)) for v in x.reshape(-1)], dtype=f32).reshape(x.shape)


def black_body_radiation(wavelength_nm, t_kelvins):
    """spectrum::black_body_radiation (render/shared/spectrum.hxx:171-189) in its float32 steps."""
    wl = (np.asarray(wavelength_nm, dtype=f32) * f32(1.0 / 1000.0)).astype(f32)
    wl5 = (wl * (wl * wl).astype(f32)).astype(f32) * (wl * wl).astype(f32)
    with np.errstate(over="ignore"):
        e0 = _expf((f32(1.4387752e+4) / (wl * f32(t_kelvins)).astype(f32)).astype(f32))
        d = (wl5.astype(f32) * (e0 - f32(1.0)).astype(f32)).astype(f32)
    return np.where(np.isinf(d), f32(0.0), (f32(3.7417712e+5) / d).astype(f32)).astype(f32)


def black_body_power(temperature, scale=1.0):
    return (black_body_radiation(scenes.WAVELENGTHS, temperature) * f32(scale)).astype(f32)


def spd_black_body(temperature, scale=1.0, normalized=False):
    """SpectralDistribution::from_black_body / from_normalized_black_body (spectrum.cxx:118-133)."""
    if not normalized:
        return spd_from_power(black_body_power(temperature, scale))
    w = f32(2.8977729e+6) / f32(temperature)
    r = black_body_radiation(np.array([w], dtype=f32), temperature)[0]
    spd = spd_from_power(black_body_power(temperature, f32(1.0) / r))
    lum = scenes.luminance(spd["integrated"][0])
    return spd_scaled(spd, f32(scale) / lum)


def spd_scaled(spd, factor):
    """SpectralDistribution::scale (spectrum.cxx:97-102)."""
    out = spd.copy()
    out["entries"]["power"][0] = (out["entries"]["power"][0] * f32(factor)).astype(f32)
    out["integrated"][0] = (out["integrated"][0] * f32(factor)).astype(f32)
    return out


def spd_from_samples(samples):
    """SpectralDistribution::from_samples (spectrum.cxx:11-95) in its float32 steps: the wavelength unit is scaled up until the first sample is >= 100
    (nm), samples are clamped to 390..830, sorted, made unique, and resampled onto the 1 nm grid with end values held."""
    if not samples:
        return scenes.spd_constant(0.0)
    mult = f32(1.0)
    while f32(samples[0][0]) * mult < f32(100.0):
        mult = f32(mult * f32(10.0))
    pts = []
    for w, p in samples:
        w, p = f32(f32(w) * mult), f32(p)
        if not (np.isfinite(w) and np.isfinite(p)):
            continue
        pts.append((min(max(w, f32(390.0)), f32(830.0)), p))
    if not pts:
        return scenes.spd_constant(0.0)
    pts.sort(key=lambda s: s[0])
    uniq = []
    for w, p in pts:
        if not uniq or abs(f32(w - uniq[-1][0])) > f32(1.0e-4):
            uniq.append([w, p])
        else:
            uniq[-1][1] = p
    if len(uniq) == 1:
        uniq.append(list(uniq[0]))
    power = np.zeros(441, dtype=f32)
    seg = 0
    for i in range(441):
        wl = f32(390 + i)
        if wl <= uniq[0][0]:
            power[i] = uniq[0][1]
        elif wl >= uniq[-1][0]:
            power[i] = uniq[-1][1]
        else:
            while seg + 1 < len(uniq) and uniq[seg + 1][0] < wl:
                seg += 1
            (x0, y0), (x1, y1) = uniq[seg], uniq[seg + 1]
            den = f32(x1 - x0)
            t = f32(f32(wl - x0) / den) if den > 0 else f32(0.0)
            t = min(max(t, f32(0.0)), f32(1.0))
            power[i] = f32(f32(y0 * f32(f32(1.0) - t)) + f32(y1 * t))
    return spd_from_power(power)


def subsurface_remap(color, distances):
    """subsurface::remap (render/shared/scene_bssrdf_subsurface.hxx:17-44) per channel, float32."""
    a, b, c = f32(1.826052378200), f32(f32(4.985111943850) + f32(0.12735595943800)), f32(1.096861024240)
    d, e, f = f32(0.496310210422), f32(f32(4.231902997010) + f32(0.00310603949088)), f32(2.406029994080)
    color = np.maximum(f32(0.0), np.asarray(color, dtype=f32))
    blend = np.power(color, f32(0.25)).astype(f32)
    albedo = ((f32(1.0) - blend).astype(f32) * a * np.power(np.arctan((b * color).astype(f32)).astype(f32), c).astype(f32)).astype(f32)
    albedo = (albedo + (blend * d * np.power(np.arctan((e * color).astype(f32)).astype(f32), f).astype(f32)).astype(f32)).astype(f32)
    albedo = np.clip(albedo, f32(0.0), f32(1.0) - f32(1.192092896e-07)).astype(f32)  # kEpsilon (math.hxx:107)
    extinction = (f32(1.0) / np.maximum(np.asarray(distances, dtype=f32), f32(1.0 / 1024.0))).astype(f32)
    return albedo, extinction, (extinction * albedo).astype(f32)


# ---- image files ----------------------------------------------------------------------------------------------------------------------
def _read_png(path):
    d = open(path, "rb").read()
    if d[:8] != b"\x89PNG\r\n\x1a\n":
        raise LoaderError(f"{path}: not a PNG file")
    pos, idat, hdr, pal = 8, b"", None, None
    while pos < len(d):
        n, = struct.unpack(">I", d[pos:pos + 4])
        typ, body = d[pos + 4:pos + 8], d[pos + 8:pos + 8 + n]
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"PLTE":
            pal = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif typ == b"tRNS":
            raise LoaderError(f"{path}: a PNG with a tRNS chunk is left to the module's reader")
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, ctype, _, _, interlace = hdr
    if depth != 8 or interlace != 0:
        raise LoaderError(f"{path}: only 8-bit non-interlaced PNG files are read")
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * ch)
    out = np.zeros((h, w * ch), np.uint8)
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:
            cur = np.zeros_like(line)
            for x in range(w * ch):
                a = cur[x - ch] if x >= ch else 0
                b = prev[x]
                c = prev[x - ch] if x >= ch else 0
                if ft == 1:
                    p = a
                elif ft == 3:
                    p = (a + b) >> 1
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[x] = (line[x] + p) & 255
        out[y] = cur
        prev = cur
    px = out.reshape(h, w, ch)
    if ctype == 3:
        px = pal[px[..., 0]]
    rgba = np.full((h, w, 4), 255, np.uint8)
    if px.shape[2] == 1:
        rgba[..., :3] = px
    elif px.shape[2] == 2:
        rgba[...] = 0  # grey + alpha: the reference's switch over the channel count has no case for 2 (image_pool.cxx:353-381), the image stays zero-filled
    else:
        rgba[..., :px.shape[2]] = px
    return rgba


def _read_exr(path):
    d = open(path, "rb").read()
    if struct.unpack_from("<I", d, 0)[0] != 20000630:
        raise LoaderError(f"{path}: not an OpenEXR file")
    pos, attrs = 8, {}
    while d[pos] != 0:
        ne = d.index(b"\0", pos)
        te = d.index(b"\0", ne + 1)
        size, = struct.unpack_from("<I", d, te + 1)
        attrs[d[pos:ne].decode()] = d[te + 5:te + 5 + size]
        pos = te + 5 + size
    pos += 1
    comp = attrs["compression"][0]
    if comp not in (0, 2, 3):
        raise LoaderError(f"{path}: EXR compression {comp} is not read here (none / ZIPS / ZIP are)")
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    names, types, cp, ch = [], [], 0, attrs["channels"]
    while ch[cp] != 0:
        e = ch.index(b"\0", cp)
        names.append(ch[cp:e].decode())
        types.append(struct.unpack_from("<i", ch, e + 1)[0])
        cp = e + 17
    lines = {0: 1, 2: 1, 3: 16}[comp]
    blocks = (h + lines - 1) // lines
    offsets = struct.unpack_from(f"<{blocks}Q", d, pos)
    bpp = [2 if t == 1 else 4 for t in types]
    out = np.zeros((h, w, 4), f32)
    out[..., 3] = 1.0
    for off in offsets:
        by, size = struct.unpack_from("<iI", d, off)
        nl = min(lines, y1 - by + 1)
        raw = d[off + 8:off + 8 + size]
        want = nl * w * sum(bpp)
        if comp and size < want:
            z = np.frombuffer(zlib.decompress(raw), np.uint8).astype(np.int32)
            z = np.cumsum(np.concatenate([[z[0]], z[1:] - 128])) & 255  # predictor
            half = (len(z) + 1) // 2
            un = np.zeros(len(z), np.uint8)
            un[0::2] = z[:half]
            un[1::2] = z[half:]
            raw = un.tobytes()
        p = 0
        for ly in range(nl):
            for ci, name in enumerate(names):
                dt = np.float16 if types[ci] == 1 else (np.float32 if types[ci] == 2 else np.uint32)
                vals = np.frombuffer(raw, dt, w, p).astype(f32)
                p += w * bpp[ci]
                if name in "RGBA":
                    out[by - y0 + ly, :, "RGBA".index(name)] = vals
                elif name == "Y":
                    out[by - y0 + ly, :, :3] = vals[:, None]
    return out


def _read_pfm(path):
    """load_pfm (render/host/image_pool.cxx:463-541): the reference's own variant — `Pf` / `PF`, then width, height and scale each on a line of its own,
    then float rows in file order; the scale (and with it the byte order) is ignored."""
    d = open(path, "rb").read()
    pos, lines = 0, []
    for _ in range(4):
        e = d.find(b"\n", pos)
        if e < 0 or e - pos > 16:
            raise LoaderError(f"{path}: not a PFM file the reference reads")
        lines.append(d[pos:e].decode("ascii", "replace"))
        pos = e + 1
    fmt = lines[0][1:2]
    w, h = int(lines[1].split()[0]), int(lines[2].split()[0])
    float(lines[3].split()[0])
    ch = {"f": 1, "F": 3}.get(fmt)
    if ch is None:
        raise LoaderError(f"{path}: PFM format P{fmt} is not read")
    px = np.frombuffer(d, np.float32, w * h * ch, pos).reshape(h, w, ch)
    out = np.ones((h, w, 4), dtype=f32)
    out[..., :3] = px if ch == 3 else px[..., :1]
    return out


def _read_hdr(path):
    """Radiance RGBE as stb_image's stbi_loadf reads it (flat or run-length encoded scan lines; value = mantissa * 2^(exponent - 136))."""
    d = open(path, "rb").read()
    if not (d.startswith(b"#?RADIANCE") or d.startswith(b"#?RGBE")):
        raise LoaderError(f"{path}: not a Radiance HDR file")
    pos = d.index(b"\n\n") + 2
    e = d.index(b"\n", pos)
    tok = d[pos:e].split()
    if len(tok) != 4 or tok[0] != b"-Y" or tok[2] != b"+X":
        raise LoaderError(f"{path}: unsupported HDR orientation {d[pos:e]!r}")
    h, w = int(tok[1]), int(tok[3])
    pos = e + 1
    rgbe = np.zeros((h, w, 4), np.uint8)
    rle = 8 <= w < 32768 and d[pos] == 2 and d[pos + 1] == 2 and (d[pos + 2] & 0x80) == 0  # decided once, at the first scan line, like stb_image
    if not rle:
        rgbe[:] = np.frombuffer(d, np.uint8, w * h * 4, pos).reshape(h, w, 4)
    else:
        for y in range(h):
            if d[pos] != 2 or d[pos + 1] != 2 or ((d[pos + 2] << 8) | d[pos + 3]) != w:
                raise LoaderError(f"{path}: corrupt HDR scan line")
            pos += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = d[pos]
                    pos += 1
                    if n > 128:
                        n -= 128
                        rgbe[y, x:x + n, c] = d[pos]
                        pos += 1
                    else:
                        rgbe[y, x:x + n, c] = np.frombuffer(d, np.uint8, n, pos)
                        pos += n
                    x += n
    scale = np.ldexp(f32(1.0), rgbe[..., 3].astype(np.int32) - 136).astype(f32)
    out = np.ones((h, w, 4), dtype=f32)
    out[..., :3] = np.where(rgbe[..., 3:4] != 0, (rgbe[..., :3].astype(f32) * scale[..., None]).astype(f32), f32(0.0))
    return out


def read_image(path):
    """-> float32 RGBA (linear values as stored for .exr, raw 0..1 for 8-bit files; the caller applies the sRGB curve)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".png":
        try:
            return _read_png(path), True
        except LoaderError:
            # bit depths other than 8, Adam7, colour keys: the module's reader covers every PNG form the reference's does (etxb_image_file_read)
            from . import api
            try:
                return api.read_image(path), True
            except api.EtxbError as e:
                raise LoaderError(str(e))
    if ext == ".exr":
        try:
            return _read_exr(path), False
        except LoaderError:
            from . import api  # run-length encoded and PIZ blocks: the module's reader
            try:
                return api.read_image(path), False
            except api.EtxbError as e:
                raise LoaderError(str(e))
    if ext == ".hdr":
        return _read_hdr(path), False
    if ext == ".pfm":
        return _read_pfm(path), False
    # every other file goes to stb_image in the reference, which looks at the content: JPEG (and PNG under another name) through the module's readers
    from . import api
    try:
        return api.read_image(path), True
    except api.EtxbError as e:
        raise LoaderError(str(e))


# ---- the loader -----------------------------------------------------------------------------------------------------------------------------
def _orthonormal_basis(n):
    """math.hxx:737-746 on an (N, 3) float32 array."""
    x, y, z = n[:, 0], n[:, 1], n[:, 2]
    general = (x != y) | (x != z)
    a = np.where(general[:, None], np.stack([z - y, x - z, y - x], 1), np.stack([z - y, x + z, -y - x], 1)).astype(f32)
    a = _normalize(a)
    b = _normalize(_cross(n, a))
    return a, b


def _dot(a, b):
    return ((a[:, 0] * b[:, 0]).astype(f32) + (a[:, 1] * b[:, 1]).astype(f32)).astype(f32) + (a[:, 2] * b[:, 2]).astype(f32)


def _cross(a, b):
    return np.stack([(a[:, 1] * b[:, 2]).astype(f32) - (b[:, 1] * a[:, 2]).astype(f32), (a[:, 2] * b[:, 0]).astype(f32) - (b[:, 2] * a[:, 0]).astype(f32),
                     (a[:, 0] * b[:, 1]).astype(f32) - (b[:, 0] * a[:, 1]).astype(f32)], 1).astype(f32)


def _normalize(v):
    with np.errstate(invalid="ignore", divide="ignore"):
        return (v / np.sqrt(_dot(v, v)).astype(f32)[:, None]).astype(f32)


def _valid(v):
    return np.isfinite(v).all(axis=1) & (_dot(v, v) > 0)


def _clip_ears(q):
    """tinyobjloader's triangulation of a polygon with more than four corners (tiny_obj_loader.hxx:1537-1800): ears clipped in the plane of the two
    dominant axes, float32 arithmetic.  `q` = corner positions; returns index triples into the polygon."""
    n, axes, eps = len(q), [1, 2], np.finfo(f32).eps
    for k in range(n):
        a, b, c = q[k % n], q[(k + 1) % n], q[(k + 2) % n]
        e0, e1 = (b - a).astype(f32), (c - b).astype(f32)
        cx = abs(f32(f32(e0[1] * e1[2]) - f32(e0[2] * e1[1])))
        cy = abs(f32(f32(e0[2] * e1[0]) - f32(e0[0] * e1[2])))
        cz = abs(f32(f32(e0[0] * e1[1]) - f32(e0[1] * e1[0])))
        if cx > eps or cy > eps or cz > eps:
            if not (cx > cy and cx > cz):
                axes[0] = 0
                if cz > cx and cz > cy:
                    axes[1] = 1
            break
    rest, out, guess, budget, previous = list(range(n)), [], 0, n, n
    while len(rest) > 3 and budget > 0:
        m = len(rest)
        if guess >= m:
            guess -= m
        if previous != m:
            previous, budget = m, m
        else:
            budget -= 1
        ind = [rest[(guess + k) % m] for k in range(3)]
        vx = [f32(q[i][axes[0]]) for i in ind]
        vy = [f32(q[i][axes[1]]) for i in ind]
        e0x, e0y, e1x, e1y = f32(vx[1] - vx[0]), f32(vy[1] - vy[0]), f32(vx[2] - vx[1]), f32(vy[2] - vy[1])
        turn = f32(f32(e0x * e1y) - f32(e0y * e1x))
        area = f32(f32(f32(vx[0] * vy[1]) - f32(vy[0] * vx[1])) * f32(0.5))
        if f32(turn * area) < 0:
            guess += 1
            continue
        overlap = False
        for other in range(3, m):
            t = q[rest[(guess + other) % m]]
            tx, ty = f32(t[axes[0]]), f32(t[axes[1]])
            inside, j = False, 2
            for i in range(3):
                if (vy[i] > ty) != (vy[j] > ty):
                    with np.errstate(divide="ignore", invalid="ignore"):
                        x = f32(f32(f32(f32(vx[j] - vx[i]) * f32(ty - vy[i])) / f32(vy[j] - vy[i])) + vx[i])
                    if tx < x:
                        inside = not inside
                j = i
            if inside:
                overlap = True
                break
        if overlap:
            guess += 1
            continue
        out.append(tuple(ind))
        del rest[(guess + 1) % m]
    if len(rest) == 3:
        out.append(tuple(rest))
    return tuple(out)


class _ObjData:
    pass


def parse_obj(path):
    """Positions / normals / texcoords, triangulated faces (quads along the shorter diagonal, larger polygons by ear clipping, like tinyobjloader), the material name and the shape (o / g group) of every face, and the mtllib."""
    pos, nrm, tex = [], [], []
    faces, face_mtl, face_shape = [], [], []
    mtllib, cur_mtl, shape, shape_has_faces = None, None, 0, False
    with open(path, "r", errors="replace") as f:
        for line in f:
            if not line or line[0] == "#":
                continue
            t = line.split()
            if not t:
                continue
            k = t[0]
            if k == "v":
                pos.append((float(t[1]), float(t[2]), float(t[3])))
            elif k == "vn":
                nrm.append((float(t[1]), float(t[2]), float(t[3])))
            elif k == "vt":
                tex.append((float(t[1]), float(t[2]) if len(t) > 2 else 0.0))
            elif k == "f":
                idx = []
                for c in t[1:]:
                    p = c.split("/")
                    vi = int(p[0])
                    ti = int(p[1]) if len(p) > 1 and p[1] else 0
                    ni = int(p[2]) if len(p) > 2 and p[2] else 0
                    idx.append((vi - 1 if vi > 0 else len(pos) + vi, (ti - 1 if ti > 0 else len(tex) + ti) if ti else -1, (ni - 1 if ni > 0 else len(nrm) + ni) if ni else -1))
                if len(idx) == 4:
                    # tinyobjloader splits a quad along its shorter diagonal (tiny_obj_loader.hxx:1464-1527), in float32
                    q = [np.asarray(pos[i[0]], dtype=f32) for i in idx]
                    e02, e13 = (q[2] - q[0]).astype(f32), (q[3] - q[1]).astype(f32)
                    s02 = f32(f32(f32(e02[0] * e02[0]) + f32(e02[1] * e02[1])) + f32(e02[2] * e02[2]))
                    s13 = f32(f32(f32(e13[0] * e13[0]) + f32(e13[1] * e13[1])) + f32(e13[2] * e13[2]))
                    tris = ((0, 1, 2), (0, 2, 3)) if s02 < s13 else ((0, 1, 3), (1, 2, 3))
                elif len(idx) == 3:
                    tris = ((0, 1, 2),)
                elif len(idx) > 4:
                    tris = _clip_ears([np.asarray(pos[i[0]], dtype=f32) for i in idx])
                else:
                    tris = ()
                for (a, b, c) in tris:
                    faces.append((idx[a], idx[b], idx[c]))
                    face_mtl.append(cur_mtl)
                    face_shape.append(shape)
                shape_has_faces = True
            elif k == "usemtl":
                cur_mtl = line.split(None, 1)[1].strip().lower() if len(t) > 1 else None
            elif k in ("o", "g"):
                if shape_has_faces:
                    shape += 1
                    shape_has_faces = False
            elif k == "mtllib" and len(t) > 1:
                mtllib = line.split(None, 1)[1].strip()
    o = _ObjData()
    o.pos = np.asarray(pos, dtype=f32).reshape(-1, 3)
    o.nrm = np.asarray(nrm, dtype=f32).reshape(-1, 3)
    o.tex = np.asarray(tex, dtype=f32).reshape(-1, 2)
    o.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3, 3)
    o.face_mtl, o.face_shape, o.mtllib = face_mtl, np.asarray(face_shape, dtype=np.int64), mtllib
    return o


class SceneLoader:
    def __init__(self, data_folder=None):
        self.sd = scenes.SceneData()
        self.sd._images, self.sd._mediums, self.sd._distant_emitters = [], [], []
        self.named_spectra = {}
        self.medium_names = {}
        self.image_cache = {}
        self.cameras = []
        self.base_dir = ""
        self.data_folder = data_folder
        self.warnings = []
        self._init_default_values()

    # -- bookkeeping --------------------------------------------------------------------------------
    def warn(self, text):
        self.warnings.append(text)

    def _init_default_values(self):
        """SceneRepresentationImpl::init_default_values (:206-225)."""
        sd, sc = self.sd, {}
        sc["black_spectrum"] = sd.add_spectrum(scenes.spd_rgb_reflectance([0.0, 0.0, 0.0]))
        sc["white_spectrum"] = sd.add_spectrum(scenes.spd_rgb_reflectance([1.0, 1.0, 1.0]))
        atm = scenes.tables("spectra")
        for key in ("rayleigh", "mie", "ozone"):
            sc[key + "_spectrum"] = sd.add_spectrum(scenes._spd(atm[f"atmosphere_{key}.power"], atm[f"atmosphere_{key}.rgb"]))
        sc["default_dielectric_eta"] = sd.add_spectrum(scenes.spd_constant(1.5))
        sc["default_conductor_eta"] = sd.add_spectrum(scenes.spd_constant(0.0))
        sc["default_conductor_k"] = sd.add_spectrum(scenes.spd_constant(1000000.0))
        a = sd.add_material("etx::subsurface-scatter", cls=S.MAT_TRANSLUCENT)
        sd.materials[a]["reflectance"]["spectrum_index"] = sc["black_spectrum"]
        sd.materials[a]["scattering"]["spectrum_index"] = sc["white_spectrum"]
        b = sd.add_material("etx::subsurface-exit", cls=S.MAT_DIFFUSE)
        sd.materials[b]["reflectance"]["spectrum_index"] = sc["white_spectrum"]
        sd.materials[b]["scattering"]["spectrum_index"] = sc["white_spectrum"]
        sc["subsurface_scatter_material"], sc["subsurface_exit_material"] = a, b
        self.defaults = sc

    def add_named_spectrum(self, name, spd):
        i = self.sd.add_spectrum(spd)
        self.named_spectra[name] = i
        return i

    def find_file(self, name):
        """get_file (:114-121): "<folder of the material file>/<name>" for any non-empty name — the file is not looked for here; one that cannot be read
        becomes the 1 x 1 white placeholder in add_image_file, like in the reference's texture pool."""
        if not name:
            return None
        return (self.base_dir + "/" + name) if self.base_dir else name

    def add_image_file(self, path, options, offset=(0.0, 0.0), scale=(1.0, 1.0)):
        """ImagePool::add_from_file + load_image (image_pool.cxx:51-66, 162-215): 8-bit files stay RGBA8 with the sRGB curve removed and re-quantised
        (unless SkipSRGBConversion), float files are RGBA32F; rows in file order; a file that cannot be read becomes the 1 x 1 white placeholder."""
        key = os.path.abspath(path)
        if key in self.image_cache:
            return self.image_cache[key]
        options |= IMG_PERFORM_LOADING
        try:
            px, eight_bit = read_image(path)
        except (LoaderError, OSError, ValueError, KeyError) as e:
            self.warn(f"{path}: {e}; using the 1x1 white placeholder")
            px, eight_bit = np.ones((1, 1, 4), dtype=f32), False
            options |= IMG_SKIP_SRGB | IMG_REPEAT_U | IMG_REPEAT_V
        if eight_bit:
            v = (px.astype(f32) / f32(255.0)).astype(f32)
            if not (options & IMG_SKIP_SRGB):
                v[..., :3] = gamma_to_linear(v[..., :3])
            px = (np.clip(v, 0.0, 1.0).astype(f32) * f32(255.0)).astype(f32).astype(np.uint8)  # to_ubyte4 truncates (math.hxx:713-720)
        else:
            px = np.where(np.isinf(px), f32(65504.0), px)
            px = np.where(np.isnan(px) | (px < 0), f32(0.0), px).astype(f32)
        i = self.sd.add_image(px, repeat=bool(options & IMG_REPEAT_U), repeat_v=bool(options & IMG_REPEAT_V), build_table=bool(options & IMG_BUILD_TABLE),
                              uniform_table=bool(options & IMG_UNIFORM_TABLE), offset=offset, scale=scale)
        rec = self.sd._images[i]
        rec["options"] = options | (int(rec["options"][0]) & IMG_HAS_ALPHA)
        self.image_cache[key] = i
        return i

    # -- spectra directives ---------------------------------------------------------------------------
    def reflectance_spectrum(self, text):
        """load_reflectance_spectrum (:1613-1633)."""
        p = _split_params(text)
        if len(p) == 1 and p[0] in self.named_spectra:
            return self.named_spectra[p[0]]
        if len(p) == 3:
            return self.sd.add_spectrum(scenes.spd_rgb_reflectance(gamma_to_linear([_atof(v) for v in p])))
        return 0

    def illuminant_spectrum(self, text):
        """load_illuminant_spectrum (:1635-1680)."""
        p = _split_params(text)
        if len(p) == 1:
            fl = _floats(p[0], 1)
            if fl:
                return scenes.spd_rgb_luminance([fl[0]] * 3)
            if p[0] in self.named_spectra:
                return self.sd.spectra[self.named_spectra[p[0]]].copy()
        if len(p) == 3:
            return scenes.spd_rgb_luminance([_atof(v) for v in p])
        spd, scale, i = scenes.spd_rgb_luminance([1.0, 1.0, 1.0]), 1.0, 0
        while i < len(p):
            if p[i] == "blackbody" and i + 1 < len(p):
                spd = spd_black_body(_atof(p[i + 1]), 1.0)
                i += 1
            elif p[i] == "nblackbody" and i + 1 < len(p):
                spd = spd_black_body(_atof(p[i + 1]), 1.0, normalized=True)
                i += 1
            elif p[i] == "scale" and i + 1 < len(p):
                scale = _atof(p[i + 1])
                i += 1
            i += 1
        return spd_scaled(spd, scale)

    def load_ior(self, target, text):
        """the load_ior lambda of parse_material (:1846-1884)."""
        sd = self.sd
        fl = _floats(text, 2)
        if len(fl) == 1:
            target["cls"] = S.SPD_DIELECTRIC
            target["eta_index"] = sd.add_spectrum(scenes.spd_constant(fl[0]))
            target["k_index"] = S.INVALID
        elif len(fl) == 2:
            target["cls"] = S.SPD_CONDUCTOR
            target["eta_index"] = sd.add_spectrum(scenes.spd_constant(fl[0]))
            target["k_index"] = sd.add_spectrum(scenes.spd_constant(fl[1]))
        else:
            name = text.strip().lower()
            try:
                eta, k, cls = scenes.spd_named_ior(name)
            except KeyError:
                self.warn(f"unable to load IOR spectrum `{text}`, falling back to 1.5 dielectric")
                eta, k, cls = scenes.spd_constant(1.5), scenes.spd_constant(0.0), S.SPD_DIELECTRIC
            target["cls"] = cls
            target["eta_index"] = sd.add_spectrum(eta)
            target["k_index"] = sd.add_spectrum(k)

    # -- et:: blocks ------------------------------------------------------------------------------------
    def parse_camera(self, b):
        cam = dict(cls=0, viewport=(0, 0), origin=None, target=None, up=(0.0, 1.0, 0.0), fov=50.0, lens_radius=0.0, focal_distance=0.0, clip_near=None, clip_far=None,
                   lens_image=S.INVALID, medium=S.INVALID, id="", active=False)
        v = b.get("class")
        if v is not None:
            cam["cls"] = 1 if v.strip() == "eq" else 0
        v = b.get("viewport")
        if v is not None and len(v.split()) >= 2:
            cam["viewport"] = (int(v.split()[0]), int(v.split()[1]))
        for key in ("origin", "target", "up"):
            v = b.get(key)
            if v is not None and len(_floats(v, 3)) == 3:
                cam[key] = tuple(_floats(v, 3))
        v = b.get("fov")
        if v is not None and _floats(v, 1):
            cam["fov"] = _floats(v, 1)[0]
        v = b.get("focal-length")
        if v is not None and _floats(v, 1):
            cam["fov"] = float(f32(focal_length_to_fov(_floats(v, 1)[0])) * f32(180.0) / f32(math.pi))
        for key, fld in (("lens-radius", "lens_radius"), ("focal-distance", "focal_distance"), ("clip-near", "clip_near"), ("clip-far", "clip_far")):
            v = b.get(key)
            if v is not None and _floats(v, 1):
                cam[fld] = _floats(v, 1)[0]
        v = b.get("shape")
        if v is not None:
            f = self.find_file(v.strip())
            if f:
                cam["lens_image"] = self.add_image_file(f, IMG_BUILD_TABLE | IMG_UNIFORM_TABLE)
        v = b.get("ext_medium")
        if v is not None:
            cam["medium"] = self.medium_names.get(v.strip(), S.INVALID)
        v = b.get("id")
        if v is not None:
            cam["id"] = v.strip()
        v = b.get("active")
        if v is not None:
            cam["active"] = bool(int(_atof(v.split()[0]))) if v.split() else False
        self.cameras.append(cam)

    def parse_medium(self, b):
        name = b.get("id")
        if name is None:
            self.warn("medium does not have identifier - skipped")
            return
        name = name.strip()
        g = 0.0
        for key in ("g", "anisotropy"):
            v = b.get(key)
            if v is not None and _floats(v, 1):
                g = _floats(v, 1)[0]

        def rgb(text):
            fl = _floats(text, 3)
            if len(fl) == 3:
                return fl
            if len(fl) >= 1:
                return [fl[0]] * 3
            return None
        s_a, s_t = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
        spd_a = spd_t = None
        for key in ("absorption", "absorbtion"):
            v = b.get(key)
            if v is not None and rgb(v) is not None:
                s_a = rgb(v)
        v = b.get("scattering")
        if v is not None and rgb(v) is not None:
            s_t = rgb(v)
        for key in ("rayleigh", "mie"):
            v = b.get(key)
            if v is not None:
                base = self.sd.spectra[self.defaults[key + "_spectrum"]]
                scale, p = 1.0, _split_params(v)
                for i, tok in enumerate(p):
                    if tok == "scale" and i + 1 < len(p):
                        scale = _atof(p[i + 1])
                spd_t = spd_scaled(base, f32(scale) / f32(base["entries"]["power"][0].max()))
        v = b.get("parametric")
        if v is not None:
            # colour + scattering distances -> absorption / scattering through subsurface::remap (scene_bssrdf_subsurface.hxx:17-44), :1254-1296
            color, dist, scale = [1.0, 1.0, 1.0], [0.25, 0.25, 0.25], 1.0
            p, i = _split_params(v), 0
            while i < len(p):
                if p[i] == "color" and i + 3 < len(p):
                    color = [_atof(p[i + 1]), _atof(p[i + 2]), _atof(p[i + 3])]
                    i += 3
                if i < len(p) and p[i] == "distance" and i + 1 < len(p):
                    dist = [_atof(p[i + 1])] * 3
                    i += 1
                if i < len(p) and p[i] == "distances" and i + 3 < len(p):
                    dist = [_atof(p[i + 1]), _atof(p[i + 2]), _atof(p[i + 3])]
                    i += 3
                if i < len(p) and p[i] == "scale" and i + 1 < len(p):
                    scale = _atof(p[i + 1])
                    i += 1
                i += 1
            albedo, extinction, scattering = subsurface_remap(np.asarray(color, dtype=f32), (f32(scale) * np.asarray(dist, dtype=f32)).astype(f32))
            s_t = [float(x) for x in scattering]
            s_a = [float(x) for x in np.maximum(f32(0.0), (extinction - scattering).astype(f32))]
        explicit = b.get("enclosed") is None
        v = b.get("volume")
        m = self.sd.add_medium(absorption=s_a, scattering=s_t, g=g, explicit_connections=explicit)
        rec = self.sd._mediums[m]
        if v is not None and v.strip():
            # MediumPool::add (medium_pool.cxx:41-59) on the dense grid of the module's NanoVDB reader (etxb_nvdb_density, csrc/scene_loader_nvdb.inl)
            from . import api
            file = os.path.join(self.base_dir, v.strip())
            if not file.lower().endswith(".nvdb"):
                raise LoaderError(f"{file}: only NanoVDB (.nvdb) volume files are read")
            lib, dims, err = api.load_library("fast"), np.zeros(3, np.uint32), C.create_string_buffer(512)
            if lib.etxb_nvdb_density(os.fsencode(file), dims.ctypes.data, None, 0, err, len(err)) != 0:
                raise LoaderError(err.value.decode(errors="replace"))
            if int(dims.prod()) > 0:
                grid = np.zeros(int(dims.prod()), f32)
                if lib.etxb_nvdb_density(os.fsencode(file), dims.ctypes.data, grid.ctypes.data, grid.size, err, len(err)) != 0:
                    raise LoaderError(err.value.decode(errors="replace"))
                top = f32(grid.max())
                if top > 0:
                    grid = (grid / top).astype(f32)
                    rec["cls"] = 1
                    rec["density"]["a"] = grid.ctypes.data
                    rec["density"]["count"] = grid.size
                    rec["dimensions"][0] = dims
                    self.sd._keep.append(grid)
        if spd_t is not None:
            self.sd.spectra[int(rec["scattering_index"][0])] = spd_t
        # max_sigma = the two spectra's maximum powers added up (scene_data.hxx:137-143)
        rec["max_sigma"] = f32(self.sd.spectra[int(rec["absorption_index"][0])]["entries"]["power"][0].max()) + f32(self.sd.spectra[int(rec["scattering_index"][0])]["entries"]["power"][0].max())
        self.medium_names[name] = m

    def parse_directional(self, b):
        v = b.get("color")
        spd = self.illuminant_spectrum(v) if v is not None else scenes.spd_rgb_luminance([1.0, 1.0, 1.0])
        d = [1.0, 1.0, 1.0]
        v = b.get("direction")
        if v is not None and len(_floats(v, 3)) == 3:
            d = _floats(v, 3)
        ang = 0.0
        v = b.get("angular_diameter")
        if v is not None and _floats(v, 1):
            ang = _floats(v, 1)[0]
        self.sd.add_directional_emitter(d, [1.0, 1.0, 1.0], ang)
        p, _ = self.sd._distant_emitters[-1]
        self.sd.spectra[int(p["emission"]["spectrum_index"][0])] = spd
        p["angular_size"] = f32(f32(ang) * f32(math.pi) / f32(180.0))
        p["equivalent_disk_size"] = f32(2.0) * f32(math.tan(float(p["angular_size"][0]) / 2.0))
        p["angular_size_cosine"] = f32(math.cos(float(p["angular_size"][0]) / 2.0))
        v = b.get("image")
        if v is not None:
            f = self.find_file(v.strip())
            if f:
                p["emission"]["image_index"] = self.add_image_file(f, 0)

    def parse_env(self, b):
        v = b.get("image")
        image = S.INVALID
        rotation, u_scale = 0.0, 1.0
        r = b.get("rotation")
        if r is not None:
            rotation = float(-f32(_atof(r.split()[0] if r.split() else "0")) / f32(360.0))
        s = b.get("scale")
        if s is not None and _floats(s, 1):
            u_scale = _floats(s, 1)[0]
        name = os.path.join(self.base_dir, v.strip()) if (v is not None and v.strip()) else os.path.join(self.base_dir, f"image-{len(self.sd._images)}")
        # a missing / unnamed image is the 1 x 1 white placeholder: a constant-colour environment (scene_data.hxx:104-107, image_pool.cxx:172-184)
        image = self.add_image_file(name, IMG_BUILD_TABLE | IMG_REPEAT_U, offset=(rotation, 0.0), scale=(u_scale, 1.0))
        c = b.get("color")
        spd = self.illuminant_spectrum(c) if c is not None else scenes.spd_rgb_luminance([1.0, 1.0, 1.0])
        self.sd.add_environment_emitter(image)
        p, _ = self.sd._distant_emitters[-1]
        self.sd.spectra[int(p["emission"]["spectrum_index"][0])] = spd

    def parse_atmosphere(self, b):
        """parse_atmosphere_light (:1376-1495): a sun (Directional, 128 x 128 limb-darkened extinction image) and a sky (Environment, single-scattering
        image with its sampling table).  The two images come from the module's host code (etxb_atmosphere_images, csrc/scene_loader_atmosphere.inl:
        render/host/scattering.cxx restated) — the C++ loader builds the same block from the same function."""
        from . import api

        def scalar(key, fallback):
            v = b.get(key)
            return _floats(v, 1)[0] if (v is not None and _floats(v, 1)) else fallback
        quality, scale, sun_scale, sky_scale = scalar("quality", 1.0), scalar("scale", 1.0), scalar("sun_scale", 1.0), scalar("sky_scale", 1.0)
        d = np.array([1.0, 1.0, 1.0], f32)
        v = b.get("direction")
        if v is not None and len(_floats(v, 3)) == 3:
            d = np.array(_floats(v, 3), f32)
        d = (d / np.sqrt((d * d).sum(dtype=f32), dtype=f32)).astype(f32)
        ang_deg = scalar("angular_diameter", 0.5422)
        ang = f32(ang_deg) * (f32(math.pi) / f32(180.0))
        prm = np.array([scalar("altitude", 1000.0), scalar("anisotropy", 0.825), scalar("rayleigh", 1.0), scalar("mie", 1.0), scalar("ozone", 1.0)], f32)
        radiance_scale = f32(scale) * (f32(2.0 * f32(math.pi)) * (f32(1.0) - f32(np.cos(f32(0.5) * ang, dtype=f32))))
        sun_spectrum = spd_black_body(5900.0, radiance_scale)
        sky_w, sky_h = max(64, int(f32(2048) * f32(quality))), max(64, int(f32(1024) * f32(quality)))
        sun = np.zeros((128, 128, 4), f32)
        sky = np.zeros((sky_h, sky_w, 4), f32)
        lib = api.load_library("fast")
        rc = lib.etxb_atmosphere_images(None, d.ctypes.data, C.c_float(float(ang)), prm.ctypes.data, sky_w, sky_h,
                                        sun.ctypes.data if ang > 0 else None, sky.ctypes.data)
        if rc != 0:
            raise LoaderError(f"etxb_atmosphere_images failed ({rc})")
        self.sd.add_directional_emitter(d, [1.0, 1.0, 1.0], 0.0)
        p, _ = self.sd._distant_emitters[-1]
        self.sd.spectra[int(p["emission"]["spectrum_index"][0])] = spd_scaled(sun_spectrum, sun_scale)
        p["direction"][0] = d
        p["angular_size"] = ang
        p["equivalent_disk_size"] = f32(2.0) * f32(math.tan(float(ang) / 2.0))
        p["angular_size_cosine"] = f32(math.cos(float(ang) / 2.0))
        if ang > 0:
            p["emission"]["image_index"] = self.sd.add_image(sun, repeat=False, build_table=False)
        image = self.sd.add_image(sky, repeat=False, build_table=True)
        self.sd.add_environment_emitter(image)
        p, _ = self.sd._distant_emitters[-1]
        self.sd.spectra[int(p["emission"]["spectrum_index"][0])] = spd_scaled(sun_spectrum, sky_scale)
        p["direction"][0] = d

    def add_default_atmosphere(self):
        """load_from_file :805-820: what a scene file that declares no distant emitter gets."""
        b = MtlBlock("et::atmosphere")
        b.params = [("direction", "0.0 2.0 1.0"), ("quality", "0.125"), ("angular_diameter", "0.5422"), ("anisotropy", "0.825"), ("altitude", "1000.0"), ("scale", "1.0"),
                    ("sky_scale", "1.0"), ("sun_scale", "1.0"), ("rayleigh", "1.0"), ("mie", "1.0"), ("ozone", "1.0")]
        self.parse_atmosphere(b)

    def parse_spectrum(self, b):
        name = b.get("id")
        if name is None:
            self.warn("spectrum does not have identifier - skipped")
            return
        name = name.strip()
        scale = _atof(b.get("scale").split()[0]) if b.get("scale") and b.get("scale").split() else 1.0
        illuminant = b.get("illuminant") is not None
        spd = None
        if b.get("rgb") is not None:
            p = _split_params(b.get("rgb"))
            if len(p) < 3:
                return
            value = gamma_to_linear([_atof(p[0]), _atof(p[1]), _atof(p[2])])
            spd = scenes.spd_rgb_luminance(value) if illuminant else scenes.spd_rgb_reflectance(value)
        elif b.get("blackbody") is not None:
            p = _split_params(b.get("blackbody"))
            if not p:
                return
            spd = spd_black_body(_atof(p[0]), scale)
        elif b.get("nblackbody") is not None:
            p = _split_params(b.get("nblackbody"))
            if not p:
                return
            sc2 = 1.0
            for i, tok in enumerate(p):
                if tok == "scale" and i + 1 < len(p):
                    sc2 = _atof(p[i + 1])
            spd = spd_black_body(_atof(p[0]), sc2, normalized=True)
        smp = b.get("samples")
        if spd is None and smp is None:
            return
        if spd is None:
            p = _split_params(smp)
            if len(p) % 2:
                return
            spd = spd_from_samples([(_atof(p[i]), _atof(p[i + 1])) for i in range(0, len(p), 2)])
            nrm = b.get("normalize")
            if nrm is not None:
                xyz = integrate_to_xyz(spd["entries"]["power"][0])
                rgbv = _xyz_to_rgb(xyz)
                lum = f32(xyz[1]) if nrm.strip() == "luminance" else f32(max(f32(0.0), rgbv.max()))  # :1597-1600
                if lum > f32(1.192092896e-07):
                    spd = spd_scaled(spd, f32(1.0) / lum)
        self.add_named_spectrum(name, spd_scaled(spd, scale))

    # -- materials -----------------------------------------------------------------------------------------
    def parse_material(self, b):
        sd = self.sd
        if b.name in sd.material_names:
            mi = sd.material_names[b.name]
        else:
            mi = sd.add_material(b.name)
        m = sd.materials[mi]
        m["cls"] = S.MAT_DIFFUSE
        m["emission"]["spectrum_index"] = S.INVALID
        m["emission"]["image_index"] = S.INVALID
        m["emission_collimation"] = 0.0
        v = b.get("base")
        if v is not None and v.strip() in sd.material_names:
            m[:] = sd.materials[sd.material_names[v.strip()]]
        for key, fld in (("Kd", "scattering"), ("Ks", "reflectance"), ("Kt", "scattering")):
            v = b.get(key)
            if v is not None:
                m[fld]["spectrum_index"] = self.reflectance_spectrum(v)
        v = b.get("two_sided")
        if v is not None:
            lead = _leading_int(v)  # an integer, else the whole value against "true" / "on" (:1714-1722)
            m["two_sided"] = (1 if lead != 0 else 0) if lead is not None else (1 if v in ("true", "on") else 0)
        v = b.get("opacity")
        if v is not None and _floats(v, 1):
            m["opacity"] = min(max(f32(_floats(v, 1)[0]), f32(0.0)), f32(1.0))
        v = b.get("Pr")
        if v is not None:
            fl = _floats(v, 2)
            if len(fl) == 2:
                m["roughness"]["value"][0][:] = [f32(fl[0]) * f32(fl[0]), f32(fl[1]) * f32(fl[1]), 0.0, 0.0]
            elif len(fl) == 1:
                m["roughness"]["value"][0][:] = [f32(fl[0]) * f32(fl[0]), f32(fl[0]) * f32(fl[0]), 0.0, 0.0]
        for key in ("metalness", "transmission"):
            v = b.get(key)
            if v is not None and _floats(v, 1):
                m[key]["value"][0][:] = f32(_floats(v, 1)[0])
        # map_Ml / map_Tm: metalness / transmission maps with an optional `channel N` (:1772-1800); `map_Pr` never reaches parse_material (the .mtl reader
        # consumes it as a standard texture key)
        for key, fld in (("map_Ml", "metalness"), ("map_Tm", "transmission")):
            v = b.get(key)
            if v is None:
                continue
            p, ch, i = _split_params(v), 0, 0
            while i < len(p):
                if p[i] == "channel" and i + 1 < len(p):
                    lead = _leading_int(p[i + 1])
                    ch = max(0, lead if lead is not None else 0)
                    i += 1
                i += 1
            f = self.find_file(p[0])
            if f:
                m[fld]["image_index"] = self.add_image_file(f, IMG_REPEAT)
                m[fld]["channel"] = ch
        for slot, fld in (("diffuse", "scattering"), ("specular", "reflectance"), ("transmittance", "scattering")):
            f = self.find_file(b.textures.get(slot))
            if f:
                m[fld]["image_index"] = self.add_image_file(f, IMG_REPEAT)
        v = b.get("material")
        if v is not None:
            p = _split_params(v)
            for i, tok in enumerate(p):
                if tok == "class" and i + 1 < len(p):
                    m["cls"] = MATERIAL_CLASSES.get(p[i + 1].lower(), S.MAT_DIFFUSE)
        v = b.get("diffuse")
        if v is not None and v.split() and v.split()[0].isdigit():
            m["diffuse_variation"] = int(v.split()[0])
        v = b.get("int_ior")
        if v is not None:
            self.load_ior(m["int_ior"], v)
        else:
            m["int_ior"]["cls"] = S.SPD_DIELECTRIC
            m["int_ior"]["eta_index"] = sd.add_spectrum(scenes.spd_constant(1.5))
            m["int_ior"]["k_index"] = sd.add_spectrum(scenes.spd_constant(0.0))
        v = b.get("ext_ior")
        if v is not None:
            self.load_ior(m["ext_ior"], v)
        else:
            m["ext_ior"]["cls"] = S.SPD_DIELECTRIC
            m["ext_ior"]["eta_index"] = sd.add_spectrum(scenes.spd_constant(1.0))
            m["ext_ior"]["k_index"] = sd.add_spectrum(scenes.spd_constant(0.0))
        for key in ("int_medium", "ext_medium"):
            v = b.get(key)
            if v is not None:
                if v.strip() not in self.medium_names:
                    self.warn(f"medium {v.strip()} was not declared, but used in material {b.name}")
                m[key] = self.medium_names.get(v.strip(), S.INVALID)
        v = b.get("normalmap")
        if v is not None:
            p, i = _split_params(v), 0
            while i < len(p):
                if p[i] == "image" and i + 1 < len(p):
                    f = self.find_file(p[i + 1])
                    if f:
                        m["normal_image_index"] = self.add_image_file(f, IMG_REPEAT | IMG_SKIP_SRGB)
                    i += 1
                if i < len(p) and p[i] == "scale" and i + 1 < len(p):
                    m["normal_scale"] = f32(_atof(p[i + 1]))
                    i += 1
                i += 1
        v = b.get("thinfilm")
        if v is not None:
            p, i = _split_params(v), 0
            while i < len(p):
                if p[i] == "image" and i + 1 < len(p):
                    f = self.find_file(p[i + 1])
                    if f:
                        m["thinfilm"]["thickness_image"] = self.add_image_file(f, IMG_REPEAT)
                    i += 1
                if i < len(p) and p[i] == "range" and i + 2 < len(p):
                    m["thinfilm"]["min_thickness"] = f32(_atof(p[i + 1]))
                    m["thinfilm"]["max_thickness"] = f32(_atof(p[i + 2]))
                    i += 2
                if i < len(p) and p[i] == "ior" and i + 1 < len(p):
                    fl = _floats(p[i + 1], 1)
                    if fl:
                        m["thinfilm"]["ior"]["cls"] = S.SPD_DIELECTRIC
                        m["thinfilm"]["ior"]["eta_index"] = sd.add_spectrum(scenes.spd_constant(fl[0]))
                        m["thinfilm"]["ior"]["k_index"] = S.INVALID
                    else:
                        self.load_ior(m["thinfilm"]["ior"], p[i + 1])
                i += 1
        v = b.get("subsurface")
        if v is not None:
            m["subsurface"]["cls"] = 1
            scale, dist = 1.0, [1.0, 0.2, 0.04]
            p, i = _split_params(v), 0
            while i < len(p):
                if p[i] == "path" and i + 1 < len(p):
                    m["subsurface"]["path"] = 1 if p[i + 1] in ("refracted", "refraction", "refract") else 0
                if p[i] == "distances" and i + 3 < len(p):
                    dist = [_atof(p[i + 1]), _atof(p[i + 2]), _atof(p[i + 3])]
                    i += 3
                if i < len(p) and p[i] == "scale" and i + 1 < len(p):
                    scale = _atof(p[i + 1])
                    i += 1
                if i < len(p) and p[i] == "class" and i + 1 < len(p):
                    if p[i + 1] == "approximate":
                        m["subsurface"]["cls"] = 2
                    i += 1
                i += 1
            m["subsurface"]["spectrum_index"] = sd.add_spectrum(spd_scaled(scenes.spd_rgb_reflectance(dist), scale))
        # emission (:2009-2078)
        spd, defined, is_emitter, pending = scenes.spd_constant(0.0), False, False, f32(1.0)
        coll = float(m["emission_collimation"][0])
        v = b.get("Ke")
        if v is not None:
            is_emitter, spd, defined = True, self.illuminant_spectrum(v), True
            f = self.find_file(b.textures.get("emissive"))
            if f:
                m["emission"]["image_index"] = self.add_image_file(f, IMG_REPEAT | IMG_BUILD_TABLE)
        v = b.get("emitter")
        if v is not None:
            is_emitter = True
            p, i = _split_params(v), 0
            while i < len(p):
                if p[i] == "image" and i + 1 < len(p) and self.find_file(p[i + 1]):
                    m["emission"]["image_index"] = self.add_image_file(self.find_file(p[i + 1]), IMG_REPEAT | IMG_BUILD_TABLE)
                    # the reference resolves the file name INTO the buffer its parameter pointers refer to (get_file -> data_buffer, :2029): whatever
                    # follows `image <file>` on the line is lost there.  Same here, so that a scene file means the same thing in both.
                    break
                elif p[i] == "twosided":
                    m["two_sided"] = 1
                elif p[i] == "collimated" and i + 1 < len(p):
                    coll = _atof(p[i + 1])
                    i += 1
                elif p[i] == "color" and i + 3 < len(p):
                    spd, defined = scenes.spd_rgb_luminance([_atof(p[i + 1]), _atof(p[i + 2]), _atof(p[i + 3])]), True
                    i += 3
                elif p[i] == "blackbody" and i + 1 < len(p):
                    spd, defined = spd_black_body(_atof(p[i + 1]), 1.0), True
                    i += 1
                elif p[i] == "nblackbody" and i + 1 < len(p):
                    spd, defined = spd_black_body(_atof(p[i + 1]), 1.0, normalized=True), True
                    i += 1
                elif p[i] == "scale" and i + 1 < len(p):
                    pending = f32(pending * f32(_atof(p[i + 1])))
                    i += 1
                i += 1
            coll = min(max(coll, 0.0), 1.0)
        if is_emitter:
            spd = spd_scaled(spd, pending)
            m["emission_collimation"] = f32(coll)
            if defined and float(scenes.luminance(spd["integrated"][0])) > 0.0:
                m["emission"]["spectrum_index"] = sd.add_spectrum(spd)
            elif (not defined) and int(m["emission"]["spectrum_index"][0]) != S.INVALID:
                pass
            else:
                m["emission"]["spectrum_index"] = S.INVALID
            if int(m["emission"]["spectrum_index"][0]) == S.INVALID:
                m["emission"]["image_index"] = S.INVALID
        elif int(m["emission"]["spectrum_index"][0]) == S.INVALID:
            m["emission"]["image_index"] = S.INVALID
            m["emission_collimation"] = 0.0

    def parse_materials(self, mtl_path):
        self.base_dir = os.path.dirname(mtl_path)
        for b in parse_mtl(mtl_path):
            if b.name == "et::camera":
                self.parse_camera(b)
            elif b.name == "et::medium":
                self.parse_medium(b)
            elif b.name == "et::dir":
                self.parse_directional(b)
            elif b.name == "et::env":
                self.parse_env(b)
            elif b.name == "et::atmosphere":
                self.parse_atmosphere(b)
            elif b.name == "et::spectrum":
                self.parse_spectrum(b)
            else:
                self.parse_material(b)

    # -- geometry --------------------------------------------------------------------------------------------
    def load_obj(self, obj_path, mtl_path):
        sd = self.sd
        o = parse_obj(obj_path)
        if not mtl_path:
            if not o.mtllib:
                raise LoaderError(f"{obj_path}: no material file")
            mtl_path = os.path.join(os.path.dirname(obj_path), o.mtllib)
        self.parse_materials(mtl_path)
        names = sd.material_names
        keep = np.array([(n is not None) and (n in names) for n in o.face_mtl], dtype=bool)
        mats = np.array([names[n] for n, k in zip(o.face_mtl, keep) if k], dtype=np.uint32)
        shapes = o.face_shape[keep]
        if keep.all():
            faces = o.faces
        else:
            # load_from_obj skips a face whose material is unknown WITHOUT advancing its index cursor (scene_representation.cxx:1005-1008, 1029): inside
            # that shape the j-th face that is kept reads the indices of the shape's j-th face.  Kept as the reference does it.
            faces = np.zeros((int(keep.sum()), 3, 3), dtype=np.int64)
            at = 0
            for sh in np.unique(o.face_shape):
                sel = o.face_shape == sh
                n_keep = int(keep[sel].sum())
                faces[at:at + n_keep] = o.faces[sel][:n_keep]
                at += n_keep
            order = np.argsort(o.face_shape[keep], kind="stable")
            assert (np.diff(order) > 0).all()  # shapes appear in file order: the per-shape blocks above are already in face order
        nf = faces.shape[0]
        v = np.zeros(nf * 3, dtype=S.VERTEX)
        flat = faces.reshape(-1, 3)
        v["pos"] = o.pos[flat[:, 0]]
        has_n, has_t = flat[:, 2] >= 0, flat[:, 1] >= 0
        if has_n.any():
            v["nrm"][has_n] = o.nrm[flat[has_n, 2]]
        if has_t.any():
            v["tex"][has_t] = o.tex[flat[has_t, 1]]
        tri = np.zeros(nf, dtype=S.TRIANGLE)
        tri["i"] = np.arange(nf * 3, dtype=np.uint32).reshape(nf, 3)
        tri["material_index"] = mats
        p = v["pos"].reshape(nf, 3, 3)
        gn = _cross((p[:, 1] - p[:, 0]).astype(f32), (p[:, 2] - p[:, 0]).astype(f32))
        gl = np.sqrt(_dot(gn, gn)).astype(f32)
        ok = gl != 0  # validate_triangle (:252-260): a degenerate triangle is dropped, its vertices stay
        with np.errstate(invalid="ignore", divide="ignore"):
            tri["geo_n"] = (gn / gl[:, None]).astype(f32)
        # medium bounds (:1036-1048): the running bounding box of the shape at the last triangle that carries the medium
        int_medium = np.array([int(sd.materials[int(mi)]["int_medium"][0]) for mi in range(len(sd.materials))], dtype=np.int64)
        tm = int_medium[mats]
        for med in np.unique(tm[tm != S.INVALID]):
            last = int(np.nonzero(tm == med)[0][-1])
            sel = (shapes == shapes[last]) & (np.arange(nf) <= last)
            pts = p[sel].reshape(-1, 3)
            rec = sd._mediums[int(med)]
            rec["bounds_min"][0] = pts.min(axis=0)
            rec["bounds_max"][0] = pts.max(axis=0)
        self.vertices, self.triangles = v, tri[ok]

    def finish_geometry(self, force_tangents):
        v, tri = self.vertices, self.triangles
        nt = tri.shape[0]
        referenced = np.zeros(v.shape[0], dtype=bool)
        referenced[tri["i"].reshape(-1)] = True
        # validate_normals (:304-335): an invalid vertex normal becomes the (area-weighted) geometric normal of its triangles
        p = v["pos"][tri["i"].reshape(-1)].reshape(nt, 3, 3)
        cr = _cross((p[:, 1] - p[:, 0]).astype(f32), (p[:, 2] - p[:, 0]).astype(f32))
        area = (f32(0.5) * np.sqrt(_dot(cr, cr)).astype(f32)).astype(f32)
        bad = ~_valid(v["nrm"])
        idx = tri["i"].reshape(-1)
        fix = bad[idx]
        if fix.any():
            # every vertex belongs to exactly one triangle (load_from_obj unrolls the faces): the first contribution is an assignment (:318-320)
            contrib = (np.repeat(tri["geo_n"], 3, axis=0) * np.repeat(area, 3)[:, None]).astype(f32)
            v["nrm"][idx[fix]] = _normalize(contrib[fix])
        # build_tangents (:337-398): without texture coordinates nothing; with them the tangent-space generator the reference calls, which the module's
        # host code restates (etxb_mesh_tangents, csrc/scene_loader_tangents.inl) — the C++ loader runs the same function
        span = v["tex"].max(axis=0) - v["tex"].min(axis=0) if v.shape[0] else np.zeros(2, f32)
        if float(span[0] * span[0] + span[1] * span[1]) > 1.192092896e-07:
            from . import api
            v = np.ascontiguousarray(v)
            tri = np.ascontiguousarray(tri)
            rc = api.load_library("fast").etxb_mesh_tangents(v.ctypes.data, v.shape[0], tri.ctypes.data, tri.shape[0])
            if rc != 0:
                raise LoaderError(f"etxb_mesh_tangents failed ({rc})")
        # validate_tangents (:400-418)
        need = (np.ones(v.shape[0], bool) if force_tangents else ~(_valid(v["tan"]) & _valid(v["btn"]))) & (referenced | force_tangents)
        if need.any():
            a, b2 = _orthonormal_basis(v["nrm"][need])
            v["tan"][need], v["btn"][need] = a, b2
        with np.errstate(invalid="ignore", divide="ignore"):
            b0 = v["btn"].copy()
            n = _normalize(v["nrm"])
            t = _normalize((v["tan"] - (_dot(v["tan"], n)[:, None] * n).astype(f32)).astype(f32))
            bt = _normalize(_cross(n, t))
            bt = (bt * np.where(_dot(b0, bt) > 0, f32(1.0), f32(-1.0))[:, None]).astype(f32)
        v["nrm"], v["tan"], v["btn"] = n, t, bt
        self.sd.vertices, self.sd.triangles = [v], [tri]


def _leading_int(text):
    """what sscanf("%d") / atoi read at the start of a string; None when there is no integer"""
    import re
    m = re.match(r"\s*([+-]?\d+)", text)
    return int(m.group(1)) if m else None


def _split_params(text):
    """split_params (scene_representation.cxx:463-478): cut at every single space, empty pieces kept."""
    return text.split(" ")


def focal_length_to_fov(focal_len):
    """camera.hxx focal_length_to_fov: 2 atan(kFilmHorizontalSize / (2 f))."""
    return float(f32(2.0) * np.arctan(f32(36.0) / (f32(2.0) * f32(focal_len))).astype(f32))


def load_scene(file_name, data_folder=None):
    """SceneRepresentation::load_from_file (:679-838): `file_name` is a .json scene description or an .obj file.  Returns a finalized SceneData."""
    file_name = os.path.abspath(file_name)
    base = os.path.dirname(file_name)
    ld = SceneLoader(data_folder)
    geometry, materials = file_name, ""
    samples, rr_start, max_len, min_len = 256, 6, 65535, 0  # Scene defaults (scene.hxx:41-44)
    spectral = force_tangents = False
    dflt = f32(5.0) + (f32(-5.0) / np.sqrt(f32(75.0), dtype=f32))  # the default camera's position + its (normalised) direction (:694)
    cam = dict(cls=0, viewport=(0, 0), origin=(5.0, 5.0, 5.0), target=(float(dflt),) * 3, up=(0.0, 1.0, 0.0), fov=26.99, focal=None, lens_radius=0.0, focal_distance=0.0,
               clip_near=None, clip_far=None)
    if file_name.lower().endswith(".json"):
        try:
            with open(file_name) as f:
                js = json.load(f)
        except (OSError, ValueError) as e:
            raise LoaderError(f"{file_name}: {e}")
        for key in sorted(js):
            val = js[key]
            # json_get_int / _float / _bool / _string (core/json.hxx:41-72) look at the value's type: a number where a bool is expected (or the other way
            # round) is ignored
            number = isinstance(val, (int, float)) and not isinstance(val, bool)
            if key == "samples" and number:
                samples = max(1, int(val))
            elif key == "random-termination-start" and number:
                rr_start = max(1, int(val))
            elif key == "max-path-length" and number:
                max_len = max(1, int(val))
            elif key == "min-path-length" and number:
                min_len = max(1, int(val))  # the reference clamps this one to 1 as well (:716)
            elif key == "geometry" and isinstance(val, str):
                geometry = os.path.join(base, val)
            elif key == "materials" and isinstance(val, str):
                materials = os.path.join(base, val)
            elif key == "spectral" and isinstance(val, bool):
                spectral = val
            elif key == "force-tangents" and isinstance(val, bool):
                force_tangents = val
            elif key == "camera" and isinstance(val, dict):
                for ck in sorted(val):
                    cv = val[ck]
                    cnum = isinstance(cv, (int, float)) and not isinstance(cv, bool)
                    if ck == "class":
                        cam["cls"] = 1 if cv == "eq" else 0
                    elif ck == "fov" and cnum:
                        cam["fov"] = float(cv)
                    elif ck == "focal-length" and cnum:
                        cam["focal"] = float(cv)
                    elif ck in ("lens-radius", "focal-distance", "clip-near", "clip-far") and cnum:
                        cam[ck.replace("-", "_")] = float(cv)
                    elif ck in ("origin", "target", "up") and isinstance(cv, list) and len(cv) >= 3:
                        cam[ck] = tuple(float(x) for x in cv[:3])
                    elif ck == "viewport" and isinstance(cv, list) and len(cv) >= 2:
                        cam["viewport"] = (int(cv[0]), int(cv[1]))
    if cam["viewport"][0] * cam["viewport"][1] == 0:
        cam["viewport"] = (1280, 720)
    if not geometry.lower().endswith(".obj"):
        raise LoaderError(f"{geometry}: only Wavefront .obj geometry is read by this loader (glTF is not)")
    ld.load_obj(geometry, materials)
    sd = ld.sd
    if not getattr(sd, "_distant_emitters", None):
        ld.add_default_atmosphere()
    # camera (:789-804)
    if ld.cameras:
        sel = next((c for c in ld.cameras if c["active"]), ld.cameras[0])
        origin = sel["origin"] if sel["origin"] is not None else (0.0, 0.0, 0.0)
        target = sel["target"] if sel["target"] is not None else (origin[0], origin[1], origin[2] - 1.0)
        vp = sel["viewport"] if sel["viewport"][0] * sel["viewport"][1] else (1280, 720)
        sd.set_camera(origin, target, sel["up"], vp[0], vp[1], sel["fov"], lens_radius=sel["lens_radius"], focal_distance=sel["focal_distance"], f32_trig=True,
                      **{k: sel[k] for k in ("clip_near", "clip_far") if sel[k] is not None})
        sd.camera["cls"] = sel["cls"]
        sd.camera["lens_image"] = sel["lens_image"]
        sd.camera["medium_index"] = sel["medium"]
    else:
        fov = cam["fov"]
        if cam["focal"] is not None:
            fov = float(f32(focal_length_to_fov(cam["focal"])) * f32(180.0) / f32(math.pi))
        sd.set_camera(cam["origin"], cam["target"], cam["up"], cam["viewport"][0], cam["viewport"][1], fov, lens_radius=cam["lens_radius"],
                      focal_distance=cam["focal_distance"], f32_trig=True, **{k: cam[k] for k in ("clip_near", "clip_far") if cam[k] is not None})
        sd.camera["cls"] = cam["cls"]
    ld.finish_geometry(force_tangents)
    finalize_loaded(ld, samples, spectral, max_len, min_len, rr_start)
    sd.name = "file:" + os.path.basename(file_name)
    sd.loader_warnings = ld.warnings
    return sd


def filter_image():
    """Film::generate_filter_image(PixelFilterBlackmanHarris) (film.cxx:63-67, 123-135): 128 x 128, centred."""
    n = 128
    y, x = np.mgrid[0:n, 0:n].astype(f32)
    px, py = (x - f32(n * 0.5)).astype(f32), (y - f32(n * 0.5)).astype(f32)
    dist = np.sqrt((px * px).astype(f32) + (py * py).astype(f32)).astype(f32)
    r = (f32(2.0 * math.pi) * np.clip((f32(0.5) + dist / f32(2.0 * n * 0.5)).astype(f32), 0.0, 1.0)).astype(f32)
    val = (f32(0.35875) - f32(0.48829) * np.cos(r) + f32(0.14128) * np.cos(f32(2.0) * r) - f32(0.01168) * np.cos(f32(3.0) * r)).astype(f32)
    img = np.ones((n, n, 4), dtype=f32)
    img[..., :3] = val[..., None]
    return img


def finalize_loaded(ld, samples, spectral, max_len, min_len, rr_start):
    """validate_materials (:262-302) + commit (:420-455): the loader's defaults instead of the generators' ones."""
    sd, d = ld.sd, ld.defaults
    for m in sd.materials:  # every missing spectrum is a NEW entry of the pool, like the reference's data.add_spectrum calls
        if int(m["reflectance"]["spectrum_index"][0]) == S.INVALID:
            m["reflectance"]["spectrum_index"] = sd.add_spectrum(scenes.spd_rgb_reflectance([1.0, 1.0, 1.0]))
        if int(m["scattering"]["spectrum_index"][0]) == S.INVALID:
            m["scattering"]["spectrum_index"] = sd.add_spectrum(scenes.spd_rgb_reflectance([1.0, 1.0, 1.0]))
        if int(m["subsurface"]["spectrum_index"][0]) == S.INVALID:
            m["subsurface"]["spectrum_index"] = sd.add_spectrum(scenes.spd_rgb_reflectance([1.0, 0.2, 0.04]))
        if int(m["emission"]["spectrum_index"][0]) == S.INVALID:
            m["emission"]["spectrum_index"] = sd.add_spectrum(scenes.spd_constant(0.0))
        r = m["roughness"]["value"][0]
        if r[0] > 0 or r[1] > 0:
            r[0], r[1] = max(f32(1e-6), r[0]), max(f32(1e-6), r[1])
        cond = int(m["cls"][0]) == S.MAT_CONDUCTOR
        if int(m["int_ior"]["eta_index"][0]) == S.INVALID:
            m["int_ior"]["eta_index"] = d["default_conductor_eta"] if cond else d["default_dielectric_eta"]
        if int(m["int_ior"]["k_index"][0]) == S.INVALID:
            m["int_ior"]["k_index"] = d["default_conductor_k"] if cond else sd.add_spectrum(scenes.spd_constant(0.0))
        if int(m["thinfilm"]["ior"]["k_index"][0]) == S.INVALID:
            m["thinfilm"]["ior"]["k_index"] = sd.add_spectrum(scenes.spd_constant(0.0))
        if int(m["thinfilm"]["ior"]["eta_index"][0]) == S.INVALID:
            m["thinfilm"]["ior"]["eta_index"] = sd.add_spectrum(scenes.spd_constant(1.0))
    pixel_filter = sd.add_image(filter_image(), repeat=False, build_table=True, uniform_table=True)
    sd._images[pixel_filter]["options"] = IMG_BUILD_TABLE | IMG_UNIFORM_TABLE
    sd.finalize_arrays(samples, spectral, max_len, min_len, rr_start, distant_first=True)
    sc = sd.scene
    sc["pixel_sampler_image"] = pixel_filter
    sc["pixel_sampler_radius"] = 1.5
    for key, val in d.items():
        sc[key] = val
