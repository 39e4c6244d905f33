"""Film export (RTApplication::on_save_image_selected, sources/raytracer/app.cxx:261-295): the module's own EXR / PNG writers
and tone map.  CPU tests read the files back with zlib, with a minimal EXR parser, and — where oracle/_ref/libreference_loader.so exists —
with the reference's OWN readers (tinyexr LoadEXR, stb_image stbi_load) compiled from the reference tree; the GPU test renders, tone-maps on
the device and saves through the context."""
import ctypes as C
import os
import struct
import zlib

import numpy as np
import pytest

from conftest import ROOT
from etx_tracer_b200 import api, scenes, structs as S


def _image(w=37, h=23, seed=1):
    rng = np.random.default_rng(seed)
    img = (rng.random((h, w, 4)) * 3.0).astype(np.float32)
    img[..., 3] = 1.0
    img[0, 0, :3] = (0.0, 1e-4, 50.0)  # the linear toe of the sRGB curve, and a saturated value
    return img


def _reference_readers():
    path = os.path.join(ROOT, "oracle", "_ref", "libreference_loader.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    if not (hasattr(lib, "LoadEXR") and hasattr(lib, "stbi_load")):
        return None
    lib.LoadEXR.argtypes = [C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.POINTER(C.c_char_p)]
    lib.stbi_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    lib.stbi_load.restype = C.POINTER(C.c_ubyte)
    return lib


def _parse_exr(data):
    """Just enough of the OpenEXR layout to read back an uncompressed single-part scan-line file."""
    assert struct.unpack_from("<II", data, 0) == (20000630, 2)
    pos, attrs = 8, {}
    while data[pos] != 0:
        name_end = data.index(b"\0", pos)
        type_end = data.index(b"\0", name_end + 1)
        size, = struct.unpack_from("<I", data, type_end + 1)
        attrs[data[pos:name_end].decode()] = (data[name_end + 1:type_end].decode(), data[type_end + 5:type_end + 5 + size])
        pos = type_end + 5 + size
    pos += 1
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    assert attrs["compression"][1] == b"\0" and attrs["lineOrder"][1] == b"\0"
    names, cp = [], 0
    ch = attrs["channels"][1]
    while ch[cp] != 0:
        e = ch.index(b"\0", cp)
        names.append(ch[cp:e].decode())
        assert struct.unpack_from("<i", ch, e + 1)[0] == 2  # FLOAT
        cp = e + 1 + 16
    offsets = struct.unpack_from(f"<{h}Q", data, pos)
    out = np.zeros((h, w, 4), np.float32)
    for y, off in enumerate(offsets):
        yy, size = struct.unpack_from("<iI", data, off)
        assert yy == y and size == w * 4 * len(names)
        plane = np.frombuffer(data, np.float32, w * len(names), off + 8).reshape(len(names), w)
        for k, n in enumerate(names):
            out[y, :, "RGBA".index(n)] = plane[k]
    return out, names


def test_exr_writer_round_trips(tmp_path):
    img = _image()
    f = str(tmp_path / "film.exr")
    api.write_exr(f, img)
    back, names = _parse_exr(open(f, "rb").read())
    assert names == ["A", "B", "G", "R"] and np.array_equal(back.view(np.uint32), img.view(np.uint32))
    ref = _reference_readers()
    if ref is not None:  # the reference's own tinyexr accepts the file and returns the same floats
        out, w, h, err = C.POINTER(C.c_float)(), C.c_int(), C.c_int(), C.c_char_p()
        assert ref.LoadEXR(C.byref(out), C.byref(w), C.byref(h), f.encode(), C.byref(err)) == 0, err.value
        got = np.ctypeslib.as_array(out, (h.value, w.value, 4))
        assert (w.value, h.value) == (img.shape[1], img.shape[0]) and np.array_equal(got.view(np.uint32), img.view(np.uint32))


def test_tonemap_and_png_writer(tmp_path):
    img = _image()
    for exposure in (1.0, 0.25):
        ldr = api.tonemap(img, exposure)
        # app.cxx:271-281 restated in numpy (float32 steps)
        tm = (np.float32(1.0) - np.exp(-np.float32(exposure) * img[..., :3])).astype(np.float32)
        g = np.where(tm <= np.float32(0.0031308), np.float32(12.92) * tm, np.float32(1.055) * np.power(tm, np.float32(1.0 / 2.4), dtype=np.float32) - np.float32(0.055))
        want = (np.float32(255.0) * np.clip(g, 0.0, 1.0).astype(np.float32)).astype(np.uint8)
        assert np.abs(ldr[..., :3].astype(int) - want.astype(int)).max() <= 1 and (ldr[..., 3] == 255).all()
        assert ldr[0, 0, 0] == 0 and ldr[0, 0, 1] == int(255.0 * 12.92 * (1.0 - np.exp(-exposure * 1e-4))) and ldr[0, 0, 2] >= 254
    f = str(tmp_path / "film.png")
    ldr = api.tonemap(img, 1.0)
    api.write_png(f, ldr)
    d = open(f, "rb").read()
    assert d[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, header = 8, b"", None
    while pos < len(d):
        n, = struct.unpack(">I", d[pos:pos + 4])
        typ, body = d[pos + 4:pos + 8], d[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + body) == struct.unpack(">I", d[pos + 8 + n:pos + 12 + n])[0], typ
        if typ == b"IHDR":
            header = struct.unpack(">IIBBBBB", body)
        if typ == b"IDAT":
            idat += body
        pos += 12 + n
    assert header == (img.shape[1], img.shape[0], 8, 6, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(img.shape[0], 1 + img.shape[1] * 4)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(ldr.shape), ldr)
    ref = _reference_readers()
    if ref is not None:  # stb_image, the reader the reference's texture pool uses
        w, h, n = C.c_int(), C.c_int(), C.c_int()
        ref.stbi_set_flip_vertically_on_load(0)  # a process-wide switch the reference's texture pool leaves on (image_pool.cxx:344)
        p = ref.stbi_load(f.encode(), C.byref(w), C.byref(h), C.byref(n), 4)
        assert bool(p) and (w.value, h.value, n.value) == (img.shape[1], img.shape[0], 4)
        assert np.array_equal(np.ctypeslib.as_array(p, ldr.shape), ldr)
    # a frame larger than one stored deflate block (65535 bytes)
    big = np.random.default_rng(3).integers(0, 256, (200, 150, 4), dtype=np.uint8)
    api.write_png(str(tmp_path / "big.png"), big)
    d = open(tmp_path / "big.png", "rb").read()
    at = d.index(b"IDAT")
    n, = struct.unpack(">I", d[at - 4:at])
    rows = np.frombuffer(zlib.decompress(d[at + 4:at + 4 + n]), np.uint8).reshape(200, 1 + 150 * 4)
    assert np.array_equal(rows[:, 1:].reshape(big.shape), big)


def test_writers_reject_bad_arguments(tmp_path):
    lib = api.load_library("fast")
    img = _image(4, 4)
    assert lib.etxb_write_exr(None, api._p(img), 4, 4) < 0 and lib.etxb_write_exr(b"/nonexistent-dir/x.exr", api._p(img), 4, 4) < 0
    assert lib.etxb_write_png(str(tmp_path / "z.png").encode(), api._p(img), 0, 4) < 0


@pytest.mark.gpu
def test_device_tonemap_and_save_through_the_context(tmp_path):
    sd = scenes.cornell_box(64, 48, samples=16, spectral=True, sphere=True)
    g = api.GPUVCM(sd, flavor="fast")
    g.render(4)
    hdr = g.film(S.FILM_RESULT)
    for exposure in (1.0, 3.0):
        ldr = g.film_ldr(S.FILM_RESULT, exposure)
        host = api.tonemap(hdr, exposure)
        assert np.abs(ldr.astype(int) - host.astype(int)).max() <= 1 and (ldr[..., 3] == 255).all()  # device exp / pow against glibc's: at most one code value
        assert ldr[..., :3].mean() > 20
    g.save_image(str(tmp_path / "frame.exr"), S.FILM_RESULT)
    back, _ = _parse_exr(open(tmp_path / "frame.exr", "rb").read())
    assert np.array_equal(back.view(np.uint32), hdr.view(np.uint32))
    g.save_image(str(tmp_path / "frame.png"), S.FILM_CAMERA, tonemapped=True, exposure=2.0)
    d = open(tmp_path / "frame.png", "rb").read()
    at = d.index(b"IDAT")
    n, = struct.unpack(">I", d[at - 4:at])
    rows = np.frombuffer(zlib.decompress(d[at + 4:at + 4 + n]), np.uint8).reshape(48, 1 + 64 * 4)
    assert np.array_equal(rows[:, 1:].reshape(48, 64, 4), g.film_ldr(S.FILM_CAMERA, 2.0))
    with pytest.raises(api.EtxbError):
        g.save_image("/nonexistent-dir/frame.exr")
    g.close()
