"""The C++ host adapter (etx_tracer_b200/host/gpu_vcm.hpp) compiles against the C ABI and keeps the Integrator contract
without a GPU: no device => enabled()==false, run()/update() are no-ops, state stays Stopped."""
import os
import subprocess
import sys

from conftest import ROOT
from etx_tracer_b200 import build as etx_build

SRC = r'''
#include <cstdio>
#include <cstring>
#include "etx_tracer_b200/host/gpu_vcm.hpp"
int main() {
  etxb::GPUVCM vcm(0);
  if (std::strcmp(vcm.name(), "VCM (B200)") != 0) return 1;
  if (vcm.set_option("vcm-merging", 0.0) != 0) return 2;
  if ((vcm.options().options & ETXB_VCM_ENABLE_MERGING) != 0) return 3;
  if (vcm.set_option("vcm-bogus", 1.0) == 0) return 4;
  if (vcm.state() != etxb::GPUVCM::State::Stopped) return 5;
  if (!vcm.enabled()) {            // CPU box: everything must fail closed
    if (vcm.can_run()) return 6;
    vcm.run();
    vcm.update();
    if (vcm.state() != etxb::GPUVCM::State::Stopped) return 7;
    float px[4];
    if (vcm.read_film(0, px, sizeof(px)) == 0) return 8;
    std::puts("no-device contract ok");
  } else {
    std::puts("device present");
  }
  return 0;
}
'''


def test_host_adapter_compiles_and_fails_closed_without_device(tmp_path):
    lib = etx_build.lib_path("fast")
    if not os.path.exists(lib):
        etx_build.build(("fast",))
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{ROOT}", str(src), lib, f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)


LOADER_SRC = r'''
#include <cstdio>
#include <cstdint>
#include "etx_tracer_b200/host/gpu_vcm.hpp"
extern "C" {
void* refloader_load(const char* data_folder, const char* scene_file);
void refloader_free(void*);
const void* refloader_scene(void*, uint64_t*);
const void* refloader_camera(void*, uint64_t*);
}
int main(int argc, char** argv) {
  // the hand-over the application performs: reference loader -> Scene / Camera PODs -> the integrator
  void* loaded = refloader_load(argv[1], argv[2]);
  if (!loaded) return 1;
  uint64_t scene_bytes = 0, camera_bytes = 0;
  const void* scene = refloader_scene(loaded, &scene_bytes);
  const void* camera = refloader_camera(loaded, &camera_bytes);
  if (scene_bytes != sizeof(etxb_scene) || camera_bytes != sizeof(etxb_camera)) return 2;  // the C mirrors ARE the reference PODs
  const etxb_scene* s = static_cast<const etxb_scene*>(scene);
  if (s->triangles.count != 138318 || s->samples != 32) return 3;
  etxb::GPUVCM vcm(0);
  int rc = vcm.commit_scene(scene, scene_bytes, camera, camera_bytes, s->samples);
  if (!vcm.enabled()) {
    if (rc == 0 || vcm.can_run()) return 4;  // no device: refused, nothing runs
    std::puts("loader -> adapter hand-over ok (no device: refused)");
  } else {
    if (rc != 0) { std::printf("upload failed: %s\n", vcm.status_str()); return 5; }
    std::puts("loader -> adapter hand-over ok (scene on the device)");
  }
  refloader_free(loaded);
  return 0;
}
'''


def test_reference_loader_hands_its_pods_to_the_host_adapter(tmp_path):
    """C++ end to end on the host side: the reference's own loader (oracle/_ref/libreference_loader.so) produces the Scene / Camera PODs, the
    adapter passes them to etxb_upload_scene; the C mirrors in include/etx_b200.h have the PODs' sizes.  Skipped without the reference tree."""
    import pytest
    from oracle import oracle_py
    if not oracle_py.ReferenceScene.available():
        pytest.skip("reference tree or oracle/_ref/libreference_loader.so not present")
    lib = etx_build.lib_path("fast")
    if not os.path.exists(lib):
        etx_build.build(("fast",))
    loader = os.path.join(ROOT, "oracle", "_ref", "libreference_loader.so")
    src = tmp_path / "l.cpp"
    src.write_text(LOADER_SRC)
    exe = tmp_path / "l"
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{ROOT}", str(src), lib, loader, f"-Wl,-rpath,{os.path.dirname(lib)}", f"-Wl,-rpath,{os.path.dirname(loader)}",
                           "-o", str(exe)])
    ref = oracle_py.REFERENCE_ROOT
    out = subprocess.run([str(exe), os.path.join(ref, "bin"), os.path.join(ref, "bin", "assets", "cornellbox", "cornellbox.json")], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout[-500:], out.stderr[-500:])
    assert "hand-over ok" in out.stdout


RENDER_SRC = r'''
// The application's side of the boundary, in C++, on a real device: PODs in host memory -> GPUVCM::commit_scene -> run() -> update() pumped
// like IntegratorThread does once per frame (non-blocking) until the integrator stops itself at scene.samples -> read_film.
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <thread>
#include <vector>
#include "etx_tracer_b200/host/gpu_vcm.hpp"
static std::vector<uint8_t> slurp(FILE* f, uint64_t n) { std::vector<uint8_t> v(n); if (n && fread(v.data(), 1, n, f) != n) v.clear(); return v; }
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  uint64_t counts[10];
  if (fread(counts, 8, 10, f) != 10) return 2;
  etxb_scene scene; etxb_camera camera;
  if (fread(&scene, sizeof(scene), 1, f) != 1 || fread(&camera, sizeof(camera), 1, f) != 1) return 3;
  std::vector<std::vector<uint8_t>> arrays;
  for (int k = 0; k < 10; ++k) arrays.push_back(slurp(f, counts[k]));
  std::vector<uint8_t> xyz = slurp(f, 441 * 3 * 4), rgbr = slurp(f, 391 * 3 * 4), sobol = slurp(f, 256 * 256), scr = slurp(f, 128 * 128 * 8), rnk = slurp(f, 128 * 128 * 8);
  fclose(f);
  etxb_array_view* views[9] = {&scene.vertices, &scene.triangles, &scene.triangle_to_emitter, &scene.materials, &scene.emitter_profiles, &scene.emitter_instances,
                               &scene.images, &scene.mediums, &scene.spectrums};
  for (int k = 0; k < 9; ++k) views[k]->a = arrays[k].empty() ? nullptr : arrays[k].data();
  scene.emitters_distribution.values.a = arrays[9].data();
  etxb::GPUVCM vcm(0);
  if (!vcm.enabled()) return 4;
  if (vcm.upload_tables((const float*)xyz.data(), (const float*)rgbr.data(), sobol.data(), scr.data(), rnk.data()) != 0) return 5;
  if (vcm.commit_scene(&scene, sizeof(scene), &camera, sizeof(camera), 6) != 0) { std::printf("%s\n", vcm.status_str()); return 6; }
  vcm.run();
  if (vcm.state() != etxb::GPUVCM::State::Running) return 7;
  uint32_t pumps = 0, idle_returns = 0;
  while (vcm.state() != etxb::GPUVCM::State::Stopped) {
    auto t0 = std::chrono::steady_clock::now();
    vcm.update();
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms < 2.0) idle_returns += 1;  // update() must not wait for an iteration
    pumps += 1;
    std::this_thread::sleep_for(std::chrono::microseconds(200));
    if (pumps > 2000000) return 8;
  }
  if (vcm.status().completed_iterations != 6) return 9;
  std::vector<float> film(size_t(camera.film_size[0]) * camera.film_size[1] * 4);
  if (vcm.read_film(ETXB_FILM_RESULT, film.data(), film.size() * 4) != 0) return 10;
  FILE* out = fopen(argv[2], "wb");
  fwrite(film.data(), 4, film.size(), out);
  fclose(out);
  std::printf("pumps %u non-blocking %u iterations %u\n", pumps, idle_returns, vcm.status().completed_iterations);
  return (idle_returns * 10 >= pumps * 9) ? 0 : 11;
}
'''


import pytest  # noqa: E402


@pytest.mark.gpu
def test_cpp_adapter_renders_on_the_device_with_a_non_blocking_update(tmp_path):
    """The C++ adapter on a real GPU: the Integrator contract end to end (run / non-blocking update pumped until Stopped / film), same film as
    the ctypes mirror renders through the same C ABI."""
    import ctypes as C
    import numpy as np
    from etx_tracer_b200 import api, scenes, structs as S
    sd = scenes.cornell_box(48, 40, samples=6, spectral=True, sphere=True)
    sc = sd.scene
    names = ["vertices", "triangles", "triangle_to_emitter", "materials", "emitter_profiles", "emitter_instances", "images", "mediums", "spectrums"]
    sizes = [S.VERTEX.itemsize, S.TRIANGLE.itemsize, 4, S.MATERIAL.itemsize, S.EMITTER_PROFILE.itemsize, S.EMITTER.itemsize, S.IMAGE.itemsize, S.MEDIUM.itemsize,
             S.SPECTRUM.itemsize]
    blobs = []
    for n, sz in zip(names, sizes):
        cnt = int(sc[n]["count"][0])
        blobs.append(bytes((C.c_char * (cnt * sz)).from_address(int(sc[n]["a"][0]))) if cnt else b"")
    assert int(sc["images"]["count"][0]) == 0 and int(sc["mediums"]["count"][0]) == 0  # flat arrays only in this scene
    ne = int(sc["emitter_instances"]["count"][0])
    blobs.append(bytes((C.c_char * ((ne + 1) * S.DIST_ENTRY.itemsize)).from_address(int(sc["emitters_distribution"]["values"]["a"][0]))))
    ct, bn = scenes.tables("color_tables"), scenes.tables("bluenoise")
    spp = 8  # next_power(min(samples = 6, 256))
    with open(tmp_path / "scene.bin", "wb") as f:
        f.write(np.array([len(b) for b in blobs], dtype=np.uint64).tobytes())
        f.write(sc.tobytes())
        f.write(sd.camera.tobytes())
        for b in blobs:
            f.write(b)
        f.write(np.ascontiguousarray(ct["xyz_441x3"], dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(ct["rgb_response_391x3"], dtype=np.float32).tobytes())
        for k in ("sobol", f"scrambling_{spp}", f"ranking_{spp}"):
            f.write(np.ascontiguousarray(bn[k], dtype=np.uint8).tobytes())
    lib = etx_build.lib_path("fast")
    src = tmp_path / "r.cpp"
    src.write_text(RENDER_SRC)
    exe = tmp_path / "r"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", f"-I{ROOT}", str(src), lib, f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", str(exe)])
    out = subprocess.run([str(exe), str(tmp_path / "scene.bin"), str(tmp_path / "film.bin")], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout[-500:], out.stderr[-500:])
    film = np.fromfile(tmp_path / "film.bin", dtype=np.float32).reshape(sd.height, sd.width, 4)
    g = api.GPUVCM(sd, flavor="fast")
    g.render(6)
    ref = g.film(S.FILM_RESULT)
    g.close()
    # same module, same iterations: the camera part is deterministic, the light image is a float-atomic sum (order differs from run to run)
    err = float(np.sqrt(((film[..., :3].astype(np.float64) - ref[..., :3]) ** 2).sum()) / np.sqrt((ref[..., :3].astype(np.float64) ** 2).sum()))
    assert err < 1e-5, f"C++ adapter and ctypes mirror must render the same film: relative L2 {err:.3e}"


PT_SRC = r'''
#include <cstdio>
#include <cstring>
#include "etx_tracer_b200/host/gpu_pt.hpp"
int main() {
  etxb::GPUPathTracing pt(0);
  if (std::strcmp(pt.name(), "Path Tracing (B200)") != 0) return 1;
  if (pt.set_option("nee", 0.0) != 0 || pt.options().nee != 0u || pt.options().mis != 1u) return 2;
  if (pt.set_option("bn", 0.0) != 0 || pt.options().blue_noise != 0u) return 3;
  if (pt.set_option("vcm-merging", 1.0) == 0) return 4;  // not one of CPUPathTracing's option ids
  if (pt.state() != etxb::GPUPathTracing::State::Stopped) return 5;
  if (!pt.enabled()) {
    if (pt.can_run()) return 6;
    pt.run();
    pt.update();
    if (pt.state() != etxb::GPUPathTracing::State::Stopped) return 7;
    float px[4];
    if (pt.read_film(ETXB_FILM_NORMALS, px, sizeof(px)) == 0) return 8;
    if (pt.have_updated_light_image()) return 9;
    std::puts("no-device contract ok");
  } else {
    std::puts("device present");
  }
  return 0;
}
'''


def test_path_tracing_adapter_compiles_and_fails_closed_without_device(tmp_path):
    lib = etx_build.lib_path("fast")
    if not os.path.exists(lib):
        etx_build.build(("fast",))
    src = tmp_path / "p.cpp"
    src.write_text(PT_SRC)
    exe = tmp_path / "p"
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{ROOT}", str(src), lib, f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)


@pytest.mark.gpu
def test_cpp_path_tracing_adapter_renders_on_the_device(tmp_path):
    """The same application-side program as above with the path-tracing adapter (host/gpu_pt.hpp): run() schedules iteration 0 itself, update() is
    pumped without blocking until the integrator stops at scene.samples; the camera layer equals what the ctypes mirror renders, bit for bit (the
    scene's shadow rays are traced inline, nothing in the pass is order dependent)."""
    import ctypes as C
    import numpy as np
    from etx_tracer_b200 import api, scenes, structs as S
    sd = scenes.cornell_box(48, 40, samples=6, spectral=True, sphere=True)
    sc = sd.scene
    names = ["vertices", "triangles", "triangle_to_emitter", "materials", "emitter_profiles", "emitter_instances", "images", "mediums", "spectrums"]
    sizes = [S.VERTEX.itemsize, S.TRIANGLE.itemsize, 4, S.MATERIAL.itemsize, S.EMITTER_PROFILE.itemsize, S.EMITTER.itemsize, S.IMAGE.itemsize, S.MEDIUM.itemsize,
             S.SPECTRUM.itemsize]
    blobs = []
    for n, sz in zip(names, sizes):
        cnt = int(sc[n]["count"][0])
        blobs.append(bytes((C.c_char * (cnt * sz)).from_address(int(sc[n]["a"][0]))) if cnt else b"")
    ne = int(sc["emitter_instances"]["count"][0])
    blobs.append(bytes((C.c_char * ((ne + 1) * S.DIST_ENTRY.itemsize)).from_address(int(sc["emitters_distribution"]["values"]["a"][0]))))
    ct, bn = scenes.tables("color_tables"), scenes.tables("bluenoise")
    with open(tmp_path / "scene.bin", "wb") as f:
        f.write(np.array([len(b) for b in blobs], dtype=np.uint64).tobytes())
        f.write(sc.tobytes())
        f.write(sd.camera.tobytes())
        for b in blobs:
            f.write(b)
        f.write(np.ascontiguousarray(ct["xyz_441x3"], dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(ct["rgb_response_391x3"], dtype=np.float32).tobytes())
        for k in ("sobol", "scrambling_8", "ranking_8"):
            f.write(np.ascontiguousarray(bn[k], dtype=np.uint8).tobytes())
    code = (RENDER_SRC.replace("gpu_vcm.hpp", "gpu_pt.hpp").replace("etxb::GPUVCM", "etxb::GPUPathTracing").replace("sizeof(camera), 6)", "sizeof(camera))")
            .replace("ETXB_FILM_RESULT", "ETXB_FILM_CAMERA"))
    assert "GPUPathTracing pt" not in code and "etxb::GPUPathTracing vcm(0)" in code
    lib = etx_build.lib_path("fast")
    src = tmp_path / "rp.cpp"
    src.write_text(code)
    exe = tmp_path / "rp"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", f"-I{ROOT}", str(src), lib, f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", str(exe)])
    out = subprocess.run([str(exe), str(tmp_path / "scene.bin"), str(tmp_path / "film.bin")], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout[-500:], out.stderr[-500:])
    film = np.fromfile(tmp_path / "film.bin", dtype=np.float32).reshape(sd.height, sd.width, 4)
    g = api.GPUPathTracing(sd, flavor="fast")
    g.render(6)
    ref = g.film(S.FILM_CAMERA)
    g.close()
    assert np.array_equal(film[..., :3].view(np.uint32), ref[..., :3].view(np.uint32))


def _build_native_renderer(tmp_path):
    lib = etx_build.lib_path("fast")
    if not os.path.exists(lib):
        etx_build.build(("fast",))
    exe = tmp_path / "etx_render"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-Wall", "-Werror", f"-I{ROOT}", os.path.join(ROOT, "etx_tracer_b200", "host", "render_main.cpp"), lib,
                           f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", str(exe)])
    return str(exe)


def _tiny_scene(tmp_path):
    (tmp_path / "s.obj").write_text("mtllib s.mtl\nv -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nv -0.3 1.9 -0.3\nv 0.3 1.9 -0.3\nv 0.3 1.9 0.3\nv -0.3 1.9 0.3\n"
                                    "usemtl floor\nf 1 4 3 2\nusemtl lamp\nf 5 6 7 8\n")
    (tmp_path / "s.mtl").write_text("newmtl floor\nKd 0.6 0.5 0.4\n\nnewmtl lamp\nKd 0 0 0\nKe 9 9 9\n\nnewmtl et::camera\nviewport 40 30\norigin 0 1 3.5\ntarget 0 0.8 0\nfov 45\n")
    (tmp_path / "s.json").write_text('{"geometry": "s.obj", "materials": "s.mtl", "samples": 4, "max-path-length": 6}')
    return str(tmp_path / "s.json")


def test_native_renderer_loads_the_scene_and_fails_closed_without_device(tmp_path):
    """host/render_main.cpp (scene file -> C++ loader -> integrator adapter -> film export) compiles warning-free against the C ABI; on a box without a
    GPU it gets as far as the loaded scene and stops with exit code 4, a scene it cannot read is exit code 3."""
    exe = _build_native_renderer(tmp_path)
    out = subprocess.run([exe, _tiny_scene(tmp_path), "-o", str(tmp_path / "o.exr"), "--option", "vcm-merging=0"], capture_output=True, text=True)
    assert out.returncode in (0, 4), (out.returncode, out.stdout[-300:], out.stderr[-300:])
    if out.returncode == 4:
        assert "no CUDA device" in out.stderr and not os.path.exists(tmp_path / "o.exr")
    out = subprocess.run([exe, str(tmp_path / "absent.json")], capture_output=True, text=True)
    assert out.returncode == 3 and "absent.json" in out.stderr
    assert subprocess.run([exe], capture_output=True).returncode == 2


def test_scene_file_tables_are_the_shipped_tables(tmp_path):
    """tables.bin (what the C++ side reads) holds the same colour, blue-noise and spectrum tables as the npz files the Python side reads."""
    import numpy as np
    from etx_tracer_b200 import api, scenes
    sf = api.SceneFile(_tiny_scene(tmp_path))
    for stem in ("color_tables", "bluenoise", "spectra"):
        z = scenes.tables(stem)
        for key in z:
            if key.startswith("blackbody_") or key.startswith("nblackbody_"):
                continue
            t = sf.table(f"{stem}/{key}", z[key].dtype)
            assert t is not None and np.array_equal(t, np.asarray(z[key]).reshape(-1)), f"{stem}/{key}: run tools/make_tables_bin.py"
    assert sf.table("bluenoise/none") is None
    assert (sf.width, sf.height, sf.triangle_count) == (40, 30, 4) and int(sf.scene["samples"][0]) == 4
    sf.lib.etxb_scene_file_set_samples(sf.h, 9)
    assert int(sf.scene["samples"][0]) == 9
    sf.close()


def test_native_renderer_reads_the_applications_options_file(tmp_path):
    """`etx_render --options options.json`: the scene and the integrator come from the application's own options file (util/options.cxx format; the ids
    RTApplication::init reads, raytracer/app.cxx:88-105); "Bidirectional (CPU)" maps to VCM with merging off (vcm_shared.hxx:33)."""
    import ctypes as C
    import json
    from etx_tracer_b200 import api
    exe = _build_native_renderer(tmp_path)
    scene = _tiny_scene(tmp_path)
    values = [{"class": 5, "description": "Integrator", "id": "integrator", "meta": 0, "value": "Bidirectional (CPU)"},
              {"class": 5, "description": "Scene", "id": "scene", "meta": 0, "value": "./" + os.path.basename(scene)},
              {"class": 1, "description": "Samples", "id": "spp", "meta": 0, "value": 4}]
    (tmp_path / "options.json").write_text(json.dumps({"values": values}, indent=2))
    lib = api.load_library("fast")
    buf = C.create_string_buffer(256)
    assert lib.etxb_options_file_string(str(tmp_path / "options.json").encode(), b"integrator", buf, len(buf)) == len("Bidirectional (CPU)") and buf.value == b"Bidirectional (CPU)"
    assert lib.etxb_options_file_string(str(tmp_path / "options.json").encode(), b"recent-0", buf, len(buf)) == 0
    assert lib.etxb_options_file_string(str(tmp_path / "options.json").encode(), b"spp", buf, len(buf)) == 0  # not a string option
    assert lib.etxb_options_file_string(str(tmp_path / "absent.json").encode(), b"scene", buf, len(buf)) < 0
    out = subprocess.run([exe, "--options", str(tmp_path / "options.json"), "-o", str(tmp_path / "o.exr")], capture_output=True, text=True)
    assert out.returncode in (0, 4), (out.returncode, out.stdout[-300:], out.stderr[-300:])  # 4: the scene was found and loaded, no device here
    if out.returncode == 4:
        assert "VCM (B200)" in out.stderr
    values[1]["value"] = "./missing.json"
    (tmp_path / "options.json").write_text(json.dumps({"values": values}))
    assert subprocess.run([exe, "--options", str(tmp_path / "options.json")], capture_output=True).returncode == 3
