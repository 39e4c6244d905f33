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
