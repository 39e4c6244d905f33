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
