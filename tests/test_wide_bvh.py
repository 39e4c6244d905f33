"""The 4-wide quantised tree (bvh.h WideNode, bvh_build.cpp build_wide_bvh) on the CPU: built from the BVH2 of a random triangle soup, every
leaf of the BVH2 appears exactly once, every decoded child box contains all the geometry below it (the quantisation is conservative in the float
arithmetic the device uses), and the traversal stack bound holds.  The traversal itself is tested on the GPU (test_gpu_parity.py: brute force)."""
import os
import subprocess

from conftest import ROOT

SRC = r'''
#include <cstdio>
#include <cstring>
#include <cmath>
#include <random>
#include <vector>
#include "etx_tracer_b200/csrc/bvh_build.h"
using namespace etxb;
struct Box { float lo[3], hi[3]; };
static int failures = 0;
static std::vector<int> leaf_seen;
static Box leaf_box(const Bvh& bvh, int32_t ref) {
  uint32_t r = uint32_t(~ref), first = r >> 2, count = (r & 3u) + 1u;
  Box b = {{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}};
  for (uint32_t s = first; s < first + count; ++s) {
    leaf_seen[s] += 1;
    for (int k = 0; k < 3; ++k) {
      const F4& p = bvh.tri_pos[s * 3 + k];
      const float v[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; ++a) { b.lo[a] = std::fmin(b.lo[a], v[a]); b.hi[a] = std::fmax(b.hi[a], v[a]); }
    }
  }
  return b;
}
static Box check(const Bvh& bvh, const WideBvh& w, int32_t node) {
  const WideNode& n = w.nodes[size_t(node)];
  Box all = {{1e30f, 1e30f, 1e30f}, {-1e30f, -1e30f, -1e30f}};
  if (n.count < 1 || n.count > 4) failures++;
  for (int k = 0; k < n.count; ++k) {
    Box below = (n.child[k] >= 0) ? check(bvh, w, n.child[k]) : leaf_box(bvh, n.child[k]);
    for (int a = 0; a < 3; ++a) {
      uint32_t bits = uint32_t(n.exp[a]) << 23;
      float step;
      std::memcpy(&step, &bits, 4);
      float lo = n.origin[a] + float(n.qlo[k][a]) * step, hi = n.origin[a] + float(n.qhi[k][a]) * step;  // the device's decode
      if (lo > below.lo[a] || hi < below.hi[a]) failures++;
      all.lo[a] = std::fmin(all.lo[a], below.lo[a]);
      all.hi[a] = std::fmax(all.hi[a], below.hi[a]);
    }
  }
  return all;
}
int main() {
  std::mt19937 rng(1234);
  std::uniform_real_distribution<float> u(-1.0f, 1.0f), small(-0.02f, 0.02f);
  const uint32_t tris = 20000;
  std::vector<float> pos(size_t(tris) * 9);
  std::vector<uint32_t> idx(size_t(tris) * 3);
  for (uint32_t t = 0; t < tris; ++t) {
    float c[3] = {u(rng) * 50.0f, u(rng) * 3.0f, u(rng) * 0.01f + 1000.0f};  // anisotropic extents, far from the origin
    for (int k = 0; k < 3; ++k) {
      for (int a = 0; a < 3; ++a) pos[(size_t(t) * 3 + k) * 3 + a] = c[a] + small(rng);
      idx[size_t(t) * 3 + k] = t * 3 + k;
    }
  }
  Bvh bvh;
  build_bvh(pos.data(), 12, idx.data(), 12, tris, bvh);
  WideBvh wide;
  build_wide_bvh(bvh, wide);
  leaf_seen.assign(bvh.tri_index.size(), 0);
  check(bvh, wide, 0);
  for (int s : leaf_seen) if (s != 1) failures++;
  std::printf("bvh2 nodes %zu depth %u, wide nodes %zu, max stack %u, failures %d\n", bvh.nodes.size(), bvh.max_depth, wide.nodes.size(), wide.max_stack, failures);
  if (wide.max_stack >= uint32_t(kWideStackSize)) return 2;
  if (wide.nodes.size() * 2 > bvh.nodes.size() + 2) return 3;  // four children per node: at most about half the nodes
  return failures ? 1 : 0;
}
'''


def test_wide_tree_is_a_conservative_cover_of_the_bvh2(tmp_path):
    src = tmp_path / "w.cpp"
    src.write_text(SRC)
    exe = tmp_path / "w"
    subprocess.check_call(["g++", "-std=c++17", "-O2", f"-I{ROOT}", str(src), os.path.join(ROOT, "etx_tracer_b200", "csrc", "bvh_build.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
