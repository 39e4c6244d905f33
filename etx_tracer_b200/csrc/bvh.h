// bvh.h — BVH2 layout + the one traversal routine used by BOTH the CUDA kernels and the CPU oracle.
//
// Replaces the reference's Embree scene (sources/etx/rt/rt.cxx:66-88 build, :250-279 rtcIntersect1).
// Embree is not available offline, and the reference draws one RNG value per *candidate* triangle hit
// inside the Embree filter callback (rt.cxx:436-457 -> scene_bsdf.hxx:128-144), so the sampler stream
// depends on traversal order.  Sharing this header between device and oracle pins that order: same nodes,
// same near-first rule, same triangle test with the same float operation order.
//
// Layout (HBM):
//   BvhNode (64 B, one 128-bit x4 load): both children's boxes + two child references.
//   child >= 0 : inner node index;  child < 0 : leaf, ~child = (first_slot << 2) | (count-1), count<=4.
//   tri_pos   : float4[3*slots] — leaf-ordered vertex positions (48 B/triangle), .w of vertex 0 holds
//               the original triangle index (bit-cast uint32).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BVH_FN __host__ __device__ __forceinline__
#else
#define BVH_FN inline
#endif

namespace etxb {

struct alignas(16) BvhNode {
  float lo0[3], hi0[3];
  float lo1[3], hi1[3];
  int32_t child0, child1;
  uint32_t pad0, pad1;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

struct alignas(16) F4 {
  float x, y, z, w;
};

constexpr int kBvhStackSize = 64;
constexpr int kBvhMaxLeafTris = 4;

// Wide, quantised form of the same tree for the product build's dedicated traversal kernels (dwide.cuh): up to four children per node — the
// BVH2 node's grandchildren, largest box opened first — whose boxes are stored as 8-bit offsets from the node's own corner on a per-axis
// power-of-two grid (conservative: the decoded box contains the exact one).  64 bytes cover four children instead of two: half the dependent
// fetches per ray and half the bytes per box.  Leaves are the BVH2's leaves (same references into tri_pos).  Built by build_wide_bvh().
struct alignas(16) WideNode {
  float origin[3];     // min corner of the union of the children
  uint8_t exp[3];      // per axis: grid step = 2^(exp - 127)
  uint8_t count;       // children in use (1..4)
  uint8_t qlo[4][3];   // child box, min corner, in grid steps from origin (rounded down)
  uint8_t qhi[4][3];   // max corner (rounded up)
  int32_t child[4];    // >= 0: wide node index; < 0: leaf, ~child = (first_slot << 2) | (count - 1)
  uint32_t pad[2];
};
static_assert(sizeof(WideNode) == 64, "WideNode must be 64 bytes");
constexpr int kWideStackSize = 96;

BVH_FN uint32_t f2u(float f) {
  union { float f; uint32_t u; } c;
  c.f = f;
  return c.u;
}
BVH_FN float u2f(uint32_t u) {
  union { float f; uint32_t u; } c;
  c.u = u;
  return c.f;
}
BVH_FN float bmin(float a, float b) { return a < b ? a : b; }
BVH_FN float bmax(float a, float b) { return a > b ? a : b; }

// Slab test against [tmin, tmax]; returns entry distance in t_entry.
BVH_FN bool slab(const float* lo, const float* hi, float ox, float oy, float oz, float ix, float iy, float iz, float tmin, float tmax, float& t_entry) {
  float tx0 = (lo[0] - ox) * ix, tx1 = (hi[0] - ox) * ix;
  float ty0 = (lo[1] - oy) * iy, ty1 = (hi[1] - oy) * iy;
  float tz0 = (lo[2] - oz) * iz, tz1 = (hi[2] - oz) * iz;
  float tn = bmax(bmax(bmin(tx0, tx1), bmin(ty0, ty1)), bmax(bmin(tz0, tz1), tmin));
  float tf = bmin(bmin(bmax(tx0, tx1), bmax(ty0, ty1)), bmin(bmax(tz0, tz1), tmax));
  tf = tf * 1.0000004f;  // conservative far plane (2 ulp)
  t_entry = tn;
  return tn <= tf;
}

// Moeller-Trumbore with a fixed operation order.  Returns true if tmin < t < tmax.
BVH_FN bool tri_test(const F4& a, const F4& b, const F4& c, float ox, float oy, float oz, float dx, float dy, float dz, float tmin, float tmax, float& t, float& u, float& v) {
  float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z;
  float e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
  float px = dy * e2z - dz * e2y;
  float py = dz * e2x - dx * e2z;
  float pz = dx * e2y - dy * e2x;
  float det = e1x * px + e1y * py + e1z * pz;
  if (det == 0.0f) return false;
  float inv = 1.0f / det;
  float tx = ox - a.x, ty = oy - a.y, tz = oz - a.z;
  u = (tx * px + ty * py + tz * pz) * inv;
  if (!(u >= 0.0f && u <= 1.0f)) return false;
  float qx = ty * e1z - tz * e1y;
  float qy = tz * e1x - tx * e1z;
  float qz = tx * e1y - ty * e1x;
  v = (dx * qx + dy * qy + dz * qz) * inv;
  if (!(v >= 0.0f && (u + v) <= 1.0f)) return false;
  t = (e2x * qx + e2y * qy + e2z * qz) * inv;
  return (t > tmin) && (t < tmax);
}

// Candidate actions returned by the visitor.
enum : int {
  kCandIgnore = 0,     // reject, keep tmax            (Embree: *valid = 0)
  kCandAccept = 1,     // accept, shrink tmax to t     (Embree: valid stays -1 for rtcIntersect1)
  kCandTerminate = 2,  // accept and stop traversal    (occlusion found)
};

struct TraverseStats {
  uint32_t nodes = 0;
  uint32_t tris = 0;
};

// Generic near-first stack traversal.  `visit(tri_index, u, v, t)` is called for every triangle hit
// with tmin < t < current tmax, in traversal order — this is where the reference's per-candidate RNG
// draw happens.  NodeLoad/TriLoad abstract how 64-B nodes / 16-B vertices are fetched (plain pointers on
// the host, 128-bit read-only loads on the device).
template <class NodeLoad, class TriLoad, class Visitor>
BVH_FN void traverse(const NodeLoad& load_node, const TriLoad& load_tri, float ox, float oy, float oz, float dx, float dy, float dz, float tmin, float tmax, Visitor& visit,
  TraverseStats* stats) {
  float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
  int32_t stack[kBvhStackSize];
  int sp = 0;
  int32_t cur = 0;  // root is always an inner node
  for (;;) {
    if (cur >= 0) {
      BvhNode n = load_node(cur);
      if (stats) stats->nodes++;
      float t0, t1;
      bool h0 = slab(n.lo0, n.hi0, ox, oy, oz, ix, iy, iz, tmin, tmax, t0);
      bool h1 = slab(n.lo1, n.hi1, ox, oy, oz, ix, iy, iz, tmin, tmax, t1);
      if (h0 && h1) {
        bool first0 = t0 <= t1;
        int32_t nearc = first0 ? n.child0 : n.child1;
        int32_t farc = first0 ? n.child1 : n.child0;
        if (sp < kBvhStackSize) stack[sp++] = farc;
        cur = nearc;
        continue;
      } else if (h0) {
        cur = n.child0;
        continue;
      } else if (h1) {
        cur = n.child1;
        continue;
      }
    } else {
      uint32_t ref = uint32_t(~cur);
      uint32_t first = ref >> 2;
      uint32_t count = (ref & 3u) + 1u;
      bool stop = false;
      for (uint32_t k = 0; k < count; ++k) {
        uint32_t slot = first + k;
        F4 a = load_tri(slot * 3u + 0u);
        F4 b = load_tri(slot * 3u + 1u);
        F4 c = load_tri(slot * 3u + 2u);
        if (stats) stats->tris++;
        float t, u, v;
        if (tri_test(a, b, c, ox, oy, oz, dx, dy, dz, tmin, tmax, t, u, v)) {
          int action = visit(f2u(a.w), u, v, t);
          if (action == kCandAccept) {
            tmax = t;
          } else if (action == kCandTerminate) {
            stop = true;
            break;
          }
        }
      }
      if (stop) return;
    }
    if (sp == 0) return;
    cur = stack[--sp];
  }
}

}  // namespace etxb
