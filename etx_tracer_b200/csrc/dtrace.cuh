// dtrace.cuh — device ray queries over the shared BVH (bvh.h): closest hit and transmittance.
// Restates Raytracing::trace / trace_transmittance (sources/etx/rt/rt.cxx:428-466, 468-579) with the
// reference's filter semantics: Void materials are skipped, and EVERY other candidate hit draws one value
// from the path's sampler for the stochastic alpha test (scene_bsdf.hxx:128-144).
#pragma once
#include "dscene.cuh"

namespace etxb {

struct DevNodeLoad {
  const BvhNode* nodes;
  DEV BvhNode operator()(int32_t i) const {
    const float4* p = reinterpret_cast<const float4*>(nodes + i);
    float4 a = __ldg(p + 0), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    BvhNode n;
    n.lo0[0] = a.x; n.lo0[1] = a.y; n.lo0[2] = a.z;
    n.hi0[0] = a.w; n.hi0[1] = b.x; n.hi0[2] = b.y;
    n.lo1[0] = b.z; n.lo1[1] = b.w; n.lo1[2] = c.x;
    n.hi1[0] = c.y; n.hi1[1] = c.z; n.hi1[2] = c.w;
    n.child0 = __float_as_int(d.x);
    n.child1 = __float_as_int(d.y);
    n.pad0 = 0;
    n.pad1 = 0;
    return n;
  }
};
struct DevTriLoad {
  const float4* pos;
  DEV F4 operator()(uint32_t i) const {
    float4 v = __ldg(pos + i);
    return F4{v.x, v.y, v.z, v.w};
  }
};

// alpha_test_pass (scene_bsdf.hxx:128-144)
DEV bool alpha_test_rejects(const DeviceScene& sc, const etxb_material& mat, uint32_t triangle_index, float u, float v, Smp& smp) {
  float alpha_diffuse = 1.0f;
  if (mat.scattering.image_index != kInvalidIndex) {
    const DImage& img = sc.images[mat.scattering.image_index];
    if (img.options & kImageHasAlpha) {
      V2 uv = lerp_uv(sc, load_triangle(sc, triangle_index), barycentrics_uv(u, v));
      alpha_diffuse = image_evaluate_alpha(img, uv);
    }
  }
  float alpha_test_value = alpha_diffuse * mat.opacity;
  return alpha_test_value <= smp.next();
}

struct HitRec {
  float u, v, t;
  uint32_t tri;  // kInvalidIndex = miss
};

struct ClosestVisitor {
  const DeviceScene& sc;
  Smp& smp;
  HitRec best;
  DEV int operator()(uint32_t triangle_index, float u, float v, float t) {
    const etxb_material& mat = sc.materials[load_triangle_material(sc, triangle_index)];
    if (mat.cls == ETXB_MAT_VOID) return kCandIgnore;
    if (alpha_test_rejects(sc, mat, triangle_index, u, v, smp)) return kCandIgnore;
    best = {u, v, t, triangle_index};
    return kCandAccept;
  }
};

DEV HitRec trace_closest(const DeviceScene& sc, V3 o, V3 d, float tmin, float tmax, Smp& smp, TraverseStats* stats) {
  ClosestVisitor vis{sc, smp, {0.0f, 0.0f, 0.0f, kInvalidIndex}};
  DevNodeLoad nl{sc.bvh_nodes};
  DevTriLoad tl{sc.bvh_tris};
  traverse(nl, tl, o.x, o.y, o.z, d.x, d.y, d.z, tmin, tmax, vis, stats);
  return vis.best;
}

struct ShadowVisitor {
  const DeviceScene& sc;
  Smp& smp;
  bool occluded;
  DEV int operator()(uint32_t triangle_index, float u, float v, float t) {
    const etxb_material& mat = sc.materials[load_triangle_material(sc, triangle_index)];
    if (mat.cls == ETXB_MAT_VOID) return kCandIgnore;
    if (alpha_test_rejects(sc, mat, triangle_index, u, v, smp)) return kCandIgnore;
    // Boundary materials (participating media interfaces) are not on the device yet: upload rejects them,
    // so any surviving candidate occludes (rt.cxx:505-509)
    occluded = true;
    return kCandTerminate;
  }
};

// Returns transmittance 1 or 0 between p0 and p1 (no media on the device yet).
DEV float trace_transmittance(const DeviceScene& sc, V3 p0, V3 p1, Smp& smp, TraverseStats* stats) {
  V3 direction = p1 - p0;
  float t_max = dot(direction, direction);
  if (t_max <= kRayEpsilon) return 1.0f;
  t_max = sqrtf(t_max);
  direction /= t_max;
  t_max -= fmaxf(kRayEpsilon, t_max * kRayEpsilon);
  ShadowVisitor vis{sc, smp, false};
  DevNodeLoad nl{sc.bvh_nodes};
  DevTriLoad tl{sc.bvh_tris};
  traverse(nl, tl, p0.x, p0.y, p0.z, direction.x, direction.y, direction.z, kRayEpsilon, t_max, vis, stats);
  return vis.occluded ? 0.0f : 1.0f;
}

}  // namespace etxb
