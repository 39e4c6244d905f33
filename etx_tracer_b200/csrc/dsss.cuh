// dsss.cuh — subsurface scattering exits used by the VCM steps: volumetric random walk inside the mesh (gather_rw) and
// Christensen-Burley disk probes (gather_cb), plus the two ray queries they need (closest hit restricted to one material,
// collect up to N hits).  Restates sources/etx/rt/shared/path_tracing_shared.hxx:43-232, render/shared/scene_bssrdf_subsurface.hxx
// and Raytracing::trace_material / continuous_trace (sources/etx/rt/rt.cxx:327-426).
#pragma once
#include "dtrace.cuh"

namespace etxb {

constexpr uint32_t kSSDirections = 3u, kSSPerDirection = 8u, kSSTotal = kSSDirections * kSSPerDirection;

// subsurface::Gather, compacted: hit records instead of full Intersections (rebuilt on demand with make_intersection)
template <bool SP>
struct SSGather {
  HitRec hits[kSSTotal];
  V3 w_i[kSSTotal];
  Spec<SP> weights[kSSTotal];
  uint32_t count, selected;
  float selected_sample_weight, total_weight;
};

// One visitor serves both queries (rt.cxx:336-360 and :384-419): hits on other materials, Void materials and alpha-test rejects are
// skipped; accepted hits are recorded until the buffer is full, and only a full buffer lets the hit shorten the ray.  With a
// one-entry buffer that is the closest hit (every accepted candidate overwrites the entry and shortens the ray).
struct MaterialHitVisitor {
  const DeviceScene& sc;
  Smp& smp;
  uint32_t material_id;
  HitRec* buffer;
  uint32_t count, max_count;
  bool closest;
  DEV int operator()(uint32_t triangle_index, float u, float v, float t) {
    uint32_t mi = load_triangle_material(sc, triangle_index);
    if ((material_id != kInvalidIndex) && (mi != material_id)) return kCandIgnore;
    const etxb_material& mat = sc.materials[mi];
    if (mat.cls == ETXB_MAT_VOID) return kCandIgnore;
    if (alpha_test_rejects(sc, mat, triangle_index, u, v, smp)) return kCandIgnore;
    if (closest) {
      buffer[0] = {u, v, t, triangle_index};
      count = 1u;
      return kCandAccept;
    }
    if (count < max_count) {
      buffer[count] = {u, v, t, triangle_index};
      count += 1u;
    }
    return (count < max_count) ? kCandIgnore : kCandAccept;
  }
};
DEVN uint32_t trace_material_hits(const DeviceScene& sc, V3 o, V3 d, float tmin, float tmax, uint32_t material_id, HitRec* buffer, uint32_t max_count, bool closest, Smp& smp,
                                  TraverseStats* stats) {
  MaterialHitVisitor vis{sc, smp, material_id, buffer, 0u, max_count, closest};
  DevNodeLoad nl{sc.bvh_nodes};
  DevTriLoad tl{sc.bvh_tris};
  traverse(nl, tl, o.x, o.y, o.z, d.x, d.y, d.z, tmin, tmax, vis, stats);
  return vis.count;
}
// Raytracing::trace_material (rt.cxx:327-371)
DEV HitRec trace_material(const DeviceScene& sc, V3 o, V3 d, float tmin, float tmax, uint32_t material_id, Smp& smp, TraverseStats* stats) {
  HitRec best = {0.0f, 0.0f, 0.0f, kInvalidIndex};
  trace_material_hits(sc, o, d, tmin, tmax, material_id, &best, 1u, true, smp, stats);
  return best;
}
// Raytracing::continuous_trace (rt.cxx:373-426)
DEV uint32_t continuous_trace(const DeviceScene& sc, V3 o, V3 d, float tmin, float tmax, uint32_t material_id, HitRec* buffer, uint32_t max_count, Smp& smp, TraverseStats* stats) {
  return trace_material_hits(sc, o, d, tmin, tmax, material_id, buffer, max_count, false, smp, stats);
}

// scene_bssrdf_subsurface.hxx:17-47
DEV void remap_channel(float color, float scattering_distances, float& albedo, float& extinction, float& scattering) {
  constexpr float a = 1.826052378200f;
  constexpr float b = 4.985111943850f + 0.12735595943800f;
  constexpr float c = 1.096861024240f;
  constexpr float d = 0.496310210422f;
  constexpr float e = 4.231902997010f + 0.00310603949088f;
  constexpr float f = 2.406029994080f;
  constexpr float kMinScattering = 1.0f / 1024.0f;
  color = fmaxf(0.0f, color);
  float blend = m_pow(color, 0.25f);
  albedo = (1.0f - blend) * a * m_pow(m_atan(b * color), c) + blend * d * m_pow(m_atan(e * color), f);
  albedo = albedo < 0.0f ? 0.0f : (albedo > (1.0f - kEpsilon) ? (1.0f - kEpsilon) : albedo);
  extinction = 1.0f / fmaxf(scattering_distances, kMinScattering);
  scattering = extinction * albedo;
}
template <bool SP>
DEV void remap(Spec<SP> color, Spec<SP> distances, Spec<SP>& albedo, Spec<SP>& extinction, Spec<SP>& scattering) {
  if constexpr (SP) {
    remap_channel(color.v, distances.v, albedo.v, extinction.v, scattering.v);
  } else {
    remap_channel(color.x, distances.x, albedo.x, extinction.x, scattering.x);
    remap_channel(color.y, distances.y, albedo.y, extinction.y, scattering.y);
    remap_channel(color.z, distances.z, albedo.z, extinction.z, scattering.z);
  }
}
DEV float safe_mul1(float a, float b) { return (a == 0.0f) || (b == 0.0f) ? 0.0f : a * b; }
DEV Spec<true> safe_mul(Spec<true> a, Spec<true> b) { return {safe_mul1(a.v, b.v)}; }
DEV Spec<false> safe_mul(Spec<false> a, Spec<false> b) { return {safe_mul1(a.x, b.x), safe_mul1(a.y, b.y), safe_mul1(a.z, b.z)}; }
DEV etxb_spectral_image subsurface_image(const etxb_material& mat) { return {mat.subsurface.spectrum_index, mat.subsurface.image_index}; }

// path_tracing_shared.hxx:61-147
template <bool SP>
DEVN bool gather_rw(const DeviceScene& sc, float wavelength, const Isect& in, Smp& smp, SSGather<SP>& result, TraverseStats* stats, uint32_t& rays) {
  constexpr uint32_t kMaxIterations = 1024u;
  const etxb_material& mat = sc.materials[in.material_index];
  float anisotropy = 0.0f;
  Spec<SP> extinction = Spec<SP>::make(0.0f), scattering = Spec<SP>::make(0.0f), albedo = Spec<SP>::make(0.0f);
  if (mat.int_medium == kInvalidIndex) {
    Spec<SP> color = apply_image<SP>(sc, mat.scattering, in.tex, wavelength);
    Spec<SP> distances = apply_image<SP>(sc, subsurface_image(mat), in.tex, wavelength);
    remap<SP>(color, distances, albedo, extinction, scattering);
  } else {
    const DMedium& medium = sc.mediums[mat.int_medium];
    anisotropy = medium.phase_function_g;
    scattering = medium_scattering<SP>(sc, medium, wavelength);
    Spec<SP> absorption = medium_absorption<SP>(sc, medium, wavelength);
    extinction = scattering + absorption;
    albedo = calculate_albedo<SP>(scattering, extinction);
  }
  V3 rd = (mat.subsurface.path == 0u) ? sample_cosine_around(smp.next_2d(), -in.nrm, 1.0f) : in.w_i;
  V3 ro = shading_pos(sc, load_triangle(sc, in.triangle_index), in.barycentric, rd);
  float max_t = kMaxFloat;
  Spec<SP> throughput = Spec<SP>::make(1.0f);
#pragma unroll 1
  for (uint32_t i = 0; i < kMaxIterations; ++i) {
    Spec<SP> pdf = Spec<SP>::make(0.0f);
    uint32_t channel = sample_spectrum_component<SP>(albedo, throughput, smp.next(), pdf);
    float scattering_distance = extinction.component(channel);
    max_t = scattering_distance > 0.0f ? (-m_log(1.0f - smp.next()) / scattering_distance) : kMaxFloat;
    if ((i == 0) && (max_t <= kRayEpsilon)) return false;
    rays += 1;
    HitRec h = trace_material(sc, ro, rd, kRayEpsilon, max_t, in.material_index, smp, stats);
    bool found = h.tri != kInvalidIndex;
    if (found) max_t = h.t;
    Spec<SP> tr = spec_exp(-max_t * extinction);
    pdf *= found ? tr : safe_mul(tr, extinction);
    if (pdf.is_zero()) return false;
    Spec<SP> weight = found ? tr : safe_mul(tr, scattering);
    throughput *= weight / pdf.sum();
    if (throughput.maximum() <= kEpsilon) return false;
    if (found) {
      Isect local_i = make_intersection(sc, rd, h.tri, h.u, h.v, h.t);
      bool w_i_in = dot(local_i.w_i, local_i.nrm) > 0.0f;
      result.hits[0] = h;
      result.w_i[0] = rd * (w_i_in ? -1.0f : +1.0f);
      result.weights[0] = throughput;
      result.count = 1u;
      result.selected = 0;
      result.selected_sample_weight = 1.0f;
      result.total_weight = 1.0f;
      return true;
    }
    V3 prev_dir = rd;
    ro = ro + rd * max_t;
    rd = sample_phase_function(prev_dir, anisotropy, smp.next_2d());
  }
  return false;
}

// scene_bssrdf_subsurface.hxx:49-146
DEV float sample_s_r(float rnd) {
  if (rnd < 0.25f) {
    rnd = fminf(4.0f * rnd, 1.0f - kEpsilon);
    return m_log(1.0f / (1.0f - rnd));
  }
  rnd = fminf((rnd - 0.25f) / 0.75f, 1.0f - kEpsilon);
  return 3.0f * m_log(1.0f / (1.0f - rnd));
}
template <bool SP>
DEV Spec<SP> ss_evaluate(const DeviceScene& sc, float wavelength, V2 tex, const etxb_material& mat, float radius) {
  Spec<SP> sd = apply_image<SP>(sc, subsurface_image(mat), tex, wavelength);
  radius = fmaxf(radius, kEpsilon);
  Spec<SP> term_0 = spec_exp(-radius / (3.0f * sd));
  Spec<SP> term_1 = term_0 * term_0 * term_0;
  Spec<SP> div = sd * (4.0f * radius * kDoublePi);
  if constexpr (SP) {
    div.v = fmaxf(div.v, kEpsilon);
  } else {
    div = {fmaxf(div.x, kEpsilon), fmaxf(div.y, kEpsilon), fmaxf(div.z, kEpsilon)};
  }
  return (term_0 + term_1) / div;
}
struct SSSample {
  V3 ray_o, ray_d;
  float ray_min_t, ray_max_t;
  V3 u, v, w, basis_prob;
  float sampled_radius;
};
DEV SSSample ss_sample_empty() {
  SSSample s = {};
  s.ray_min_t = kRayEpsilon;  // Ray{} defaults (math.hxx:660-663)
  s.ray_max_t = kMaxFloat;
  return s;
}
template <bool SP>
DEV SSSample ss_sample(const DeviceScene& sc, float wavelength, const Isect& data, const etxb_material& mat, uint32_t direction, Smp& smp) {
  Spec<SP> sampled_distance = apply_image<SP>(sc, subsurface_image(mat), data.tex, wavelength);
  uint32_t channel = uint32_t((SP ? 1.0f : 3.0f) * smp.next());
  float scattering_distance = sampled_distance.component(channel);
  if (scattering_distance == 0.0f) return ss_sample_empty();
  SSSample r = ss_sample_empty();
  if (direction == 0u) {
    r.u = data.tan; r.v = data.btn; r.w = data.nrm;
    r.basis_prob = {0.25f, 0.25f, 0.5f};
  } else if (direction == 1u) {
    r.u = data.btn; r.v = data.nrm; r.w = data.tan;
    r.basis_prob = {0.25f, 0.50f, 0.25f};
  } else {
    r.u = data.nrm; r.v = data.tan; r.w = data.btn;
    r.basis_prob = {0.5f, 0.25f, 0.25f};
  }
  constexpr float kMaxRadius = 47.827155457397595950044717258511f;
  float r_max = scattering_distance * kMaxRadius;
  r.sampled_radius = scattering_distance * sample_s_r(smp.next());
  if (r.sampled_radius >= r_max) return ss_sample_empty();
  float phi = kDoublePi * smp.next();
  float height = sqrtf(sqr(r_max) - sqr(r.sampled_radius));
  if (height <= kRayEpsilon) return ss_sample_empty();
  r.ray_o = data.pos + height * r.w + r.sampled_radius * (m_cos(phi) * r.u + m_sin(phi) * r.v);
  r.ray_d = -r.w;
  r.ray_max_t = 2.0f * height;
  return r;
}
DEV float ss_geometric_weight(V3 nrm, const SSSample& s) {
  float pdf_t = s.basis_prob.x * fabsf(dot(nrm, s.u));
  float pdf_b = s.basis_prob.y * fabsf(dot(nrm, s.v));
  float pdf_n = s.basis_prob.z * fabsf(dot(nrm, s.w));
  return sqr(pdf_n) / (sqr(pdf_t) + sqr(pdf_b) + sqr(pdf_n));
}

// path_tracing_shared.hxx:149-221
template <bool SP>
DEVN bool gather_cb(const DeviceScene& sc, float wavelength, const Isect& in, Smp& smp, SSGather<SP>& result, TraverseStats* stats, uint32_t& rays) {
  const etxb_material& mat = sc.materials[in.material_index];
  SSSample ss[kSSDirections];
#pragma unroll 1
  for (uint32_t d = 0; d < kSSDirections; ++d) ss[d] = ss_sample<SP>(sc, wavelength, in, mat, d, smp);
  HitRec probe_hits[kSSTotal];
  uint32_t found[kSSDirections] = {0u, 0u, 0u};
  uint32_t filled = 0;
#pragma unroll 1
  for (uint32_t d = 0; d < kSSDirections; ++d) {
    found[d] = continuous_trace(sc, ss[d].ray_o, ss[d].ray_d, ss[d].ray_min_t, ss[d].ray_max_t, in.material_index, probe_hits + filled, kSSPerDirection, smp, stats);
    filled += found[d];
  }
  rays += kSSDirections;
  const uint32_t n0 = found[0], n1 = found[1];
  const uint32_t intersection_count = filled;
  if (intersection_count == 0) return false;
  Spec<SP> base_weight = apply_image<SP>(sc, mat.scattering, in.tex, wavelength);
  result.count = 0;
  result.selected = 0;
  result.selected_sample_weight = 0.0f;
  result.total_weight = 0.0f;
#pragma unroll 1
  for (uint32_t i = 0; i < intersection_count; ++i) {
    const SSSample& s = ss[(i < n0) ? 0 : (i < n0 + n1 ? 1 : 2)];
    Isect out = make_intersection(sc, s.ray_d, probe_hits[i].tri, probe_hits[i].u, probe_hits[i].v, probe_hits[i].t);
    float gw = ss_geometric_weight(out.nrm, s);
    float pdf = ss_evaluate<SP>(sc, wavelength, out.tex, mat, s.sampled_radius).average();
    if (pdf <= 0.0f) continue;
    Spec<SP> eval = ss_evaluate<SP>(sc, wavelength, out.tex, mat, length(out.pos - in.pos));
    Spec<SP> weight = base_weight * eval / pdf * gw;
    if (weight.is_zero()) continue;
    result.total_weight += weight.average();
    result.hits[result.count] = probe_hits[i];
    result.w_i[result.count] = s.ray_d;
    result.weights[result.count] = weight;
    result.count += 1u;
  }
  if (result.total_weight > 0.0f) {
    float rnd = smp.next() * result.total_weight;
    float partial_sum = 0.0f;
    for (uint32_t i = 0; i < result.count; ++i) {
      float sample_weight = result.weights[i].average();
      float next_sum = partial_sum + sample_weight;
      if (rnd < next_sum) {
        result.selected = i;
        result.selected_sample_weight = result.total_weight / sample_weight;
        break;
      }
      partial_sum = next_sum;
    }
  }
  return result.count > 0;
}

// subsurface::gather (path_tracing_shared.hxx:223-232): class 2 = ChristensenBurley, everything else random walk
template <bool SP>
DEV bool ss_gather(const DeviceScene& sc, float wavelength, const Isect& in, Smp& smp, SSGather<SP>& result, TraverseStats* stats, uint32_t& rays) {
  if (sc.materials[in.material_index].subsurface.cls == 2u) return gather_cb<SP>(sc, wavelength, in, smp, result, stats, rays);
  return gather_rw<SP>(sc, wavelength, in, smp, result, stats, rays);
}

// the gathered exit point i as a full intersection carrying `material_index`
template <bool SP>
DEV Isect ss_exit_intersection(const DeviceScene& sc, const SSGather<SP>& g, uint32_t i, uint32_t material_index) {
  Isect out = make_intersection(sc, g.w_i[i], g.hits[i].tri, g.hits[i].u, g.hits[i].v, g.hits[i].t);
  out.material_index = material_index;
  return out;
}

}  // namespace etxb
