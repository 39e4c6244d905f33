// dpt.cuh — the unidirectional path tracer's per-path routines, on the same device functions as the VCM path.
//
// Reference: sources/etx/rt/shared/path_tracing_shared.hxx — PTRayPayload :16-33, make_ray_payload :238-257, try_sampling_medium :259-268,
// handle_sampled_medium :270-298, evaluate_light :300-328, handle_direct_emitter :330-357, handle_hit_ray :359-460, handle_missed_ray :462-481,
// run_path_iteration :485-510; Film::sample (render/host/film.cxx:137-145).  Every float expression keeps the reference's operation order: the
// parity build is bit-exact against the oracle's compiled reference header.
#pragma once
#include "dvcm.cuh"
#include "dsss.cuh"

namespace etxb {

// PTOptions (path_tracing_shared.hxx:8-14) + what handle_hit_ray / the driver read from the scene and the film
struct PtParams {
  uint32_t nee, direct, mis, blue_noise;
  uint32_t iteration;
  uint32_t pixel_sampler_image;  // Scene::pixel_sampler (camera.hxx:65-73)
  float pixel_sampler_radius;
  float radiance_clamp;          // Scene::radiance_clamp (path_tracing.cxx:73-78)
};

// PTRayPayload (path_tracing_shared.hxx:16-33) in registers
template <bool SP>
struct PtState {
  V3 ray_o, ray_d;
  float ray_min_t, ray_max_t;
  Spec<SP> throughput, accumulated, view_albedo;
  V3 view_normal;
  uint32_t medium, path_length;
  float wavelength, eta, sampled_bsdf_pdf;
  Smp smp;
  bool mis_weight;
};

// math.hxx:945-950
DEV float power_heuristic(float f, float g) {
  float f2 = f * f;
  float g2 = g * g;
  float denom = f2 + g2;
  return denom > 0.0f ? saturatef(f2 / denom) : 0.0f;
}

// Film::sample (film.cxx:137-145): jittered pixel position in NDC; `radius` 0 on the first iteration (PixelFilter::empty())
DEV V2 film_sample(const DeviceScene& sc, uint32_t filter_image, float radius, uint32_t px, uint32_t py, V2 rnd) {
  V2 jitter = rnd * 2.0f - 1.0f;
  if (filter_image != kInvalidIndex) {
    float pdf = 0.0f;
    F4v value;
    jitter = image_sample(sc.images[filter_image], rnd, pdf, value) * 2.0f - 1.0f;
  }
  float u = (float(px) + 0.5f + radius * jitter.x) / float(sc.camera.film_size[0]) * 2.0f - 1.0f;
  float v = (float(py) + 0.5f + radius * jitter.y) / float(sc.camera.film_size[1]) * 2.0f - 1.0f;
  return {u, v};
}

// make_ray_payload (:238-257)
template <bool SP>
DEV PtState<SP> pt_make_payload(const DeviceScene& sc, const PtParams& pt, uint32_t px, uint32_t py, uint32_t pixel_index) {
  PtState<SP> s;
  s.smp.init(pixel_index, pt.iteration);
  s.wavelength = SP ? spectral_sample_wavelength(s.smp.next()) : -1.0f;
  const bool first = pt.iteration == 0u;
  V2 uv = film_sample(sc, first ? kInvalidIndex : pt.pixel_sampler_image, first ? 0.0f : pt.pixel_sampler_radius, px, py, s.smp.next_2d());
  generate_ray(sc, sc.camera, uv, s.smp.next_2d(), s.ray_o, s.ray_d, s.ray_min_t, s.ray_max_t);
  s.throughput = Spec<SP>::make(1.0f);
  s.accumulated = Spec<SP>::make(0.0f);
  s.view_albedo = Spec<SP>::make(0.0f);
  s.view_normal = {0.0f, 0.0f, 0.0f};
  s.medium = sc.camera.medium_index;
  s.path_length = 1u;
  s.eta = 1.0f;
  s.sampled_bsdf_pdf = 0.0f;
  s.mis_weight = true;
  return s;
}

// bsdf::albedo (scene_bsdf.hxx:99-102 and the per-class bodies: conductor bsdf_conductor.hxx:133, void / mirror / boundary bsdf_various.hxx:28,259,291;
// every other class reads the scattering image)
template <bool SP>
DEV Spec<SP> bsdf_albedo(const DeviceScene& sc, const BData& data, const etxb_material& mat) {
  switch (mat.cls) {
    case ETXB_MAT_VOID:
      return Spec<SP>::make(0.0f);
    case ETXB_MAT_MIRROR:
    case ETXB_MAT_BOUNDARY:
      return Spec<SP>::make(1.0f);
    case ETXB_MAT_CONDUCTOR:
      return apply_image<SP>(sc, mat.reflectance, data.tex, data.wavelength);
    default:
      return apply_image<SP>(sc, mat.scattering, data.tex, data.wavelength);
  }
}

// evaluate_light (:300-328).  `batch` (product build, opaque scenes): the segment joins the bounce's shadow list with `scale` x the unoccluded
// contribution and the routine returns zero (k_shadow_resolve adds it to the path's accumulated sum).
template <bool SP, bool PLAIN>
DEV Spec<SP> pt_evaluate_light(const DeviceScene& sc, const Isect& isect, const etxb_material& mat, uint32_t medium, float wavelength, const EmitterSample<SP>& es, Smp& smp, bool mis,
                               TraverseStats* stats, uint32_t& shadow_rays, ShadowBatch* batch, Spec<SP> scale) {
  const Spec<SP> zero = Spec<SP>::make(0.0f);
  if (es.pdf_dir == 0.0f) return zero;
  BData data = make_bdata(isect, isect.w_i, wavelength, medium, kPathCamera);
  BEval<SP> eval = bsdf_evaluate<SP>(sc, data, es.direction, mat, smp);
  if (eval.valid() == false) return zero;
  TriRec tri = load_triangle(sc, isect.triangle_index);
  V3 pos = shading_pos(sc, tri, isect.barycentric, es.direction);
  shadow_rays += 1u;
  Spec<SP> tr = Spec<SP>::make(1.0f);
  if (batch == nullptr) tr = trace_transmittance<SP, PLAIN>(sc, wavelength, pos, es.origin, medium, smp, stats);
  bool no_weight = (mis == false) || es.is_delta;
  float weight = no_weight ? 1.0f : power_heuristic(es.pdf_dir * es.pdf_sample, eval.pdf);
  float wscale = weight / (es.pdf_dir * es.pdf_sample);
  Spec<SP> value = eval.bsdf * es.value * tr * wscale;
  if (batch != nullptr) {
    batch->push<SP>(pos, es.origin, scale * value);
    return zero;
  }
  return value;
}

// handle_direct_emitter (:330-357)
template <bool SP, bool PLAIN>
DEV void pt_handle_direct_emitter(const DeviceScene& sc, const PtParams& pt, const Isect& isect, PtState<SP>& s, TraverseStats* stats, uint32_t& shadow_rays) {
  if ((pt.direct == 0u) || (isect.emitter_index == kInvalidIndex)) return;
  const etxb_emitter& emitter = sc.emitters[isect.emitter_index];
  float pdf_emitter_area = 0.0f, pdf_emitter_dir = 0.0f, pdf_emitter_dir_out = 0.0f;
  const bool directly_visible = s.path_length == 1u;
  Spec<SP> e = emitter_get_radiance<SP>(sc, emitter, s.wavelength, s.ray_o, isect.pos, {0.0f, 0.0f, 0.0f}, isect.tex, directly_visible, pdf_emitter_area, pdf_emitter_dir,
    pdf_emitter_dir_out);
  if (pdf_emitter_dir > 0.0f) {
    shadow_rays += 1u;
    Spec<SP> tr = trace_transmittance<SP, PLAIN>(sc, s.wavelength, s.ray_o, isect.pos, s.medium, s.smp, stats);
    float pdf_emitter_discrete = emitter_discrete_pdf(sc, emitter);
    bool no_weight = (pt.mis == 0u) || directly_visible || (s.mis_weight == false);
    float weight = no_weight ? 1.0f : power_heuristic(s.sampled_bsdf_pdf, pdf_emitter_discrete * pdf_emitter_dir);
    s.accumulated += s.throughput * e * tr * weight;
  }
}

// handle_missed_ray (:462-481)
template <bool SP>
DEV void pt_handle_missed_ray(const DeviceScene& sc, PtState<SP>& s) {
  for (uint32_t ie = 0; ie < sc.env_emitter_count; ++ie) {
    const etxb_emitter& emitter = sc.emitters[sc.env_emitters[ie]];
    float pdf_emitter_area = 0.0f, pdf_emitter_dir = 0.0f, pdf_emitter_dir_out = 0.0f;
    const bool directly_visible = s.path_length == 1u;
    Spec<SP> e = emitter_get_radiance<SP>(sc, emitter, s.wavelength, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, s.ray_d, {0.0f, 0.0f}, directly_visible, pdf_emitter_area,
      pdf_emitter_dir, pdf_emitter_dir_out);
    if ((pdf_emitter_dir > 0.0f) && (e.is_zero() == false)) {
      float pdf_emitter_discrete = emitter_discrete_pdf(sc, emitter);
      float weight = ((s.mis_weight == false) || directly_visible) ? 1.0f : power_heuristic(s.sampled_bsdf_pdf, pdf_emitter_discrete * pdf_emitter_dir);
      s.accumulated += s.throughput * e * weight;
    }
  }
}

// handle_sampled_medium (:270-298); the caller then applies random_continue (:497-500)
template <bool SP, bool PLAIN>
DEV void pt_handle_sampled_medium(const DeviceScene& sc, const PtParams& pt, const MediumSample<SP>& ms, PtState<SP>& s, TraverseStats* stats, uint32_t& shadow_rays) {
  const DMedium& medium = sc.mediums[s.medium];
  if (pt.nee && (s.path_length + 1u <= sc.max_path_length) && medium.enable_explicit_connections) {
    uint32_t emitter_index = distribution_sample(sc.emitter_dist, sc.emitter_count + 1u, s.smp.next());
    EmitterSample<SP> es = sample_emitter<SP>(sc, s.wavelength, emitter_index, s.smp.next_2d(), ms.pos);
    if (es.pdf_dir > 0.0f) {
      shadow_rays += 1u;
      Spec<SP> tr = trace_transmittance<SP, PLAIN>(sc, s.wavelength, ms.pos, es.origin, s.medium, s.smp, stats);
      float phase = phase_function(s.ray_d, es.direction, medium.phase_function_g);
      float weight = es.is_delta ? 1.0f : power_heuristic(es.pdf_dir * es.pdf_sample, phase);
      s.accumulated += s.throughput * es.value * tr * (phase * weight / (es.pdf_dir * es.pdf_sample));
    }
  }
  V3 w_o = sample_phase_function(s.ray_d, medium.phase_function_g, s.smp.next_2d());
  s.sampled_bsdf_pdf = phase_function(s.ray_d, w_o, medium.phase_function_g);
  s.mis_weight = true;
  s.ray_o = ms.pos;
  s.ray_d = w_o;
  s.ray_max_t = kMaxFloat;
  s.ray_min_t = kRayEpsilon;
  s.path_length += 1u;
}

// handle_hit_ray (:359-460).  Returns whether the path continues.
template <bool SP, bool PLAIN>
DEV bool pt_handle_hit_ray(const DeviceScene& sc, const PtParams& pt, const Isect& isect, uint32_t px, uint32_t py, PtState<SP>& s, TraverseStats* stats, uint32_t& shadow_rays,
                           ShadowBatch* batch) {
  const etxb_material& mat = sc.materials[isect.material_index];
  TriRec tri = load_triangle(sc, isect.triangle_index);
  if constexpr (!PLAIN) {
    if (mat.cls == ETXB_MAT_BOUNDARY) {
      s.medium = (dot(isect.nrm, s.ray_d) < 0.0f) ? mat.int_medium : mat.ext_medium;
      s.ray_o = shading_pos(sc, tri, isect.barycentric, s.ray_d);
      s.ray_max_t = kMaxFloat;
      s.ray_min_t = kRayEpsilon;
      return true;
    }
  }
  pt_handle_direct_emitter<SP, PLAIN>(sc, pt, isect, s, stats, shadow_rays);

  BData bsdf_data = make_bdata(isect, isect.w_i, s.wavelength, s.medium, kPathCamera);
  if (s.path_length == 1u) {
    s.view_normal = isect.nrm;
    s.view_albedo = bsdf_albedo<SP>(sc, bsdf_data, mat);
  }
  V2 rnd_bsdf = s.smp.next_2d();
  V2 rnd_em_sample = s.smp.next_2d();
  V2 rnd_support = s.smp.next_2d();
  if (pt.blue_noise && (s.path_length == 1u)) {
    rnd_bsdf = sample_blue_noise(sc, px, py, pt.iteration, 0);
    rnd_em_sample = sample_blue_noise(sc, px, py, pt.iteration, 2);
    rnd_support = sample_blue_noise(sc, px, py, pt.iteration, 4);
  }
  s.smp.push_fixed(rnd_bsdf.x, rnd_bsdf.y, rnd_support.x);
  BSample<SP> bs = bsdf_sample<SP>(sc, bsdf_data, mat, s.smp);
  s.smp.pop_fixed();

  bool subsurface_path = false, subsurface_sampled = false;
  SSGather<SP> ssg;
  ssg.count = 0;
  if constexpr (!PLAIN) {
    subsurface_path = (sc.has_subsurface != 0u) && (mat.subsurface.cls != 0u) && (bs.properties & kBsdfReflection) && (bs.properties & kBsdfDiffuse);
    if (subsurface_path) subsurface_sampled = ss_gather<SP>(sc, s.wavelength, isect, s.smp, ssg, stats, shadow_rays);
  }
  if (subsurface_path && (subsurface_sampled == false)) return false;
  if (bs.valid() == false) return false;

  s.medium = (bs.properties & kBsdfMediumChanged) ? bs.medium_index : s.medium;

  // direct light sampling (:400-425)
  if (pt.nee && (s.path_length + 1u <= sc.max_path_length)) {
    s.smp.push_fixed(rnd_em_sample.x, rnd_em_sample.y, rnd_support.x);
    uint32_t emitter_index = distribution_sample(sc.emitter_dist, sc.emitter_count + 1u, rnd_support.y);
    Spec<SP> direct_light = Spec<SP>::make(0.0f);
    if (subsurface_sampled) {
      const etxb_material& exit_mat = sc.materials[sc.subsurface_exit_material];
      for (uint32_t i = 0; i < ssg.count; ++i) {
        Isect exit_isect = ss_exit_intersection<SP>(sc, ssg, i, sc.subsurface_exit_material);
        EmitterSample<SP> local = sample_emitter<SP>(sc, s.wavelength, emitter_index, rnd_em_sample, exit_isect.pos);
        Spec<SP> light_value = pt_evaluate_light<SP, PLAIN>(sc, exit_isect, exit_mat, s.medium, s.wavelength, local, s.smp, pt.mis != 0u, stats, shadow_rays, nullptr, direct_light);
        direct_light += ssg.weights[i] * light_value;
      }
    } else {
      EmitterSample<SP> es = sample_emitter<SP>(sc, s.wavelength, emitter_index, rnd_em_sample, isect.pos);
      direct_light += pt_evaluate_light<SP, PLAIN>(sc, isect, mat, s.medium, s.wavelength, es, s.smp, pt.mis != 0u, stats, shadow_rays, batch, s.throughput);
    }
    s.accumulated += s.throughput * direct_light;
    s.smp.pop_fixed();
  }

  if (subsurface_sampled) {
    Isect out = ss_exit_intersection<SP>(sc, ssg, ssg.selected, sc.subsurface_exit_material);
    s.ray_d = sample_cosine_around(rnd_bsdf, out.nrm, 1.0f);
    s.throughput *= ssg.weights[ssg.selected] * ssg.selected_sample_weight;
    s.sampled_bsdf_pdf = fabsf(dot(s.ray_d, out.nrm)) / kPi;
    s.mis_weight = true;
    s.ray_o = shading_pos(sc, load_triangle(sc, out.triangle_index), out.barycentric, s.ray_d);
  } else {
    s.throughput *= bs.weight;
    s.sampled_bsdf_pdf = bs.pdf;
    s.mis_weight = bs.is_delta() == false;
    s.eta *= bs.eta;
    s.ray_d = bs.w_o;
    s.ray_o = shading_pos(sc, tri, isect.barycentric, s.ray_d);
  }
  if (s.throughput.is_zero()) return false;
  s.ray_max_t = kMaxFloat;
  s.ray_min_t = kRayEpsilon;
  s.path_length += 1u;
  return random_continue<SP>(s.path_length, sc.random_path_termination, s.eta, s.smp, s.throughput);
}

}  // namespace etxb
