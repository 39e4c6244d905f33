// dclosure.cuh — vertex closures: the product build's form of bsdf::evaluate + bsdf::reverse_pdf for the stages that evaluate ONE vertex
// against MANY directions (photon gather: ~20 photons per camera vertex; vertex connections: every vertex of the paired light path).
//
// The reference evaluates a BSDF as `evaluate(data, w_o, material)` followed by `reverse_pdf(...)` (vcm_shared.hxx:673-763, 841-857), and
// every one of those calls starts from the Material record: two RefractiveIndex lookups (four SpectralDistribution queries), the thin-film
// evaluation, the roughness and reflectance / scattering images (scene_bsdf.hxx:56-126, bsdf_plastic.hxx:96-160, bsdf_conductor.hxx:60-110).
// PlasticBSDF::evaluate alone repeats that preparation three times and computes the same Fresnel term three times (the layer weight, the
// specular pdf, and again for the reverse pdf) — measured on the 1M-triangle room, the photon gather spent 180 of 374 ms per iteration there
// (profiles/r1b_c3_k_camera_merge_generic_batched.raw.csv).  A closure does the direction-independent part once per vertex; the per-direction
// part keeps the reference's estimator (the same stochastic microsurface walk, bsdf_external.hxx:281-420,466-580) and returns value, pdf and
// reverse pdf together, sharing the half-vector terms.  Classes without a closure form (Translucent, Velvet, Principled, the rough / vMF
// diffuse variations, Mirror) go through the generic routines of dbsdf.cuh unchanged.
//
// The parity build does not use this header: there every call keeps the reference's order of sampler draws.
#pragma once
#include "dbsdf.cuh"

namespace etxb {

enum : uint32_t { kClLambert = 0u, kClPlastic = 1u, kClConductor = 2u, kClDielectric = 3u, kClGeneric = 4u, kClNone = 5u };

template <bool SP>
struct Closure {
  V3 nrm, tan, btn, w_i;  // shading frame as interpolated (not flipped) + incoming direction (BSDFData::w_i: pointing AT the surface)
  V2 tex;
  float wavelength;
  uint32_t kind, material, medium, path_source;
  V2 alpha;
  IorSample<SP> ext, inte;
  ThinfilmEval<SP> film;
  Spec<SP> refl, scat;  // reflectance (specular tint) and scattering (diffuse albedo / transmission tint) at this wavelength
};

template <bool SP>
struct CEval {
  Spec<SP> func, bsdf;
  float pdf, rev_pdf;
  DEV bool valid() const { return pdf > 0.0f; }
};
template <bool SP>
DEV CEval<SP> ceval_zero() {
  return {Spec<SP>::make(0.0f), Spec<SP>::make(0.0f), 0.0f, 0.0f};
}

template <bool SP>
DEV BData closure_bdata(const Closure<SP>& c) {
  return {{0.0f, 0.0f, 0.0f}, c.nrm, c.tan, c.btn, c.tex, c.w_i, c.wavelength, c.path_source, c.medium};
}

// the direction-independent half: one pass over the Material record
template <bool SP>
DEV Closure<SP> make_closure(const DeviceScene& sc, const BData& d, uint32_t material_index, Smp& smp) {
  Closure<SP> c;
  c.nrm = d.nrm;
  c.tan = d.tan;
  c.btn = d.btn;
  c.w_i = d.w_i;
  c.tex = d.tex;
  c.wavelength = d.wavelength;
  c.material = material_index;
  c.medium = d.current_medium;
  c.path_source = d.path_source;
  c.alpha = {0.0f, 0.0f};
  c.refl = Spec<SP>::make(0.0f);
  c.scat = Spec<SP>::make(0.0f);
  c.ext.cls = c.inte.cls = 0u;
  c.ext.eta = c.inte.eta = Spec<SP>::make(1.0f);
  c.ext.k = c.inte.k = Spec<SP>::make(0.0f);
  c.film.thickness = 0.0f;
  c.film.ior = c.ext;
  c.film.rgb_wavelengths = {610.0f, 537.0f, 450.0f};
  const etxb_material& m = sc.materials[material_index];
  const bool lambert_base = m.diffuse_variation == 0u;
  if ((m.cls == ETXB_MAT_DIFFUSE) && lambert_base) {
    c.kind = kClLambert;
    c.scat = apply_image<SP>(sc, m.scattering, d.tex, d.wavelength);
    return c;
  }
  const bool walk = ((m.cls == ETXB_MAT_PLASTIC) && lambert_base) || (m.cls == ETXB_MAT_CONDUCTOR) || (m.cls == ETXB_MAT_DIELECTRIC);
  if (!walk) {
    c.kind = ((m.cls == ETXB_MAT_THINFILM) || (m.cls == ETXB_MAT_BOUNDARY) || (m.cls == ETXB_MAT_VOID)) ? kClNone : kClGeneric;  // those three evaluate to zero
    return c;
  }
  c.kind = (m.cls == ETXB_MAT_PLASTIC) ? kClPlastic : ((m.cls == ETXB_MAT_CONDUCTOR) ? kClConductor : kClDielectric);
  c.alpha = evaluate_roughness(sc, m, d.tex);
  c.ext = evaluate_ior<SP>(sc, m.ext_ior, d.wavelength);
  c.inte = evaluate_ior<SP>(sc, m.int_ior, d.wavelength);
  c.film = evaluate_thinfilm<SP>(sc, d.wavelength, m.thinfilm, d.tex, smp);
  c.refl = apply_image<SP>(sc, m.reflectance, d.tex, d.wavelength);
  if (m.cls != ETXB_MAT_CONDUCTOR) c.scat = apply_image<SP>(sc, m.scattering, d.tex, d.wavelength);
  return c;
}

// visible-normal sampling density of the specular lobe around the half vector, the part that depends on which side is "incoming"
// (PlasticBSDF::pdf / DielectricBSDF::pdf: D * G1 * |i.h| / |i.n|)
DEV float closure_vndf_prob(V3 w, V3 wh, float dg, V2 alpha) {
  MicroRay ray = micro_ray(w, alpha);
  return tmax(0.0f, dot(wh, ray.w) * dg / ((1.0f + ray.Lambda) * ray.w.z));
}

// DielectricBSDF::pdf on local directions (bsdf_dielectric.hxx:219-259), the preparation already done
template <bool SP>
DEV float closure_dielectric_pdf(const Closure<SP>& c, V3 w_i, V3 w_o) {
  const bool outside = w_i.z > 0.0f;
  const bool reflection = w_i.z * w_o.z > 0.0f;
  V3 wh;
  float dwh_dwo;
  if (reflection) {
    wh = normalize(w_o + w_i);
    dwh_dwo = 1.0f / (4.0f * dot(w_o, wh));
  } else {
    float eta = outside ? (c.inte.eta / c.ext.eta).monochromatic() : (c.ext.eta / c.inte.eta).monochromatic();
    wh = normalize(w_i + w_o * eta);
    float sqrt_denom = dot(w_i, wh) + eta * dot(w_o, wh);
    dwh_dwo = sqr(eta) * dot(w_o, wh) / sqr(sqrt_denom);
  }
  wh *= (wh.z >= 0.0f) ? 1.0f : -1.0f;
  float prob = closure_vndf_prob(w_i * (outside ? 1.0f : -1.0f), wh, d_ggx(wh, c.alpha), c.alpha);
  float f = fresnel_calculate<SP>(c.wavelength, dot(w_i, wh), outside ? c.ext : c.inte, outside ? c.inte : c.ext, c.film).monochromatic();
  prob *= reflection ? f : (1.0f - f);
  return fabsf(prob * dwh_dwo) + fabsf(w_o.z);
}

// value, pdf and reverse pdf of the closure for the outgoing direction w_o (world space, pointing away from the surface)
template <bool SP>
DEV CEval<SP> closure_evaluate(const DeviceScene& sc, const Closure<SP>& c, V3 w_o, Smp& smp) {
  CEval<SP> e = ceval_zero<SP>();
  if (c.kind == kClNone) return e;
  const bool entering = dot(c.nrm, c.w_i) < 0.0f;
  if (c.kind == kClLambert) {
    // DiffuseBSDF::evaluate / pdf (bsdf_various.hxx:36-133)
    const V3 fn = entering ? c.nrm : -c.nrm;
    const float cos_o = dot(fn, w_o);
    if (cos_o <= kEpsilon) return e;
    e.func = c.scat / kPi;
    e.bsdf = e.func * cos_o;
    e.pdf = kInvPi * cos_o;
    const float facing = (dot(c.nrm, w_o) > 0.0f) ? dot(c.nrm, -c.w_i) : -dot(c.nrm, -c.w_i);  // normal facing -w_o, against the reversed outgoing -w_i
    e.rev_pdf = (facing <= kEpsilon) ? 0.0f : kInvPi * facing;
    return e;
  }
  if (c.kind == kClGeneric) {
    const etxb_material& m = sc.materials[c.material];
    BData d = closure_bdata(c);
    BEval<SP> b = bsdf_evaluate<SP>(sc, d, w_o, m, smp);
    if (b.valid() == false) return e;
    e.func = b.func;
    e.bsdf = b.bsdf;
    e.pdf = b.pdf;
    e.rev_pdf = bsdf_reverse_pdf<SP>(sc, d, w_o, m, smp);
    return e;
  }
  const Frame lf = {c.tan, c.btn, c.nrm, false};
  if (c.kind == kClPlastic) {
    // PlasticBSDF::evaluate + ::pdf of the reversed pair (bsdf_plastic.hxx:96-160): Lambert base under a dielectric microsurface
    const V3 n = entering ? c.nrm : -c.nrm;
    const V3 mh = normalize(w_o - c.w_i);
    const float n_dot_o = dot(n, w_o), m_dot_o = dot(mh, w_o);
    if ((n_dot_o <= kEpsilon) || (m_dot_o <= kEpsilon)) return e;
    const Spec<SP> fr = fresnel_calculate<SP>(c.wavelength, dot(c.w_i, mh), c.ext, c.inte, c.film);  // |i.h| = |o.h|: one Fresnel term serves both directions
    const Spec<SP> tr = 1.0f - fr;
    const float tr_m = tr.monochromatic(), fr_m = fr.monochromatic();
    const Spec<SP> diff_func = c.scat / kPi;  // local_w_o.z = n_dot_o > 0 here
    Spec<SP> spec = Spec<SP>::make(0.0f);
    float spec_pdf = 0.0f, spec_rev = 0.0f;
    const V3 wi_u = lf.to_local(-c.w_i), wo_u = lf.to_local(w_o);
    if ((wi_u.z > kEpsilon) && (wo_u.z > kEpsilon)) {
      spec = 2.0f * eval_dielectric<SP>(c.wavelength, smp, wi_u, wo_u, true, c.alpha, c.ext, c.inte, c.film) * c.refl;
      const V3 wh = normalize(wo_u + wi_u);
      const float dg = d_ggx(wh, c.alpha);
      const float dwh = 1.0f / (4.0f * dot(wo_u, wh));
      spec_pdf = fabsf(closure_vndf_prob(wi_u, wh, dg, c.alpha) * fr_m * dwh);
      spec_rev = fabsf(closure_vndf_prob(wo_u, wh, dg, c.alpha) * fr_m * dwh);
    }
    e.func = diff_func * tr + spec / n_dot_o;
    e.bsdf = diff_func * tr * n_dot_o + spec;
    e.pdf = kInvPi * n_dot_o * tr_m + spec_pdf;
    // reversed pair (w_i' = -w_o, w_o' = -w_i): its frame normal faces -(-w_o) = w_o, the same half vector, the same Fresnel term
    const V3 n_rev = (dot(c.nrm, w_o) > 0.0f) ? c.nrm : -c.nrm;
    const float n_dot_i = dot(n_rev, -c.w_i), m_dot_i = dot(mh, -c.w_i);
    e.rev_pdf = ((n_dot_i <= kEpsilon) || (m_dot_i <= kEpsilon)) ? 0.0f : (kInvPi * n_dot_i * tr_m + spec_rev);
    return e;
  }
  if (c.kind == kClConductor) {
    // ConductorBSDF::evaluate / pdf (bsdf_conductor.hxx:60-110) in the frame flipped towards the incoming side
    const float s = entering ? 1.0f : -1.0f;
    const V3 wo_l = lf.to_local(w_o) * s, wi_l = lf.to_local(-c.w_i) * s;
    if ((wo_l.z <= kEpsilon) || (wi_l.z <= kEpsilon)) return e;
    Spec<SP> value = eval_conductor<SP>(c.wavelength, smp, wi_l, wo_l, c.alpha, c.ext, c.inte, c.film);
    e.bsdf = value * c.refl;
    e.func = e.bsdf / wo_l.z;
    const float dg4 = d_ggx(normalize(wo_l + wi_l), c.alpha) * 0.25f;
    MicroRay ri = micro_ray(wi_l, c.alpha), ro = micro_ray(wo_l, c.alpha);
    e.pdf = dg4 / (1.0f + ri.Lambda) / wi_l.z + wo_l.z;
    e.rev_pdf = dg4 / (1.0f + ro.Lambda) / wo_l.z + wi_l.z;
    return e;
  }
  // kClDielectric: DielectricBSDF::evaluate / pdf (bsdf_dielectric.hxx:170-259), frame as interpolated
  const V3 w_i = lf.to_local(-c.w_i), w_ol = lf.to_local(w_o);
  if ((fabsf(w_i.z) <= kEpsilon) || (fabsf(w_ol.z) <= kEpsilon)) return e;
  const bool forward_path = c.path_source == kPathCamera;
  const float backward_scale = fabsf(1.0f / w_i.z);
  const float wl = c.wavelength;
  Spec<SP> value;
  if (w_i.z > 0) {
    if (w_ol.z >= 0) {
      value = forward_path ? eval_dielectric<SP>(wl, smp, w_i, w_ol, true, c.alpha, c.ext, c.inte, c.film)
                           : eval_dielectric<SP>(wl, smp, w_ol, w_i, true, c.alpha, c.ext, c.inte, c.film) * backward_scale;
    } else {
      value = forward_path ? eval_dielectric<SP>(wl, smp, w_i, w_ol, false, c.alpha, c.ext, c.inte, c.film)
                           : eval_dielectric<SP>(wl, smp, -w_ol, -w_i, false, c.alpha, c.inte, c.ext, c.film) * backward_scale;
    }
  } else if (w_ol.z <= 0) {
    value = forward_path ? eval_dielectric<SP>(wl, smp, -w_i, -w_ol, true, c.alpha, c.inte, c.ext, c.film)
                         : eval_dielectric<SP>(wl, smp, -w_ol, -w_i, true, c.alpha, c.inte, c.ext, c.film) * backward_scale;
  } else {
    value = forward_path ? eval_dielectric<SP>(wl, smp, -w_i, -w_ol, false, c.alpha, c.inte, c.ext, c.film)
                         : eval_dielectric<SP>(wl, smp, w_ol, w_i, false, c.alpha, c.ext, c.inte, c.film) * backward_scale;
  }
  if (value.is_zero()) return e;
  const bool reflection = w_i.z * w_ol.z > 0.0f;
  e.func = (2.0f * value) * (reflection ? c.refl : c.scat);
  e.bsdf = e.func * fabsf(w_ol.z);
  e.pdf = closure_dielectric_pdf<SP>(c, w_i, w_ol);
  e.rev_pdf = closure_dielectric_pdf<SP>(c, w_ol, w_i);
  return e;
}

}  // namespace etxb
