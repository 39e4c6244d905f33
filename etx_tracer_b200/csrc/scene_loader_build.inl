// scene_loader_build.inl — from parsed files to the Scene / Camera PODs (included by scene_loader.cpp inside its anonymous namespace).

struct ImageRecord {
  Pixels px;              // u8 (RGBA8) or f32 (RGBA32F) pixels, rows in file order
  uint32_t options = 0;
  float offset[2] = {0.0f, 0.0f}, scale[2] = {1.0f, 1.0f};
  std::vector<std::vector<etxb_distribution_entry>> x_entries;
  std::vector<etxb_distribution> x_dists;
  std::vector<etxb_distribution_entry> y_entries;
  float y_total = 0.0f, normalization = 0.0f;
  void pixel(uint32_t x, uint32_t y, float out[4]) const {  // Image::pixel (image.hxx:90-103)
    size_t i = std::min<size_t>(size_t(x) + size_t(y) * px.w, size_t(px.w) * px.h - 1u);
    if (px.eight_bit) {
      for (int c = 0; c < 4; ++c) out[c] = px.u8[i * 4 + c] / 255.0f;
    } else {
      memcpy(out, px.f32.data() + i * 4, 16);
    }
  }
  float tex_coord(float u, float size, bool repeat) const {  // image.hxx:158-172
    if (repeat) {
      float x = fmodf(u, size);
      return x < 0.0f ? (x + size) : x;
    }
    return clampf(u, 0.0f, nextafterf(size, 0.0f));
  }
  // Image::read (image.hxx:174-186): the four texels around a pixel-space position
  void read(float ux, float uy, float out[4]) const {
    const float fw = float(px.w), fh = float(px.h);
    const bool ru = options & IMG_REPEAT_U, rv = options & IMG_REPEAT_V;
    float x0 = tex_coord(ux - 0.0f, fw, ru), x1 = tex_coord(ux + 1.0f, fw, ru), y0 = tex_coord(uy - 0.0f, fh, rv), y1 = tex_coord(uy + 1.0f, fh, rv);
    float dx = x0 - floorf(x0), dy = y0 - floorf(y0);
    float p00[4], p01[4], p10[4], p11[4];
    pixel(uint32_t(x0), uint32_t(y0), p00), pixel(uint32_t(x1), uint32_t(y0), p01), pixel(uint32_t(x0), uint32_t(y1), p10), pixel(uint32_t(x1), uint32_t(y1), p11);
    for (int c = 0; c < 4; ++c) out[c] = p00[c] * (1.0f - dx) * (1.0f - dy) + p01[c] * (dx) * (1.0f - dy) + p10[c] * (1.0f - dx) * (dy) + p11[c] * (dx) * (dy);
  }
  // Image::evaluate(uv, nullptr) (image.hxx:51-88)
  void evaluate(float u, float v, float out[4]) const {
    const float fw = float(px.w), fh = float(px.h);
    float x0 = tex_coord(u * fw, fw, options & IMG_REPEAT_U), y0 = tex_coord(v * fh, fh, options & IMG_REPEAT_V);
    float dx = x0 - floorf(x0), dy = y0 - floorf(y0);
    uint32_t r0 = std::min(uint32_t(y0), px.h - 1u), r1 = std::min(r0 + 1u, px.h - 1u), c0 = std::min(uint32_t(x0), px.w - 1u), c1 = std::min(c0 + 1u, px.w - 1u);
    float p00[4], p01[4], p10[4], p11[4];
    pixel(c0, r0, p00), pixel(c1, r0, p01), pixel(c0, r1, p10), pixel(c1, r1, p11);
    for (int c = 0; c < 4; ++c) out[c] = p00[c] * (1.0f - dx) * (1.0f - dy) + p01[c] * (dx) * (1.0f - dy) + p10[c] * (1.0f - dx) * (dy) + p11[c] * (dx) * (dy);
  }
};

// DistributionBuilder::finalize (distribution_builder.hxx:30-58): size + 1 entries, `total` = the float sum in index order
float finalize_distribution(std::vector<etxb_distribution_entry>& e, uint32_t size) {
  float total = 0.0f;
  for (uint32_t i = 0; i < size; ++i) {
    e[i].cdf = total;
    total += e[i].value;
  }
  if (total == 0.0f) {
    for (uint32_t i = 0; i < size; ++i) e[i] = {1.0f, 1.0f / float(size), float(i) / float(size)};
  } else {
    for (uint32_t i = 0; i < size; ++i) {
      e[i].pdf = e[i].value / total;
      e[i].cdf /= total;
    }
  }
  e[size] = {0.0f, 0.0f, 1.0f};
  return total;
}

// build_image_sampling_table (render/host/image_pool.cxx:226-259)
void build_sampling_table(ImageRecord& img) {
  const uint32_t w = img.px.w, h = img.px.h;
  const bool uniform = img.options & IMG_UNIFORM_TABLE;
  img.x_entries.assign(h, {});
  img.x_dists.assign(h, {});
  img.y_entries.assign(h + 1, {});
  float total_weight = 0.0f;
  const float fw = float(w), fh = float(h);
  for (uint32_t y = 0; y < h; ++y) {
    float v = (float(y) + 0.5f) / fh, row_value = 0.0f;
    auto& e = img.x_entries[y];
    e.assign(w + 1, {});
    for (uint32_t x = 0; x < w; ++x) {
      float u = (float(x) + 0.5f) / fw, px[4];
      img.read(fw * u, fh * v, px);
      float lum = luminance({px[0], px[1], px[2]});
      row_value += lum;
      e[x].value = lum;
    }
    img.x_dists[y].total_weight = finalize_distribution(e, w);
    img.x_dists[y].values = {e.data(), w};  // the reference counts `size` entries and owns size + 1
    float row_weight = uniform ? 1.0f : sinf(v * kPiF);
    row_value *= row_weight;
    total_weight = total_weight + row_value;
    img.y_entries[y].value = row_value;
  }
  img.y_total = finalize_distribution(img.y_entries, h);
  img.normalization = total_weight / (fw * fh);
}

#include "scene_loader_atmosphere.inl"

struct CameraBlock {
  uint32_t cls = 0, viewport[2] = {0, 0}, lens_image = kInvalid, medium = kInvalid;
  bool has_origin = false, has_target = false, has_near = false, has_far = false, active = false;
  F3 origin, target, up = {0.0f, 1.0f, 0.0f};
  float fov = 50.0f, lens_radius = 0.0f, focal_distance = 0.0f, clip_near = 0.0f, clip_far = 0.0f;
};

struct etxb_scene_file_impl {
  std::shared_ptr<Tables> tables;
  std::string data_folder;  // where tables.bin was found (the optical-length table of the atmosphere is cached beside it)
  etxb_scene scene = {};
  etxb_camera camera = {};
  std::vector<etxb_vertex> vertices;
  std::vector<etxb_triangle> triangles;
  std::vector<uint32_t> tri_to_emitter;
  std::vector<etxb_material> materials;
  std::vector<std::string> material_names;
  std::vector<Spd> spectra;
  std::vector<std::unique_ptr<ImageRecord>> images;
  std::vector<etxb_image> image_pods;
  std::vector<etxb_medium> mediums;
  std::list<std::vector<float>> densities;  // the dense grids of heterogeneous media (addresses stay put)
  std::vector<etxb_emitter_profile> profiles;   // distant ones first (declared in the material file), area ones appended by commit
  std::vector<etxb_emitter> emitters;
  std::vector<etxb_distribution_entry> emitter_dist;
  std::vector<std::string> warnings;
  std::map<std::string, uint32_t> named_spectra, medium_names, material_index, image_cache;
  std::vector<CameraBlock> cameras;
  std::string base_dir;
  uint32_t black = 0, white = 0, rayleigh = 0, mie = 0, ozone = 0, def_diel = 0, def_cond_eta = 0, def_cond_k = 0, ss_scatter = 0, ss_exit = 0;

  const Tables& t() const { return *tables; }
  void warn(const std::string& s) { warnings.push_back(s); }
  uint32_t add_spectrum(const Spd& s) {
    spectra.push_back(s);
    return uint32_t(spectra.size() - 1);
  }

  etxb_material blank_material() const {
    etxb_material m = {};
    m.reflectance = m.scattering = m.emission = {kInvalid, kInvalid};
    m.roughness.image_index = m.metalness.image_index = m.transmission.image_index = kInvalid;
    m.roughness.channel = m.metalness.channel = m.transmission.channel = kInvalid;
    m.subsurface.spectrum_index = m.subsurface.image_index = kInvalid;
    m.thinfilm.ior.eta_index = m.thinfilm.ior.k_index = kInvalid;
    m.thinfilm.thickness_image = kInvalid;
    m.ext_ior.eta_index = m.ext_ior.k_index = m.int_ior.eta_index = m.int_ior.k_index = kInvalid;
    m.cls = ETXB_MAT_DIFFUSE;
    m.int_medium = m.ext_medium = kInvalid;
    m.normal_image_index = kInvalid;
    m.normal_scale = 1.0f;
    m.opacity = 1.0f;
    return m;
  }
  uint32_t add_material(const std::string& name) {
    materials.push_back(blank_material());
    material_names.push_back(name);
    material_index[name] = uint32_t(materials.size() - 1);
    return uint32_t(materials.size() - 1);
  }

  // SceneRepresentationImpl::init_default_values (:206-225)
  void init_default_values() {
    black = add_spectrum(spd_rgb_reflectance(t(), {0.0f, 0.0f, 0.0f}));
    white = add_spectrum(spd_rgb_reflectance(t(), {1.0f, 1.0f, 1.0f}));
    auto atmosphere = [&](const char* key) {
      return add_spectrum(spd_from_power(t().get(std::string("spectra/atmosphere_") + key + ".power").data(), t().get(std::string("spectra/atmosphere_") + key + ".rgb").data()));
    };
    rayleigh = atmosphere("rayleigh"), mie = atmosphere("mie"), ozone = atmosphere("ozone");
    def_diel = add_spectrum(spd_constant(1.5f));
    def_cond_eta = add_spectrum(spd_constant(0.0f));
    def_cond_k = add_spectrum(spd_constant(1000000.0f));
    ss_scatter = add_material("etx::subsurface-scatter");
    materials[ss_scatter].cls = ETXB_MAT_TRANSLUCENT;
    materials[ss_scatter].reflectance.spectrum_index = black;
    materials[ss_scatter].scattering.spectrum_index = white;
    ss_exit = add_material("etx::subsurface-exit");
    materials[ss_exit].reflectance.spectrum_index = white;
    materials[ss_exit].scattering.spectrum_index = white;
  }

  // get_file (:114-154): relative to the material file's folder, else as given
  // get_file (:114-121): "<folder of the material file>/<name>" for any non-empty name — the file is not looked for here; one that cannot be read
  // becomes the 1 x 1 white placeholder in add_image_file, like in the reference's texture pool
  bool find_file(const std::string& name, std::string& out) {
    if (name.empty()) return false;
    out = base_dir.empty() ? name : (base_dir + "/" + name);
    return true;
  }

  // ImagePool::add_from_file + load_image (image_pool.cxx:51-66, 162-215)
  uint32_t add_image_file(const std::string& path, uint32_t options, float off_x = 0.0f, float off_y = 0.0f, float sc_x = 1.0f, float sc_y = 1.0f) {
    auto hit = image_cache.find(path);
    if (hit != image_cache.end()) return hit->second;
    auto rec = std::make_unique<ImageRecord>();
    options |= IMG_PERFORM_LOADING;
    std::string problem;
    bool loaded = false;
    try {
      rec->px = read_image(path);
      loaded = true;
    } catch (const LoadError& e) {
      problem = e.text;
    } catch (const std::exception& e) {  // a container refusing a size a corrupt header asked for
      problem = path + ": " + e.what();
    }
    if (!loaded) {
      warn(problem + "; using the 1x1 white placeholder");
      rec->px = Pixels();
      rec->px.w = rec->px.h = 1;
      rec->px.f32 = {1.0f, 1.0f, 1.0f, 1.0f};
      options |= IMG_SKIP_SRGB | IMG_REPEAT_U | IMG_REPEAT_V;
    }
    Pixels& px = rec->px;
    if (px.eight_bit) {
      const bool convert = (options & IMG_SKIP_SRGB) == 0;
      for (size_t i = 0; i < size_t(px.w) * px.h; ++i) {
        float f[4];
        for (int c = 0; c < 4; ++c) f[c] = px.u8[i * 4 + c] / 255.0f;
        if (convert) {
          F3 lin = gamma_to_linear({f[0], f[1], f[2]});
          f[0] = lin.x, f[1] = lin.y, f[2] = lin.z;
        }
        for (int c = 0; c < 4; ++c) px.u8[i * 4 + c] = static_cast<uint8_t>(saturate(f[c]) * 255.0f);  // to_ubyte4 truncates (math.hxx:713-720)
      }
    } else {
      for (float& v : px.f32) {
        if (std::isinf(v)) v = 65504.0f;
        if (std::isnan(v) || (v < 0.0f)) v = 0.0f;
      }
    }
    float pxv[4];
    for (size_t i = 0; i < size_t(px.w) * px.h; ++i) {
      rec->pixel(uint32_t(i % px.w), uint32_t(i / px.w), pxv);
      if (pxv[3] < 1.0f) {
        options |= IMG_HAS_ALPHA;
        break;
      }
    }
    rec->options = options;
    rec->offset[0] = off_x, rec->offset[1] = off_y, rec->scale[0] = sc_x, rec->scale[1] = sc_y;
    if (options & IMG_BUILD_TABLE) build_sampling_table(*rec);
    images.push_back(std::move(rec));
    image_cache[path] = uint32_t(images.size() - 1);
    return uint32_t(images.size() - 1);
  }
  uint32_t add_image_data(Pixels&& px, uint32_t options) {  // ImagePool::add_from_data (:69-92)
    auto rec = std::make_unique<ImageRecord>();
    rec->px = std::move(px);
    rec->options = options;
    if (options & IMG_BUILD_TABLE) build_sampling_table(*rec);
    images.push_back(std::move(rec));
    return uint32_t(images.size() - 1);
  }

  // ---- spectra directives -----------------------------------------------------------------------------------------------------------------------------
  uint32_t reflectance_spectrum(const std::string& text) {  // load_reflectance_spectrum (:1613-1633)
    auto p = split_params(text);
    if (p.size() == 1 && named_spectra.count(p[0])) return named_spectra[p[0]];
    if (p.size() == 3) return add_spectrum(spd_rgb_reflectance(t(), gamma_to_linear({c_atof(p[0]), c_atof(p[1]), c_atof(p[2])})));
    return 0;
  }
  Spd illuminant_spectrum(const std::string& text) {  // load_illuminant_spectrum (:1635-1680)
    auto p = split_params(text);
    if (p.size() == 1) {
      auto fl = leading_floats(p[0], 1);
      if (!fl.empty()) return spd_rgb_luminance(t(), {fl[0], fl[0], fl[0]});
      if (named_spectra.count(p[0])) return spectra[named_spectra[p[0]]];
    }
    if (p.size() == 3) return spd_rgb_luminance(t(), {c_atof(p[0]), c_atof(p[1]), c_atof(p[2])});
    Spd spd = spd_rgb_luminance(t(), {1.0f, 1.0f, 1.0f});
    float scale = 1.0f;
    for (size_t i = 0; i < p.size(); ++i) {
      if (p[i] == "blackbody" && i + 1 < p.size()) {
        spd = spd_black_body(t(), c_atof(p[i + 1]), 1.0f);
        i += 1;
      } else if (p[i] == "nblackbody" && i + 1 < p.size()) {
        spd = spd_normalized_black_body(t(), c_atof(p[i + 1]), 1.0f);
        i += 1;
      } else if (p[i] == "scale" && i + 1 < p.size()) {
        scale = c_atof(p[i + 1]);
        i += 1;
      }
    }
    spd_scale(spd, scale);
    return spd;
  }
  bool named_ior(const std::string& name_in, Spd& eta, Spd& k, uint32_t& cls) {
    std::string name = lower(trim(name_in));
    if (!t().has("spectra/" + name + ".eta_power")) return false;
    eta = spd_from_power(t().get("spectra/" + name + ".eta_power").data(), t().get("spectra/" + name + ".eta_rgb").data());
    k = spd_from_power(t().get("spectra/" + name + ".k_power").data(), t().get("spectra/" + name + ".k_rgb").data());
    cls = tables->u32.count("spectra/" + name + ".cls") ? tables->u32.at("spectra/" + name + ".cls")[0] : uint32_t(SPD_DIELECTRIC);
    return true;
  }
  void load_ior(etxb_refractive_index& target, const std::string& text) {  // the load_ior lambda of parse_material (:1846-1884)
    auto fl = leading_floats(text, 2);
    if (fl.size() == 1) {
      target.cls = SPD_DIELECTRIC;
      target.eta_index = add_spectrum(spd_constant(fl[0]));
      target.k_index = kInvalid;
    } else if (fl.size() == 2) {
      target.cls = SPD_CONDUCTOR;
      target.eta_index = add_spectrum(spd_constant(fl[0]));
      target.k_index = add_spectrum(spd_constant(fl[1]));
    } else {
      Spd eta, k;
      uint32_t cls = SPD_DIELECTRIC;
      if (!named_ior(text, eta, k, cls)) {
        warn("unable to load IOR spectrum `" + text + "`, falling back to 1.5 dielectric");
        eta = spd_constant(1.5f), k = spd_constant(0.0f), cls = SPD_DIELECTRIC;
      }
      target.cls = cls;
      target.eta_index = add_spectrum(eta);
      target.k_index = add_spectrum(k);
    }
  }

  // ---- et:: blocks ------------------------------------------------------------------------------------------------------------------------------------------
  void parse_camera(const MtlBlock& b) {
    CameraBlock cam;
    std::string v;
    if (b.get("class", v)) cam.cls = (trim(v) == "eq") ? 1u : 0u;
    if (b.get("viewport", v)) {
      auto tk = split(v);
      if (tk.size() >= 2) cam.viewport[0] = uint32_t(strtoul(tk[0].c_str(), nullptr, 10)), cam.viewport[1] = uint32_t(strtoul(tk[1].c_str(), nullptr, 10));
    }
    auto vec3 = [&](const char* key, F3& out, bool* flag) {
      if (b.get(key, v)) {
        auto fl = leading_floats(v, 3);
        if (fl.size() == 3) {
          out = {fl[0], fl[1], fl[2]};
          if (flag) *flag = true;
        }
      }
    };
    vec3("origin", cam.origin, &cam.has_origin);
    vec3("target", cam.target, &cam.has_target);
    vec3("up", cam.up, nullptr);
    auto scalar = [&](const char* key, float& out, bool* flag) {
      if (b.get(key, v)) {
        auto fl = leading_floats(v, 1);
        if (!fl.empty()) {
          out = fl[0];
          if (flag) *flag = true;
        }
      }
    };
    scalar("fov", cam.fov, nullptr);
    float focal = 0.0f;
    bool has_focal = false;
    scalar("focal-length", focal, &has_focal);
    if (has_focal) cam.fov = (2.0f * atanf(36.0f / (2.0f * focal))) * 180.0f / kPiF;  // focal_length_to_fov (:612-614)
    scalar("lens-radius", cam.lens_radius, nullptr);
    scalar("focal-distance", cam.focal_distance, nullptr);
    scalar("clip-near", cam.clip_near, &cam.has_near);
    scalar("clip-far", cam.clip_far, &cam.has_far);
    if (b.get("shape", v)) {
      std::string f;
      if (find_file(trim(v), f)) cam.lens_image = add_image_file(f, IMG_BUILD_TABLE | IMG_UNIFORM_TABLE);
    }
    if (b.get("ext_medium", v)) cam.medium = medium_names.count(trim(v)) ? medium_names[trim(v)] : kInvalid;
    if (b.get("active", v)) {
      auto tk = split(v);
      cam.active = !tk.empty() && (atoi(tk[0].c_str()) != 0);
    }
    cameras.push_back(cam);
  }

  uint32_t add_medium(const Spd& s_a, const Spd& s_t, float g, bool explicit_connections) {  // SceneLoaderContext::add_medium (scene_data.hxx:125-147)
    etxb_medium m = {};
    m.absorption_index = add_spectrum(s_a);
    m.scattering_index = add_spectrum(s_t);
    m.max_sigma = spd_max_power(spectra[m.absorption_index]) + spd_max_power(spectra[m.scattering_index]);
    m.phase_function_g = g;
    m.enable_explicit_connections = explicit_connections ? 1 : 0;
    mediums.push_back(m);
    return uint32_t(mediums.size() - 1);
  }

  void parse_medium(const MtlBlock& b) {
    std::string v;
    if (!b.get("id", v)) {
      warn("medium does not have identifier - skipped");
      return;
    }
    std::string name = trim(v);
    float g = 0.0f;
    for (const char* key : {"g", "anisotropy"}) {
      if (b.get(key, v)) {
        auto fl = leading_floats(v, 1);
        if (!fl.empty()) g = fl[0];
      }
    }
    auto rgb = [&](const std::string& text, F3& out) {
      auto fl = leading_floats(text, 3);
      if (fl.size() == 3) {
        out = {fl[0], fl[1], fl[2]};
        return true;
      }
      if (!fl.empty()) {
        out = {fl[0], fl[0], fl[0]};
        return true;
      }
      return false;
    };
    Spd s_a = spd_constant(0.0f), s_t = spd_constant(0.0f);
    F3 c;
    for (const char* key : {"absorption", "absorbtion"})
      if (b.get(key, v) && rgb(v, c)) s_a = spd_rgb_reflectance(t(), c);
    if (b.get("scattering", v) && rgb(v, c)) s_t = spd_rgb_reflectance(t(), c);
    for (const char* key : {"rayleigh", "mie"}) {
      if (b.get(key, v)) {
        s_t = spectra[!strcmp(key, "rayleigh") ? rayleigh : mie];
        float scale = 1.0f;
        auto p = split_params(v);
        for (size_t i = 0; i < p.size(); ++i)
          if (p[i] == "scale" && i + 1 < p.size()) scale = c_atof(p[i + 1]);
        spd_scale(s_t, scale / spd_max_power(s_t));
      }
    }
    if (b.get("parametric", v)) {  // colour + distances -> absorption / scattering through subsurface::remap (scene_bssrdf_subsurface.hxx:17-44), :1254-1296
      F3 color = {1.0f, 1.0f, 1.0f}, dist = {0.25f, 0.25f, 0.25f};
      float scale = 1.0f;
      auto p = split_params(v);
      for (size_t i = 0; i < p.size(); ++i) {
        if (p[i] == "color" && i + 3 < p.size()) {
          color = {c_atof(p[i + 1]), c_atof(p[i + 2]), c_atof(p[i + 3])};
          i += 3;
        }
        if (i < p.size() && p[i] == "distance" && i + 1 < p.size()) {
          float d = c_atof(p[i + 1]);
          dist = {d, d, d};
          i += 1;
        }
        if (i < p.size() && p[i] == "distances" && i + 3 < p.size()) {
          dist = {c_atof(p[i + 1]), c_atof(p[i + 2]), c_atof(p[i + 3])};
          i += 3;
        }
        if (i < p.size() && p[i] == "scale" && i + 1 < p.size()) {
          scale = c_atof(p[i + 1]);
          i += 1;
        }
      }
      auto remap = [](float col, float distance, float& extinction, float& scattering) {
        constexpr float ra = 1.826052378200f, rb = 4.985111943850f + 0.12735595943800f, rc = 1.096861024240f, rd = 0.496310210422f, re = 4.231902997010f + 0.00310603949088f,
                        rf = 2.406029994080f;
        col = fmaxf(0.0f, col);
        float blend = powf(col, 0.25f);
        float albedo = (1.0f - blend) * ra * powf(atanf(rb * col), rc) + blend * rd * powf(atanf(re * col), rf);
        albedo = clampf(albedo, 0.0f, 1.0f - kEps);
        extinction = 1.0f / fmaxf(distance, 1.0f / 1024.0f);
        scattering = extinction * albedo;
      };
      F3 ext, sca;
      remap(color.x, scale * dist.x, ext.x, sca.x), remap(color.y, scale * dist.y, ext.y, sca.y), remap(color.z, scale * dist.z, ext.z, sca.z);
      s_t = spd_rgb_reflectance(t(), sca);
      s_a = spd_rgb_reflectance(t(), {fmaxf(0.0f, ext.x - sca.x), fmaxf(0.0f, ext.y - sca.y), fmaxf(0.0f, ext.z - sca.z)});
    }
    const bool explicit_connections = !b.has("enclosed");
    const uint32_t index = add_medium(s_a, s_t, g, explicit_connections);
    medium_names[name] = index;
    if (b.get("volume", v) && !trim(v).empty()) {  // MediumPool::add (medium_pool.cxx:41-59): a dense grid scaled to a maximum of 1
      const std::string file = join(base_dir, trim(v));
      const size_t dot = file.find_last_of('.');
      if (dot == std::string::npos || lower(file.substr(dot)) != ".nvdb") fail(file + ": only NanoVDB (.nvdb) volume files are read");
      DensityGrid grid = read_nvdb_density(file, warnings);
      float max_density = 0.0f;
      for (float f : grid.values) max_density = fmaxf(max_density, f);
      if (max_density > 0.0f) {
        for (float& f : grid.values) f /= max_density;
        densities.push_back(std::move(grid.values));
        etxb_medium& m = mediums[index];
        m.cls = 1u;
        m.density = {densities.back().data(), densities.back().size()};
        memcpy(m.dimensions, grid.dim, sizeof(m.dimensions));
      }
    }
  }

  void add_distant(uint32_t cls, const Spd& spd, uint32_t image, F3 direction, float angular_size) {
    etxb_emitter_profile p = {};
    p.emission = {add_spectrum(spd), image};
    store3(p.direction, direction);
    p.cls = cls;
    p.angular_size = angular_size;
    p.angular_size_cosine = 1.0f;
    etxb_emitter e = {};
    e.cls = cls;
    e.profile = uint32_t(profiles.size());
    e.triangle_index = kInvalid;
    profiles.push_back(p);
    emitters.push_back(e);
  }

  void parse_directional(const MtlBlock& b) {  // (:1308-1343)
    std::string v;
    Spd spd = b.get("color", v) ? illuminant_spectrum(v) : spd_rgb_luminance(t(), {1.0f, 1.0f, 1.0f});
    F3 d = {1.0f, 1.0f, 1.0f};
    if (b.get("direction", v)) {
      auto fl = leading_floats(v, 3);
      if (fl.size() == 3) d = {fl[0], fl[1], fl[2]};
    }
    d = normalize(d);
    uint32_t image = kInvalid;
    if (b.get("image", v)) {
      std::string f;
      if (find_file(trim(v), f)) image = add_image_file(f, 0);
    }
    float ang = 0.0f;
    if (b.get("angular_diameter", v)) {
      auto fl = leading_floats(v, 1);
      if (!fl.empty()) ang = fl[0] * kPiF / 180.0f;
    }
    add_distant(2u, spd, image, d, ang);
  }

  void parse_env(const MtlBlock& b) {  // (:1345-1378)
    std::string v, name;
    float rotation = 0.0f, u_scale = 1.0f;
    if (b.get("rotation", v)) rotation = -c_atof(v) / 360.0f;
    if (b.get("scale", v)) {
      auto fl = leading_floats(v, 1);
      if (!fl.empty()) u_scale = fl[0];
    }
    // a missing / unnamed image is the 1 x 1 white placeholder: a constant-colour environment (scene_data.hxx:104-107, image_pool.cxx:172-184)
    name = (b.get("image", v) && !trim(v).empty()) ? join(base_dir, trim(v)) : join(base_dir, "image-" + std::to_string(images.size()));
    uint32_t image = add_image_file(name, IMG_BUILD_TABLE | IMG_REPEAT_U, rotation, 0.0f, u_scale, 1.0f);
    Spd spd = b.get("color", v) ? illuminant_spectrum(v) : spd_rgb_luminance(t(), {1.0f, 1.0f, 1.0f});
    add_distant(1u, spd, image, {0.0f, 0.0f, 0.0f}, 0.0f);
  }

  // parse_atmosphere_light (:1376-1495): a sun (Directional, limb-darkened extinction image) and a sky (Environment, single-scattering image)
  void parse_atmosphere(const MtlBlock& b) {
    auto scalar = [&](const char* key, float fallback) {
      std::string v;
      if (b.get(key, v)) {
        auto fl = leading_floats(v, 1);
        if (!fl.empty()) return fl[0];
      }
      return fallback;
    };
    AtmosphereParameters prm;
    const float quality = scalar("quality", 1.0f), scale = scalar("scale", 1.0f), sun_scale = scalar("sun_scale", 1.0f), sky_scale = scalar("sky_scale", 1.0f);
    F3 direction = normalize(F3{1.0f, 1.0f, 1.0f});
    std::string v;
    if (b.get("direction", v)) {
      auto fl = leading_floats(v, 3);
      if (fl.size() == 3) direction = normalize(F3{fl[0], fl[1], fl[2]});
    }
    float angular_size = 0.5422f * (kPiF / 180.0f);
    if (b.get("angular_diameter", v)) {
      auto fl = leading_floats(v, 1);
      if (!fl.empty()) angular_size = fl[0] * (kPiF / 180.0f);
    }
    prm.anisotropy = scalar("anisotropy", prm.anisotropy);
    prm.altitude = scalar("altitude", prm.altitude);
    prm.rayleigh_scale = scalar("rayleigh", prm.rayleigh_scale);
    prm.mie_scale = scalar("mie", prm.mie_scale);
    prm.ozone_scale = scalar("ozone", prm.ozone_scale);
    const float radiance_scale = scale * ((2.0f * kPiF) * (1.0f - cosf(0.5f * angular_size)));
    const Spd sun_spectrum = spd_black_body(t(), 5900.0f, radiance_scale);
    uint32_t sky_w = std::max(64u, uint32_t(2048u * quality)), sky_h = std::max(64u, uint32_t(1024u * quality));
    float pr[441], pm[441], po[441];
    for (int i = 0; i < 441; ++i) pr[i] = spectra[rayleigh].entries[i].power, pm[i] = spectra[mie].entries[i].power, po[i] = spectra[ozone].entries[i].power;
    const atmosphere::Medium medium = {pr, pm, po};
    {
      Spd s = sun_spectrum;
      spd_scale(s, sun_scale);
      uint32_t image = kInvalid;
      if (angular_size > 0.0f) {
        Pixels px;
        px.w = px.h = 128;
        px.f32.assign(size_t(128) * 128 * 4, 0.0f);
        atmosphere::sun_image(t(), medium, prm, direction, angular_size, 128, 128, px.f32.data());
        image = add_image_data(std::move(px), 0u);
      }
      add_distant(2u, s, image, direction, angular_size);
    }
    {
      Spd s = sun_spectrum;
      spd_scale(s, sky_scale);
      Pixels px;
      px.w = sky_w, px.h = sky_h;
      px.f32.assign(size_t(sky_w) * sky_h * 4, 0.0f);
      auto table = atmosphere::optical_length_table(data_folder);
      atmosphere::sky_image(t(), medium, *table, prm, direction, sky_w, sky_h, px.f32.data());
      uint32_t image = add_image_data(std::move(px), IMG_BUILD_TABLE);
      add_distant(1u, s, image, direction, 0.0f);
    }
  }

  // load_from_file :805-820: a scene file that declares no distant emitter gets this atmosphere
  void add_default_atmosphere() {
    MtlBlock b;
    b.name = "et::atmosphere";
    b.params.push_back({"direction", "0.0 2.0 1.0"});
    b.params.push_back({"quality", "0.125"});
    b.params.push_back({"angular_diameter", "0.5422"});
    b.params.push_back({"anisotropy", "0.825"});
    b.params.push_back({"altitude", "1000.0"});
    b.params.push_back({"scale", "1.0"});
    b.params.push_back({"sky_scale", "1.0"});
    b.params.push_back({"sun_scale", "1.0"});
    b.params.push_back({"rayleigh", "1.0"});
    b.params.push_back({"mie", "1.0"});
    b.params.push_back({"ozone", "1.0"});
    parse_atmosphere(b);
  }

  void parse_spectrum(const MtlBlock& b) {  // (:1496-1611)
    std::string v;
    if (!b.get("id", v)) {
      warn("spectrum does not have identifier - skipped");
      return;
    }
    const std::string name = trim(v);
    float scale = b.get("scale", v) ? c_atof(v) : 1.0f;
    const bool illuminant = b.has("illuminant");
    bool initialized = false;
    Spd spd = {};
    if (b.get("rgb", v)) {
      auto p = split_params(v);
      if (p.size() < 3) return;
      F3 value = gamma_to_linear({c_atof(p[0]), c_atof(p[1]), c_atof(p[2])});
      spd = illuminant ? spd_rgb_luminance(t(), value) : spd_rgb_reflectance(t(), value);
      initialized = true;
    } else if (b.get("blackbody", v)) {
      auto p = split_params(v);
      if (p.empty()) return;
      spd = spd_black_body(t(), c_atof(p[0]), scale);
      initialized = true;
    } else if (b.get("nblackbody", v)) {
      auto p = split_params(v);
      if (p.empty()) return;
      float s2 = 1.0f;
      for (size_t i = 0; i < p.size(); ++i)
        if (i + 1 < p.size() && p[i] == "scale") s2 = c_atof(p[++i]);
      spd = spd_normalized_black_body(t(), c_atof(p[0]), s2);
      initialized = true;
    }
    const bool have_samples = b.get("samples", v);
    if (!have_samples && !initialized) return;
    if (!initialized) {
      auto p = split_params(v);
      if (p.size() % 2) return;
      std::vector<std::pair<float, float>> samples;
      for (size_t i = 0; i + 1 < p.size(); i += 2) samples.push_back({c_atof(p[i]), c_atof(p[i + 1])});
      spd = spd_from_samples(t(), samples);
      std::string mode;
      if (b.get("normalize", mode)) {
        F3 xyz = integrate_to_xyz(t(), spd);
        F3 rgb = xyz_to_rgb(xyz);
        float lum = (trim(mode) != "luminance") ? fmaxf(fmaxf(0.0f, rgb.x), fmaxf(rgb.y, rgb.z)) : xyz.y;
        if (lum > kEps) spd_scale(spd, 1.0f / lum);
      }
    }
    spd_scale(spd, scale);
    named_spectra[name] = add_spectrum(spd);
  }

  // ---- materials (:1682-2079) ---------------------------------------------------------------------------------------------------------------------------------
  static uint32_t material_class(const std::string& s) {
    static const char* names[] = {"diffuse", "translucent", "plastic", "conductor", "dielectric", "thinfilm", "mirror", "boundary", "velvet", "principled", "void"};
    for (uint32_t i = 0; i < 11; ++i)
      if (s == names[i]) return i;
    return ETXB_MAT_DIFFUSE;
  }

  void parse_material(const MtlBlock& b) {
    uint32_t mi = material_index.count(b.name) ? material_index[b.name] : add_material(b.name);
    std::string v, f;
    {
      etxb_material& m = materials[mi];
      m.cls = ETXB_MAT_DIFFUSE;
      m.emission = {kInvalid, kInvalid};
      m.emission_collimation = 0.0f;
    }
    if (b.get("base", v) && material_index.count(trim(v))) materials[mi] = materials[material_index[trim(v)]];
    if (b.get("Kd", v)) materials[mi].scattering.spectrum_index = reflectance_spectrum(v);
    if (b.get("Ks", v)) materials[mi].reflectance.spectrum_index = reflectance_spectrum(v);
    if (b.get("Kt", v)) materials[mi].scattering.spectrum_index = reflectance_spectrum(v);
    if (b.get("two_sided", v)) {  // an integer, else the whole value against "true" / "on" (:1714-1722)
      char* end = nullptr;
      long val = strtol(v.c_str(), &end, 10);
      materials[mi].two_sided = (end != v.c_str()) ? (val != 0 ? 1u : 0u) : ((v == "true" || v == "on") ? 1u : 0u);
    }
    if (b.get("opacity", v)) {
      auto fl = leading_floats(v, 1);
      if (!fl.empty()) materials[mi].opacity = clampf(fl[0], 0.0f, 1.0f);
    }
    if (b.get("Pr", v)) {
      auto fl = leading_floats(v, 2);
      float* r = materials[mi].roughness.value;
      if (fl.size() == 2) {
        r[0] = fl[0] * fl[0], r[1] = fl[1] * fl[1], r[2] = 0.0f, r[3] = 0.0f;
      } else if (fl.size() == 1) {
        r[0] = r[1] = fl[0] * fl[0], r[2] = 0.0f, r[3] = 0.0f;
      }
    }
    if (b.get("metalness", v)) {
      auto fl = leading_floats(v, 1);
      if (!fl.empty())
        for (float& x : materials[mi].metalness.value) x = fl[0];
    }
    if (b.get("transmission", v)) {
      auto fl = leading_floats(v, 1);
      if (!fl.empty())
        for (float& x : materials[mi].transmission.value) x = fl[0];
    }
    // map_Ml / map_Tm: metalness / transmission maps with an optional `channel N` (:1772-1800).  (`map_Pr`, the roughness map of the same block, is
    // dead code in the reference: the .mtl reader consumes that key as a standard texture before parse_material looks for it.)
    auto channel_map = [&](const char* key, uint32_t& image_index, uint32_t& channel) {
      if (!b.get(key, v)) return;
      auto p = split_params(v);
      int ch = 0;
      for (size_t i = 0; i < p.size(); ++i) {
        if (p[i] == "channel" && i + 1 < p.size()) {
          ch = std::max(0, atoi(p[i + 1].c_str()));
          ++i;
        }
      }
      if (find_file(p[0], f)) {
        image_index = add_image_file(f, IMG_REPEAT_U | IMG_REPEAT_V);
        channel = uint32_t(ch);
      }
    };
    channel_map("map_Ml", materials[mi].metalness.image_index, materials[mi].metalness.channel);
    channel_map("map_Tm", materials[mi].transmission.image_index, materials[mi].transmission.channel);
    auto texture = [&](const char* slot) -> std::string {
      auto it = b.textures.find(slot);
      return it == b.textures.end() ? std::string("") : it->second;
    };
    if (find_file(texture("diffuse"), f)) materials[mi].scattering.image_index = add_image_file(f, IMG_REPEAT_U | IMG_REPEAT_V);
    if (find_file(texture("specular"), f)) materials[mi].reflectance.image_index = add_image_file(f, IMG_REPEAT_U | IMG_REPEAT_V);
    if (find_file(texture("transmittance"), f)) materials[mi].scattering.image_index = add_image_file(f, IMG_REPEAT_U | IMG_REPEAT_V);
    if (b.get("material", v)) {
      auto p = split_params(v);
      for (size_t i = 0; i < p.size(); ++i)
        if (p[i] == "class" && i + 1 < p.size()) materials[mi].cls = material_class(p[++i]);
    }
    if (b.get("diffuse", v)) {
      uint32_t var = 0;
      if (leading_uint(v, var)) materials[mi].diffuse_variation = var;
    }
    if (b.get("int_ior", v)) {
      etxb_refractive_index r = materials[mi].int_ior;
      load_ior(r, v);
      materials[mi].int_ior = r;
    } else {
      uint32_t eta = add_spectrum(spd_constant(1.5f)), k = add_spectrum(spd_constant(0.0f));
      materials[mi].int_ior = {SPD_DIELECTRIC, eta, k};
    }
    if (b.get("ext_ior", v)) {
      etxb_refractive_index r = materials[mi].ext_ior;
      load_ior(r, v);
      materials[mi].ext_ior = r;
    } else {
      uint32_t eta = add_spectrum(spd_constant(1.0f)), k = add_spectrum(spd_constant(0.0f));
      materials[mi].ext_ior = {SPD_DIELECTRIC, eta, k};
    }
    if (b.get("int_medium", v)) {
      if (!medium_names.count(trim(v))) warn("medium " + trim(v) + " was not declared, but used in material " + b.name);
      materials[mi].int_medium = medium_names.count(trim(v)) ? medium_names[trim(v)] : kInvalid;
    }
    if (b.get("ext_medium", v)) {
      if (!medium_names.count(trim(v))) warn("medium " + trim(v) + " was not declared, but used in material " + b.name);
      materials[mi].ext_medium = medium_names.count(trim(v)) ? medium_names[trim(v)] : kInvalid;
    }
    if (b.get("normalmap", v)) {
      auto p = split_params(v);
      for (size_t i = 0; i < p.size(); ++i) {
        if (p[i] == "image" && i + 1 < p.size()) {
          if (find_file(p[i + 1], f)) materials[mi].normal_image_index = add_image_file(f, IMG_REPEAT_U | IMG_REPEAT_V | IMG_SKIP_SRGB);
          i += 1;
        }
        if (i < p.size() && p[i] == "scale" && i + 1 < p.size()) {
          materials[mi].normal_scale = c_atof(p[i + 1]);
          i += 1;
        }
      }
    }
    if (b.get("thinfilm", v)) {
      auto p = split_params(v);
      for (size_t i = 0; i < p.size(); ++i) {
        if (p[i] == "image" && i + 1 < p.size()) {
          if (find_file(p[i + 1], f)) materials[mi].thinfilm.thickness_image = add_image_file(f, IMG_REPEAT_U | IMG_REPEAT_V);
          i += 1;
        }
        if (i < p.size() && p[i] == "range" && i + 2 < p.size()) {
          materials[mi].thinfilm.min_thickness = c_atof(p[i + 1]);
          materials[mi].thinfilm.max_thickness = c_atof(p[i + 2]);
          i += 2;
        }
        if (i < p.size() && p[i] == "ior" && i + 1 < p.size()) {
          auto fl = leading_floats(p[i + 1], 1);
          if (!fl.empty()) {
            uint32_t eta = add_spectrum(spd_constant(fl[0]));
            materials[mi].thinfilm.ior = {SPD_DIELECTRIC, eta, kInvalid};
          } else {
            etxb_refractive_index r = materials[mi].thinfilm.ior;
            load_ior(r, p[i + 1]);
            materials[mi].thinfilm.ior = r;
          }
        }
      }
    }
    if (b.get("subsurface", v)) {
      materials[mi].subsurface.cls = 1u;
      float scale = 1.0f;
      F3 dist = {1.0f, 0.2f, 0.04f};
      auto p = split_params(v);
      for (size_t i = 0; i < p.size(); ++i) {
        if (p[i] == "path" && i + 1 < p.size()) materials[mi].subsurface.path = (p[i + 1] == "refracted" || p[i + 1] == "refraction" || p[i + 1] == "refract") ? 1u : 0u;
        if (p[i] == "distances" && i + 3 < p.size()) {
          dist = {c_atof(p[i + 1]), c_atof(p[i + 2]), c_atof(p[i + 3])};
          i += 3;
        }
        if (i < p.size() && p[i] == "scale" && i + 1 < p.size()) {
          scale = c_atof(p[i + 1]);
          i += 1;
        }
        if (i < p.size() && p[i] == "class" && i + 1 < p.size()) {
          if (p[i + 1] == "approximate") materials[mi].subsurface.cls = 2u;
          i += 1;
        }
      }
      Spd s = spd_rgb_reflectance(t(), dist);
      spd_scale(s, scale);
      materials[mi].subsurface.spectrum_index = add_spectrum(s);
    }
    // emission (:2009-2078)
    Spd spd = spd_constant(0.0f);
    bool defined = false, is_emitter = false;
    float pending = 1.0f, collimation = materials[mi].emission_collimation;
    if (b.get("Ke", v)) {
      is_emitter = true;
      spd = illuminant_spectrum(v);
      defined = true;
      if (find_file(texture("emissive"), f)) materials[mi].emission.image_index = add_image_file(f, IMG_REPEAT_U | IMG_REPEAT_V | IMG_BUILD_TABLE);
    }
    if (b.get("emitter", v)) {
      is_emitter = true;
      auto p = split_params(v);
      for (size_t i = 0; i < p.size(); ++i) {
        if (p[i] == "image" && i + 1 < p.size() && find_file(p[i + 1], f)) {
          materials[mi].emission.image_index = add_image_file(f, IMG_REPEAT_U | IMG_REPEAT_V | IMG_BUILD_TABLE);
          // the reference resolves the file name INTO the buffer its parameter pointers refer to (get_file -> data_buffer, :2029): whatever follows
          // `image <file>` on the line is lost there.  Same here, so that a scene file means the same thing in both.
          break;
        } else if (p[i] == "twosided") {
          materials[mi].two_sided = 1u;
        } else if (p[i] == "collimated" && i + 1 < p.size()) {
          collimation = c_atof(p[++i]);
        } else if (p[i] == "color" && i + 3 < p.size()) {
          spd = spd_rgb_luminance(t(), {c_atof(p[i + 1]), c_atof(p[i + 2]), c_atof(p[i + 3])});
          defined = true;
          i += 3;
        } else if (p[i] == "blackbody" && i + 1 < p.size()) {
          spd = spd_black_body(t(), c_atof(p[++i]), 1.0f);
          defined = true;
        } else if (p[i] == "nblackbody" && i + 1 < p.size()) {
          spd = spd_normalized_black_body(t(), c_atof(p[++i]), 1.0f);
          defined = true;
        } else if (p[i] == "scale" && i + 1 < p.size()) {
          pending *= c_atof(p[++i]);
        }
      }
      collimation = saturate(collimation);
    }
    etxb_material& m = materials[mi];
    if (is_emitter) {
      spd_scale(spd, pending);
      m.emission_collimation = collimation;
      if (defined && (luminance(load3(spd.integrated)) > 0.0f)) {
        m.emission.spectrum_index = add_spectrum(spd);
      } else if (!defined && m.emission.spectrum_index != kInvalid) {
        // keeps the spectrum inherited from `base`
      } else {
        m.emission.spectrum_index = kInvalid;
      }
      if (materials[mi].emission.spectrum_index == kInvalid) materials[mi].emission.image_index = kInvalid;
    } else if (m.emission.spectrum_index == kInvalid) {
      m.emission.image_index = kInvalid;
      m.emission_collimation = 0.0f;
    }
  }

  void parse_materials(const std::string& mtl_path) {
    base_dir = folder_of(mtl_path);
    for (const MtlBlock& b : parse_mtl(mtl_path)) {
      if (b.name == "et::camera") {
        parse_camera(b);
      } else if (b.name == "et::medium") {
        parse_medium(b);
      } else if (b.name == "et::dir") {
        parse_directional(b);
      } else if (b.name == "et::env") {
        parse_env(b);
      } else if (b.name == "et::atmosphere") {
        parse_atmosphere(b);
      } else if (b.name == "et::spectrum") {
        parse_spectrum(b);
      } else {
        parse_material(b);
      }
    }
  }

  // ---- geometry (:964-1052) -------------------------------------------------------------------------------------------------------------------------------------
  void load_obj(const std::string& obj_path, std::string mtl_path) {
    ObjData o = parse_obj(obj_path);
    if (mtl_path.empty()) {
      if (o.mtllib.empty()) fail(obj_path + ": no material file");
      mtl_path = join(folder_of(obj_path), o.mtllib);
    }
    parse_materials(mtl_path);
    const size_t nf = o.face_material.size();
    std::vector<int> mat_of_name(o.material_names.size(), -1);
    for (size_t i = 0; i < o.material_names.size(); ++i)
      if (material_index.count(o.material_names[i])) mat_of_name[i] = int(material_index[o.material_names[i]]);
    // load_from_obj skips a face whose material is unknown WITHOUT advancing its index cursor (:1005-1008, 1029): inside that shape the j-th face
    // that is kept reads the corners of the shape's j-th face.  Kept as the reference does it.
    std::vector<etxb_vertex>& v = vertices;
    std::vector<etxb_triangle> tris;
    std::vector<uint32_t> shapes;
    v.reserve(nf * 3);
    size_t cursor = 0;
    uint32_t current_shape = kInvalid;
    for (size_t fi = 0; fi < nf; ++fi) {
      if (o.face_shape[fi] != current_shape) {
        current_shape = o.face_shape[fi];
        cursor = fi;
      }
      const int name = o.face_material[fi];
      const int mat = (name >= 0) ? mat_of_name[size_t(name)] : -1;
      if (mat < 0) continue;
      const size_t src = cursor++;
      etxb_triangle tri = {};
      tri.material_index = uint32_t(mat);
      for (uint32_t k = 0; k < 3; ++k) {
        const ObjIndex& ix = o.corners[src * 3 + k];
        etxb_vertex vert = {};
        memcpy(vert.pos, &o.pos[size_t(ix.v) * 3], 12);
        if (ix.n >= 0 && size_t(ix.n) * 3 + 2 < o.nrm.size()) memcpy(vert.nrm, &o.nrm[size_t(ix.n) * 3], 12);
        if (ix.t >= 0 && size_t(ix.t) * 2 + 1 < o.tex.size()) memcpy(vert.tex, &o.tex[size_t(ix.t) * 2], 8);
        tri.i[k] = uint32_t(v.size());
        v.push_back(vert);
      }
      tris.push_back(tri);
      shapes.push_back(current_shape);
    }
    // medium bounds (:1036-1048): the running bounding box of the shape at the last triangle that carries the medium
    {
      F3 lo = {}, hi = {};
      uint32_t sh = kInvalid;
      for (size_t i = 0; i < tris.size(); ++i) {
        if (shapes[i] != sh) {
          sh = shapes[i];
          lo = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
          hi = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
        }
        for (uint32_t k = 0; k < 3; ++k) {
          F3 p = load3(v[tris[i].i[k]].pos);
          lo = {fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z)};
          hi = {fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z)};
        }
        uint32_t med = materials[tris[i].material_index].int_medium;
        if (med != kInvalid && med < mediums.size()) {
          store3(mediums[med].bounds_min, lo);
          store3(mediums[med].bounds_max, hi);
        }
      }
    }
    // validate_triangle (:252-260): a degenerate triangle is dropped, its vertices stay
    for (etxb_triangle& tri : tris) {
      F3 p0 = load3(v[tri.i[0]].pos), p1 = load3(v[tri.i[1]].pos), p2 = load3(v[tri.i[2]].pos);
      F3 gn = cross(p1 - p0, p2 - p0);
      float l = length(gn);
      if (l == 0.0f) continue;
      store3(tri.geo_n, gn / l);
      triangles.push_back(tri);
    }
  }

  void finish_geometry(bool force_tangents) {
    std::vector<etxb_vertex>& v = vertices;
    std::vector<bool> referenced(v.size(), false);
    // validate_normals (:304-335)
    for (const etxb_triangle& tri : triangles) {
      F3 p0 = load3(v[tri.i[0]].pos), p1 = load3(v[tri.i[1]].pos), p2 = load3(v[tri.i[2]].pos);
      float area = 0.5f * length(cross(p1 - p0, p2 - p0));
      for (uint32_t k = 0; k < 3; ++k) {
        referenced[tri.i[k]] = true;
        if (valid_vector(load3(v[tri.i[k]].nrm))) continue;
        store3(v[tri.i[k]].nrm, normalize(load3(tri.geo_n) * area));  // every vertex belongs to one triangle: the first contribution is an assignment
      }
    }
    // build_tangents (:337-398): without texture coordinates nothing; with them the tangent-space generator (scene_loader_tangents.inl)
    float lo[2] = {3.402823466e+38f, 3.402823466e+38f}, hi[2] = {-3.402823466e+38f, -3.402823466e+38f};
    for (const etxb_vertex& x : v) {
      lo[0] = fminf(lo[0], x.tex[0]), lo[1] = fminf(lo[1], x.tex[1]);
      hi[0] = fmaxf(hi[0], x.tex[0]), hi[1] = fmaxf(hi[1], x.tex[1]);
    }
    if (!v.empty() && ((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) > kEps)) tangents::generate(v, triangles);
    // validate_tangents (:400-418): orthonormal_basis (math.hxx:737-746) where no frame exists, then orthogonalize (scene.hxx:133-139) everywhere
    for (size_t i = 0; i < v.size(); ++i) {
      etxb_vertex& x = v[i];
      if (!force_tangents && valid_vector(load3(x.tan)) && valid_vector(load3(x.btn))) continue;
      if (force_tangents || referenced[i]) {
        F3 n = load3(x.nrm);
        F3 a = normalize(((n.x != n.y) || (n.x != n.z)) ? F3{n.z - n.y, n.x - n.z, +n.y - n.x} : F3{n.z - n.y, n.x + n.z, -n.y - n.x});
        F3 bb = normalize(cross(n, a));
        store3(x.tan, a);
        store3(x.btn, bb);
      }
    }
    for (etxb_vertex& x : v) {
      F3 b0 = load3(x.btn);
      F3 n = normalize(load3(x.nrm));
      F3 tn = normalize(load3(x.tan) - n * dot(load3(x.tan), n));
      F3 bt = normalize(cross(n, tn));
      bt = bt * (dot(b0, bt) > 0.0f ? 1.0f : -1.0f);
      store3(x.nrm, n), store3(x.tan, tn), store3(x.btn, bt);
    }
  }

  // build_camera (:579-598)
  void build_camera(F3 origin, F3 target, F3 up, uint32_t width, uint32_t height, float fov_deg) {
    etxb_camera& cam = camera;
    F3 f = normalize(target - origin), s = normalize(cross(f, up)), u = cross(s, f);
    float view[4][4] = {};  // view[col][row]
    view[0][0] = s.x, view[1][0] = s.y, view[2][0] = s.z;
    view[0][1] = u.x, view[1][1] = u.y, view[2][1] = u.z;
    view[0][2] = -f.x, view[1][2] = -f.y, view[2][2] = -f.z;
    view[3][0] = -dot(s, origin), view[3][1] = -dot(u, origin), view[3][2] = dot(f, origin), view[3][3] = 1.0f;
    float fov = fov_deg * kPiF / 180.0f;
    float w = cosf(0.5f * fov) / sinf(0.5f * fov);
    float aspect = float(width) / float(height);
    float proj[4][4] = {};
    proj[0][0] = w;
    proj[1][1] = w * aspect;
    proj[2][2] = cam.clip_far / (cam.clip_near - cam.clip_far);
    proj[2][3] = -1.0f;
    proj[3][2] = -(cam.clip_far * cam.clip_near) / (cam.clip_far - cam.clip_near);
    for (int j = 0; j < 4; ++j) {
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int k = 0; k < 4; ++k)
        for (int r = 0; r < 4; ++r) acc[r] = acc[r] + proj[k][r] * view[j][k];
      memcpy(cam.view_proj + j * 4, acc, 16);
    }
    store3(cam.target, target);
    store3(cam.position, origin);  // inverse(view).col[3]: the origin up to a rounding step
    store3(cam.side, s);
    store3(cam.up, u);
    store3(cam.direction, f);
    cam.tan_half_fov = 1.0f / fabsf(w);
    cam.aspect = proj[1][1] / proj[0][0];
    float plane_w = 2.0f * cam.tan_half_fov, plane_h = 2.0f * cam.tan_half_fov / cam.aspect;
    cam.area = plane_w * plane_h;
    cam.film_size[0] = width, cam.film_size[1] = height;
    cam.image_plane = float(width) / (2.0f * cam.tan_half_fov);
  }

  // Film::generate_filter_image(PixelFilterBlackmanHarris) (film.cxx:63-67, 123-135)
  uint32_t add_pixel_filter() {
    Pixels px;
    px.w = px.h = 128;
    px.f32.resize(128 * 128 * 4);
    for (uint32_t y = 0; y < 128; ++y) {
      for (uint32_t x = 0; x < 128; ++x) {
        float dx = float(x) - 64.0f, dy = float(y) - 64.0f;
        float dist = sqrtf(dx * dx + dy * dy);
        float r = (2.0f * kPiF) * saturate(0.5f + dist / (2.0f * 64.0f));
        float value = 0.35875f - 0.48829f * cosf(r) + 0.14128f * cosf(2.0f * r) - 0.01168f * cosf(3.0f * r);
        float* o = px.f32.data() + (size_t(x) + size_t(y) * 128) * 4;
        o[0] = o[1] = o[2] = value, o[3] = 1.0f;
      }
    }
    return add_image_data(std::move(px), IMG_BUILD_TABLE | IMG_UNIFORM_TABLE);
  }

  // add_area_emitters_for_triangle (:840-905), incl. the emission-image factor
  void add_area_emitters() {
    tri_to_emitter.assign(triangles.size(), kInvalid);
    std::map<uint32_t, uint32_t> profile_of_material;
    for (uint32_t ti = 0; ti < triangles.size(); ++ti) {
      const etxb_triangle& tri = triangles[ti];
      const etxb_material& m = materials[tri.material_index];
      if (m.emission.spectrum_index == kInvalid || m.emission.spectrum_index >= spectra.size()) continue;
      float texture_emission = 1.0f;
      if (m.emission.image_index != kInvalid) {
        const ImageRecord& img = *images[m.emission.image_index];
        const float* t0 = vertices[tri.i[0]].tex;
        const float* t1 = vertices[tri.i[1]].tex;
        const float* t2 = vertices[tri.i[2]].tex;
        float min_u = fminf(t0[0], fminf(t1[0], t2[0])), min_v = fminf(t0[1], fminf(t1[1], t2[1]));
        float max_u = fmaxf(t0[0], fmaxf(t1[0], t2[0])), max_v = fmaxf(t0[1], fmaxf(t1[1], t2[1]));
        float u_size = 4.0f * fmaxf(1.0f, ceilf((max_u - min_u) * float(img.px.w)));
        float du = 1.0f / u_size;
        float v_size = 4.0f * fmaxf(1.0f, ceilf((max_v - min_v) * float(img.px.h)));
        float dv = 1.0f / v_size;
        for (float vv = 0.0f; vv < 1.0f; vv += dv) {
          for (float uu = 0.0f; uu < 1.0f; uu += dv) {  // the reference steps the inner loop by dv too
            float r1 = sqrtf(uu);
            float bc[3] = {1.0f - r1, r1 * (1.0f - vv), r1 * vv};  // random_barycentric (math.hxx:768-771)
            float tu = t0[0] * bc[0] + t1[0] * bc[1] + t2[0] * bc[2], tv = t0[1] * bc[0] + t1[1] * bc[1] + t2[1] * bc[2];
            float val[4];
            img.evaluate(tu, tv, val);
            texture_emission += luminance({val[0], val[1], val[2]}) * du * dv * val[3];
          }
        }
      }
      F3 p0 = load3(vertices[tri.i[0]].pos), p1 = load3(vertices[tri.i[1]].pos), p2 = load3(vertices[tri.i[2]].pos);
      float tri_area = 0.5f * length(cross(p1 - p0, p2 - p0));
      float spectrum_weight = luminance(load3(spectra[m.emission.spectrum_index].integrated));
      float additional_weight = (m.two_sided ? 2.0f : 1.0f) * (tri_area * kPiF) * texture_emission;
      if ((additional_weight <= 0.0f) || (spectrum_weight <= 0.0f)) continue;
      uint32_t profile;
      auto it = profile_of_material.find(tri.material_index);
      if (it != profile_of_material.end()) {
        profile = it->second;
      } else {
        profile = uint32_t(profiles.size());
        profile_of_material[tri.material_index] = profile;
        etxb_emitter_profile p = {};
        p.cls = 0u;
        p.angular_size_cosine = 1.0f;
        profiles.push_back(p);
      }
      profiles[profile].emission = m.emission;
      etxb_emitter e = {};
      e.cls = 0u;
      e.profile = profile;
      e.triangle_index = ti;
      e.triangle_area = tri_area;
      e.additional_weight = additional_weight;
      e.spectrum_weight = spectrum_weight;
      tri_to_emitter[ti] = uint32_t(emitters.size());
      emitters.push_back(e);
    }
  }

  void finalize(uint32_t samples, bool spectral, uint32_t max_len, uint32_t min_len, uint32_t rr_start) {
    // validate_materials (:262-302): every missing spectrum is a NEW entry of the pool
    for (etxb_material& m : materials) {
      if (m.reflectance.spectrum_index == kInvalid) m.reflectance.spectrum_index = add_spectrum(spd_rgb_reflectance(t(), {1.0f, 1.0f, 1.0f}));
      if (m.scattering.spectrum_index == kInvalid) m.scattering.spectrum_index = add_spectrum(spd_rgb_reflectance(t(), {1.0f, 1.0f, 1.0f}));
      if (m.subsurface.spectrum_index == kInvalid) m.subsurface.spectrum_index = add_spectrum(spd_rgb_reflectance(t(), {1.0f, 0.2f, 0.04f}));
      if (m.emission.spectrum_index == kInvalid) m.emission.spectrum_index = add_spectrum(spd_constant(0.0f));
      if ((m.roughness.value[0] > 0.0f) || (m.roughness.value[1] > 0.0f)) {
        m.roughness.value[0] = fmaxf(1e-6f, m.roughness.value[0]);
        m.roughness.value[1] = fmaxf(1e-6f, m.roughness.value[1]);
      }
      const bool conductor = m.cls == ETXB_MAT_CONDUCTOR;
      if (m.int_ior.eta_index == kInvalid) m.int_ior.eta_index = conductor ? def_cond_eta : def_diel;
      if (m.int_ior.k_index == kInvalid) m.int_ior.k_index = conductor ? def_cond_k : add_spectrum(spd_constant(0.0f));
      if (m.thinfilm.ior.k_index == kInvalid) m.thinfilm.ior.k_index = add_spectrum(spd_constant(0.0f));
      if (m.thinfilm.ior.eta_index == kInvalid) m.thinfilm.ior.eta_index = add_spectrum(spd_constant(1.0f));
    }
    // commit (:420-455)
    const uint32_t pixel_filter = add_pixel_filter();
    F3 lo = {-1.0f, -1.0f, -1.0f}, hi = {1.0f, 1.0f, 1.0f};
    if (!triangles.empty()) {
      lo = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
      hi = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
      for (const etxb_triangle& tri : triangles) {
        for (uint32_t k = 0; k < 3; ++k) {
          F3 p = load3(vertices[tri.i[k]].pos);
          lo = {fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z)};
          hi = {fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z)};
        }
      }
    }
    F3 center = (lo + hi) * 0.5f;
    const float radius = length(hi - center);
    add_area_emitters();
    // build_emitters_distribution (:2460-2497)
    for (uint32_t i = 0; i < profiles.size(); ++i) {
      etxb_emitter_profile& p = profiles[i];
      if (p.cls != 0u) {
        p.equivalent_disk_size = 2.0f * tanf(p.angular_size / 2.0f);
        p.angular_size_cosine = cosf(p.angular_size / 2.0f);
        for (etxb_emitter& e : emitters)
          if (e.profile == i) e.additional_weight = kPiF * radius * radius;
      }
    }
    emitter_dist.assign(emitters.size() + 1, {});
    uint32_t env_count = 0;
    etxb_scene& sc = scene;
    for (uint32_t i = 0; i < emitters.size(); ++i) {
      etxb_emitter& e = emitters[i];
      const etxb_emitter_profile& p = profiles[e.profile];
      e.spectrum_weight = (p.emission.spectrum_index != kInvalid) ? luminance(load3(spectra[p.emission.spectrum_index].integrated)) : 0.0f;
      float total = e.spectrum_weight * e.additional_weight;
      emitter_dist[i].value = total;
      if (e.cls == 0u) {
        tri_to_emitter[e.triangle_index] = i;
      } else if (total > 0.0f && env_count < 63u) {
        sc.environment_emitters[env_count++] = i;
      }
    }
    sc.emitters_distribution.total_weight = finalize_distribution(emitter_dist, uint32_t(emitters.size()));
    sc.emitters_distribution.values = {emitter_dist.data(), uint32_t(emitters.size())};
    sc.environment_emitter_count = env_count;
    // images -> PODs
    image_pods.assign(images.size(), {});
    for (size_t i = 0; i < images.size(); ++i) {
      ImageRecord& r = *images[i];
      etxb_image& pod = image_pods[i];
      pod.pixels = {r.px.eight_bit ? static_cast<const void*>(r.px.u8.data()) : static_cast<const void*>(r.px.f32.data()), uint64_t(r.px.w) * r.px.h};
      if (r.options & IMG_BUILD_TABLE) {
        pod.x_distributions = {r.x_dists.data(), r.px.h};
        pod.y_distribution.values = {r.y_entries.data(), r.px.h};
        pod.y_distribution.total_weight = r.y_total;
        pod.normalization = r.normalization;
      }
      pod.fsize[0] = float(r.px.w), pod.fsize[1] = float(r.px.h);
      pod.isize[0] = r.px.w, pod.isize[1] = r.px.h;
      memcpy(pod.offset, r.offset, 8);
      memcpy(pod.scale, r.scale, 8);
      pod.options = r.options;
      pod.format = r.px.eight_bit ? 2u : 1u;
      pod.data_size = uint32_t(size_t(r.px.w) * r.px.h * (r.px.eight_bit ? 4u : 16u));
    }
    sc.vertices = {vertices.data(), vertices.size()};
    sc.triangles = {triangles.data(), triangles.size()};
    sc.triangle_to_emitter = {tri_to_emitter.data(), tri_to_emitter.size()};
    sc.materials = {materials.data(), materials.size()};
    sc.emitter_profiles = {profiles.data(), profiles.size()};
    sc.emitter_instances = {emitters.data(), emitters.size()};
    sc.images = {image_pods.data(), image_pods.size()};
    sc.mediums = {mediums.data(), mediums.size()};
    sc.spectrums = {spectra.data(), spectra.size()};
    store3(sc.bounding_sphere_center, center);
    sc.bounding_sphere_radius = radius;
    sc.pixel_sampler_image = pixel_filter;
    sc.pixel_sampler_radius = 1.5f;
    sc.min_path_length = min_len, sc.max_path_length = max_len, sc.samples = samples, sc.random_path_termination = rr_start;
    sc.noise_threshold = 0.1f;
    sc.radiance_clamp = 0.0f;
    sc.black_spectrum = black, sc.white_spectrum = white, sc.rayleigh_spectrum = rayleigh, sc.mie_spectrum = mie, sc.ozone_spectrum = ozone;
    sc.subsurface_scatter_material = ss_scatter, sc.subsurface_exit_material = ss_exit;
    sc.default_dielectric_eta = def_diel, sc.default_conductor_eta = def_cond_eta, sc.default_conductor_k = def_cond_k;
    sc.flags = ETXB_SCENE_COMMITTED | (spectral ? ETXB_SCENE_SPECTRAL : 0u);
  }

  // SceneRepresentation::load_from_file (:679-838)
  void load(const std::string& file_name) {
    init_default_values();
    const std::string base = folder_of(file_name);
    std::string geometry = file_name, materials_file;
    uint32_t samples = 256, rr_start = 6, max_len = 65535, min_len = 0;  // Scene defaults (scene.hxx:41-44)
    bool spectral = false, force_tangents = false;
    CameraBlock cam;
    cam.origin = {5.0f, 5.0f, 5.0f};
    {
      float t5 = 5.0f + (-5.0f / sqrtf(75.0f));  // the default camera's position + its (normalised) direction (:694)
      cam.target = {t5, t5, t5};
    }
    cam.fov = 26.99f;
    bool has_focal = false;
    float focal = 0.0f;
    const size_t dot_at = file_name.find_last_of('.');
    const std::string ext = dot_at == std::string::npos ? std::string("") : lower(file_name.substr(dot_at));
    if (ext == ".json") {
      std::string text = read_file(file_name);
      JsonParser parser{text};
      Json js = parser.value();
      if (js.kind != Json::Object) fail(file_name + ": a JSON object was expected");
      // json_get_int / _float / _bool / _string (core/json.hxx:41-72) look at the value's type: a number where a bool is expected (or the other way round) is ignored
      auto number = [](const Json& j) { return j.number; };
      auto is_number = [](const Json& j) { return j.kind == Json::Number; };
      for (const auto& kv : js.members) {
        const std::string& key = kv.first;
        const Json& val = kv.second;
        if (key == "samples" && is_number(val)) samples = uint32_t(std::max<int64_t>(1, int64_t(number(val))));
        else if (key == "random-termination-start" && is_number(val)) rr_start = uint32_t(std::max<int64_t>(1, int64_t(number(val))));
        else if (key == "max-path-length" && is_number(val)) max_len = uint32_t(std::max<int64_t>(1, int64_t(number(val))));
        else if (key == "min-path-length" && is_number(val)) min_len = uint32_t(std::max<int64_t>(1, int64_t(number(val))));  // the reference clamps this one to 1 as well (:716)
        else if (key == "geometry" && val.kind == Json::String) geometry = join(base, val.text);
        else if (key == "materials" && val.kind == Json::String) materials_file = join(base, val.text);
        else if (key == "spectral" && val.kind == Json::Bool) spectral = val.b;
        else if (key == "force-tangents" && val.kind == Json::Bool) force_tangents = val.b;
        else if (key == "camera" && val.kind == Json::Object) {
          for (const auto& ck : val.members) {
            const Json& cv = ck.second;
            auto vec = [&](F3& out) {
              if (cv.kind == Json::Array && cv.items.size() >= 3) out = {float(cv.items[0].number), float(cv.items[1].number), float(cv.items[2].number)};
            };
            if (ck.first == "class") cam.cls = (cv.text == "eq") ? 1u : 0u;
            else if (ck.first == "fov" && is_number(cv)) cam.fov = float(number(cv));
            else if (ck.first == "focal-length" && is_number(cv)) focal = float(number(cv)), has_focal = true;
            else if (ck.first == "lens-radius" && is_number(cv)) cam.lens_radius = float(number(cv));
            else if (ck.first == "focal-distance" && is_number(cv)) cam.focal_distance = float(number(cv));
            else if (ck.first == "clip-near" && is_number(cv)) cam.clip_near = float(number(cv)), cam.has_near = true;
            else if (ck.first == "clip-far" && is_number(cv)) cam.clip_far = float(number(cv)), cam.has_far = true;
            else if (ck.first == "origin") vec(cam.origin);
            else if (ck.first == "target") vec(cam.target);
            else if (ck.first == "up") vec(cam.up);
            else if (ck.first == "viewport" && cv.kind == Json::Array && cv.items.size() >= 2)
              cam.viewport[0] = uint32_t(cv.items[0].number), cam.viewport[1] = uint32_t(cv.items[1].number);
          }
        }
      }
    }
    if (cam.viewport[0] * cam.viewport[1] == 0u) cam.viewport[0] = 1280, cam.viewport[1] = 720;
    {
      const size_t gd = geometry.find_last_of('.');
      if (gd == std::string::npos || lower(geometry.substr(gd)) != ".obj") fail(geometry + ": only Wavefront .obj geometry is read by this loader (glTF is not)");
    }
    load_obj(geometry, materials_file);
    if (profiles.empty()) add_default_atmosphere();
    // camera (:789-804); clip planes default to Camera's own (camera.hxx: 1 / 256 and 1024)
    const CameraBlock* sel = nullptr;
    if (!cameras.empty()) {
      sel = &cameras.front();
      for (const CameraBlock& c : cameras) {
        if (c.active) {
          sel = &c;
          break;
        }
      }
    }
    camera = {};
    camera.lens_image = kInvalid;
    camera.medium_index = kInvalid;
    camera.clip_near = 1.0f / 256.0f;
    camera.clip_far = 1024.0f;
    if (sel != nullptr) {
      F3 origin = sel->has_origin ? sel->origin : F3{0.0f, 0.0f, 0.0f};
      F3 target = sel->has_target ? sel->target : F3{origin.x, origin.y, origin.z - 1.0f};
      uint32_t w = sel->viewport[0], h = sel->viewport[1];
      if (w * h == 0u) w = 1280, h = 720;
      camera.lens_radius = sel->lens_radius, camera.focal_distance = sel->focal_distance;
      if (sel->has_near) camera.clip_near = sel->clip_near;
      if (sel->has_far) camera.clip_far = sel->clip_far;
      build_camera(origin, target, sel->up, w, h, sel->fov);
      camera.cls = sel->cls;
      camera.lens_image = sel->lens_image;
      camera.medium_index = sel->medium;
    } else {
      float fov = cam.fov;
      if (has_focal) fov = (2.0f * atanf(36.0f / (2.0f * focal))) * 180.0f / kPiF;
      camera.lens_radius = cam.lens_radius, camera.focal_distance = cam.focal_distance;
      if (cam.has_near) camera.clip_near = cam.clip_near;
      if (cam.has_far) camera.clip_far = cam.clip_far;
      build_camera(cam.origin, cam.target, cam.up, cam.viewport[0], cam.viewport[1], fov);
      camera.cls = cam.cls;
    }
    finish_geometry(force_tangents);
    finalize(samples, spectral, max_len, min_len, rr_start);
  }
};
