// scene_loader_atmosphere.inl — the procedural sun disk and sky dome (included by scene_loader.cpp inside its anonymous namespace).
//
// The reference gives EVERY scene file without an et::dir / et::env / et::atmosphere block a default atmosphere (scene_representation.cxx:805-820),
// and an explicit `newmtl et::atmosphere` block the same thing with the file's parameters (parse_atmosphere_light :1376-1495): a Directional emitter
// with a 128 x 128 limb-darkened extinction image and an Environment emitter with a single-scattering sky image (render/host/scattering.cxx:24-384).
// Restated here: the same physical model and fitted constants (Rayleigh / Mie / ozone, the density profile, the adaptive march), float arithmetic in
// the reference's order so the images agree to rounding; the work is spread over the host threads in fixed chunks and reduced in chunk order (the
// reference adds its per-task partial sums into atomics, so its own result varies in the last bits from run to run).
//
// The optical-length table (1024 x 1024, a function of the constants only; ~30 s on 8 cores) is kept for the life of the process and, when the data
// folder is writable, beside tables.bin as atmosphere_optical_length.bin.

struct AtmosphereParameters {  // scattering::Parameters (render/shared/scattering.hxx:11-17)
  float altitude = 1000.0f, anisotropy = 0.825f, rayleigh_scale = 1.0f, mie_scale = 1.0f, ozone_scale = 1.0f;
};

namespace atmosphere {

constexpr float kPlanetRadius = 6371e+3f, kShellThickness = 120e+3f, kOuterRadius = kPlanetRadius + kShellThickness;
constexpr float kDensityStep = 0.01f, kRayleighScaleHeight = 7994.0f, kMieScaleHeight = 1200.0f;
constexpr uint32_t kTableSize = 1024u;

inline F3 mul(F3 a, F3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }

template <class Fn>
void parallel_chunks(uint32_t count, uint32_t chunks, Fn fn) {  // fn(chunk, begin, end)
  chunks = std::max(1u, std::min(chunks, count));
  uint32_t workers = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
  std::atomic<uint32_t> next{0};
  auto body = [&]() {
    for (uint32_t c = next.fetch_add(1); c < chunks; c = next.fetch_add(1)) {
      uint32_t begin = uint32_t(uint64_t(count) * c / chunks), end = uint32_t(uint64_t(count) * (c + 1) / chunks);
      fn(c, begin, end);
    }
  };
  std::vector<std::thread> pool;
  for (uint32_t i = 1; i < workers; ++i) pool.emplace_back(body);
  body();
  for (auto& t : pool) t.join();
}

// relative densities of the three constituents at a height above the surface (scattering.cxx:52-68): two exponentials and a fitted ozone profile
F3 density(float height) {
  float h = fmaxf(0.0f, height);
  float x = h / 1000.0f, x2 = x * x, x3 = x2 * x, x4 = x2 * x2, x5 = x4 * x, x6 = x3 * x3;
  float f = 3.759384e-08f * x6 - 1.067250e-05f * x5 + 1.080311e-03f * x4 - 4.851181e-02f * x3 + 9.185432e-01f * x2 - 4.886021e+00f * x + 7.900478e+00f;
  constexpr float kOzoneScale = 1.0f / 30.8491249f;
  return {expf(-h / kRayleighScaleHeight), expf(-h / kMieScaleHeight), fmaxf(0.0f, f * kOzoneScale)};
}

// the march's step: short where the density changes fast (scattering.cxx:98-104)
float step_size(float travelled, float total, F3 origin, F3 direction, F3 d0) {
  F3 grad = density(length(origin + direction * (1.0f + travelled)) - kPlanetRadius) - d0;
  float l0 = logf((1.0f + grad.x) / kDensityStep) * kRayleighScaleHeight;
  float l1 = logf((1.0f + grad.y) / kDensityStep) * kMieScaleHeight;
  float calculated = sqrtf(kDensityStep * (l0 * l0 + l1 * l1));
  return fminf(total - travelled, calculated);
}

float distance_to_sphere(F3 origin, F3 direction, float radius) {  // math.hxx:1023-1034, sphere at the world origin
  float b = dot(direction, origin);
  float d = (b * b) - dot(origin, origin) + (radius * radius);
  if (d < 0.0f) return 0.0f;
  d = sqrtf(d);
  float a0 = -b - d, a1 = -b + d;
  return (a0 < 0.0f) ? ((a1 < 0.0f) ? 0.0f : a1) : a0;
}

F3 optical_length(F3 origin, F3 direction, float total) {  // scattering.cxx:106-124
  F3 result = {};
  F3 d = density(length(origin) - kPlanetRadius);
  float t = 0.0f;
  while (t < total) {
    float dt = step_size(t, total, origin, direction, d);
    F3 p = origin + direction * (t + 0.5f * dt);
    t += dt;
    d = density(length(p) - kPlanetRadius);
    result = result + d * dt;
  }
  return result;
}

// the table: u = ((n . l) / 2 + 1/2)^2, v = sqrt(height / shell) (scattering.cxx:70-80); RGBA32F, clamped bilinear lookups
std::shared_ptr<ImageRecord> optical_length_table(const std::string& data_folder) {
  static std::mutex guard;
  static std::shared_ptr<ImageRecord> table;
  std::lock_guard<std::mutex> lock(guard);
  if (table) return table;
  auto rec = std::make_shared<ImageRecord>();
  rec->px.w = rec->px.h = kTableSize;
  rec->px.f32.assign(size_t(kTableSize) * kTableSize * 4, 0.0f);
  const std::string cache = data_folder + "/atmosphere_optical_length.bin";
  const size_t bytes = rec->px.f32.size() * sizeof(float);
  const uint32_t tag[4] = {0x4c54504fu /* OPTL */, kTableSize, kTableSize, 1u};
  if (FILE* f = fopen(cache.c_str(), "rb")) {
    uint32_t head[4] = {};
    bool ok = fread(head, 4, 4, f) == 4 && memcmp(head, tag, sizeof(tag)) == 0 && fread(rec->px.f32.data(), 1, bytes, f) == bytes;
    fclose(f);
    if (ok) {
      table = rec;
      return table;
    }
  }
  float* image = rec->px.f32.data();
  parallel_chunks(kTableSize * kTableSize, 1024u, [image](uint32_t, uint32_t begin, uint32_t end) {
    for (uint32_t i = begin; i < end; ++i) {
      uint32_t x = i % kTableSize, y = i / kTableSize;
      float u = float(x) / float(kTableSize), v = float(y) / float(kTableSize);
      float height = (v * v) * kShellThickness, n_dot_l = sqrtf(u) * 2.0f - 1.0f;
      F3 direction = {sqrtf(1.0f - n_dot_l * n_dot_l), n_dot_l, 0.0f};
      F3 origin = {0.0f, kPlanetRadius + height, 0.0f};
      F3 value = optical_length(origin, direction, distance_to_sphere(origin, direction, kOuterRadius));
      image[size_t(i) * 4 + 0] = value.x, image[size_t(i) * 4 + 1] = value.y, image[size_t(i) * 4 + 2] = value.z, image[size_t(i) * 4 + 3] = 0.0f;
    }
  });
  const std::string tmp = cache + ".tmp" + std::to_string(uint64_t(getpid()));
  if (FILE* f = fopen(tmp.c_str(), "wb")) {  // best effort: a read-only installation recomputes the table once per process
    bool ok = fwrite(tag, 4, 4, f) == 4 && fwrite(image, 1, bytes, f) == bytes;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), cache.c_str()) != 0) remove(tmp.c_str());
  }
  table = rec;
  return table;
}

F3 sample_optical_length(const ImageRecord& table, F3 pos, F3 light_direction) {  // scattering.cxx:89-96
  float height = length(pos);
  float n_dot_l = dot(pos / height, light_direction);
  float half = n_dot_l * 0.5f + 0.5f;
  float u = half * half, v = sqrtf(saturate((height - kPlanetRadius) / kShellThickness));
  float e[4];
  table.evaluate(u, v, e);
  return {e[0], e[1], e[2]};
}

struct Medium {
  const float *rayleigh, *mie, *ozone;  // 441 coefficients each, 390..830 nm (the `power` column of scene spectra 2, 3, 4)
};

// single scattering along a view ray (scattering.cxx:126-181) -> 441 spectral radiances
void sky_radiance(const Medium& m, const ImageRecord& table, F3 view, F3 light, const AtmosphereParameters& prm, float* out441) {
  const F3 origin = {0.0f, kPlanetRadius + prm.altitude, 0.0f};
  const float l_dot_v = dot(light, view), g = prm.anisotropy;
  const float phase_r = (3.0f / 4.0f) * (1.0f + l_dot_v * l_dot_v) * (1.0f / (2.0f * kPiF));
  const float phase_m = (3.0f / 2.0f) * ((1.0f - g * g) * (1.0f + l_dot_v * l_dot_v)) / ((2.0f + g * g) * powf(1.0f + g * g - 2.0f * g * l_dot_v, 1.5f)) * (1.0f / (2.0f * kPiF));
  const F3 scale = {prm.rayleigh_scale, prm.mie_scale, prm.ozone_scale};
  F3 view_path = {};
  F3 current = density(length(origin) - kPlanetRadius);
  for (int i = 0; i < 441; ++i) out441[i] = 0.0f;
  float t = 0.0f;
  float to_space = distance_to_sphere(origin, view, kOuterRadius), to_planet = distance_to_sphere(origin, view, kPlanetRadius);
  if (to_planet > 0.0f) to_space = to_planet;
  while (t < to_space) {
    float dt = step_size(t, to_space, origin, view, current);
    F3 p = origin + view * (t + 0.5f * dt);
    float height = length(p) - kPlanetRadius;
    t += dt;
    if (height < -kRayleighScaleHeight) break;
    current = density(height);
    view_path = view_path + mul(scale * dt, current);
    F3 total = view_path + mul(scale, sample_optical_length(table, p, light));
    const float in_r = phase_r, in_m = phase_m;
    for (int i = 0; i < 441; ++i) {
      float r = m.rayleigh[i], mi = m.mie[i], o = m.ozone[i];
      float tr = r * total.x + mi * total.y + o * total.z;
      out441[i] += expf(-tr) * dt * (in_r * r * scale.x * current.x + in_m * mi * scale.y * current.y);
    }
  }
}

// transmittance towards the sun disk (scattering.cxx:183-222); black where the next row's direction already meets the planet
void sun_extinction(const Medium& m, F3 view, F3 next, const AtmosphereParameters& prm, float* out441) {
  const F3 origin = {0.0f, kPlanetRadius + prm.altitude, 0.0f};
  float to_space = distance_to_sphere(origin, view, kOuterRadius);
  if (distance_to_sphere(origin, next, kPlanetRadius) > 0.0f) {
    for (int i = 0; i < 441; ++i) out441[i] = 0.0f;
    return;
  }
  const F3 scale = {prm.rayleigh_scale, prm.mie_scale, prm.ozone_scale};
  F3 path = {};
  F3 current = density(length(origin) - kPlanetRadius);
  float t = 0.0f;
  while (t < to_space) {
    float dt = step_size(t, to_space, origin, view, current);
    F3 p = origin + view * (t + 0.5f * dt);
    t += dt;
    current = density(length(p) - kPlanetRadius);
    path = path + mul(scale * dt, current);
  }
  for (int i = 0; i < 441; ++i) out441[i] = expf(-(m.rayleigh[i] * path.x + m.mie[i] * path.y + m.ozone[i] * path.z));
}

F3 spectrum_to_rgb(const Tables& t, const float* power441, float factor) {
  Spd s = {};
  s.entry_count = 441;
  for (int i = 0; i < 441; ++i) s.entries[i] = {float(i + 390), power441[i]};
  F3 rgb = xyz_to_rgb(integrate_to_xyz(t, s) * factor);
  return {fmaxf(0.0f, rgb.x), fmaxf(0.0f, rgb.y), fmaxf(0.0f, rgb.z)};
}

// generate_sun_image (scattering.cxx:343-384)
void sun_image(const Tables& t, const Medium& m, const AtmosphereParameters& prm, F3 light, float angular_size, uint32_t w, uint32_t h, float* rgba) {
  F3 bu = normalize(((light.x != light.y) || (light.x != light.z)) ? F3{light.z - light.y, light.x - light.z, +light.y - light.x} : F3{light.z - light.y, light.x + light.z, -light.y - light.x});
  F3 bv = normalize(cross(light, bu));
  const float tan_half = tanf(0.5f * angular_size);
  parallel_chunks(w * h, 256u, [&](uint32_t, uint32_t begin, uint32_t end) {
    float spectrum[441];
    for (uint32_t i = begin; i < end; ++i) {
      uint32_t x = i % w, y = i / w;
      float u = float(x + 0.5f) / float(w) * 2.0f - 1.0f;
      float v0 = float(y + 0.5f) / float(h) * 2.0f - 1.0f, v1 = float(y + 1.5f) / float(h) * 2.0f - 1.0f;
      F3 d0 = normalize((bu * u + bv * v0) * tan_half + light), d1 = normalize((bu * u + bv * v1) * tan_half + light);
      sun_extinction(m, d0, d1, prm, spectrum);
      float darkening = (1.0f - 0.6f * (1.0f - fmaxf(0.0f, 1.0f - (u * u + v0 * v0))));
      F3 rgb = spectrum_to_rgb(t, spectrum, darkening);
      float* o = rgba + size_t(x + w * y) * 4;
      o[0] = rgb.x, o[1] = rgb.y, o[2] = rgb.z, o[3] = 1.0f;
    }
  });
}

// generate_sky_image (scattering.cxx:277-341): the dome, rows stored top down, plus the "average of the upper hemisphere" lift
void sky_image(const Tables& t, const Medium& m, const ImageRecord& table, const AtmosphereParameters& prm, F3 light, uint32_t w, uint32_t h, float* rgba) {
  const uint32_t chunks = 512u;
  std::vector<float> partial(size_t(chunks) * 4, 0.0f);
  parallel_chunks(w * h, chunks, [&](uint32_t chunk, uint32_t begin, uint32_t end) {
    float spectrum[441];
    F3 avg = {};
    float weight_sum = 0.0f;
    for (uint32_t i = begin; i < end; ++i) {
      uint32_t x = i % w, y = i / w;
      float u = float(x + 0.5f) / float(w) * 2.0f - 1.0f, v = float(y + 0.5f) / float(h) * 2.0f - 1.0f;
      float phi = u * kPiF, theta = v * (0.5f * kPiF);
      float cos_p = cosf(phi), sin_p = sinf(phi), cos_t = cosf(theta), sin_t = sinf(theta);
      F3 direction = {1.0f * cos_p * cos_t, 1.0f * sin_t, 1.0f * sin_p * cos_t};  // from_spherical (math.hxx:961-971)
      sky_radiance(m, table, direction, light, prm, spectrum);
      F3 rgb = spectrum_to_rgb(t, spectrum, 1.0f);
      if (v > 0.0f) {
        float weight = sinf(v * (0.5f * kPiF));
        weight_sum += weight;
        avg = avg + rgb * weight;
      }
      float* o = rgba + size_t(x + w * (h - y - 1u)) * 4;
      o[0] = rgb.x, o[1] = rgb.y, o[2] = rgb.z, o[3] = 1.0f;
    }
    float* p = partial.data() + size_t(chunk) * 4;
    p[0] = avg.x, p[1] = avg.y, p[2] = avg.z, p[3] = weight_sum;
  });
  float total[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (uint32_t c = 0; c < chunks; ++c)
    for (int k = 0; k < 4; ++k) total[k] = total[k] + partial[size_t(c) * 4 + k];
  const F3 average = F3{total[0], total[1], total[2]} / total[3];
  const float two_pi = 2.0f * kPiF;
  for (size_t i = 0; i < size_t(w) * h; ++i) {
    float* o = rgba + i * 4;
    o[0] += two_pi * average.x * o[0] + average.x;
    o[1] += two_pi * average.y * o[1] + average.y;
    o[2] += two_pi * average.z * o[2] + average.z;
  }
}

}  // namespace atmosphere
