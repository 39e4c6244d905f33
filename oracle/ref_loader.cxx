// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product library.
//
// The reference's OWN scene loader (sources/etx/render/host/scene_representation.cxx:679-2497 with image_pool.cxx, medium_pool.cxx,
// scattering.cxx, spectrum.cxx, ior_database.hxx and its third-party readers) compiled in place from /root/reference and exposed through a
// small C interface, so that tests can hand the Scene / Camera PODs the reference builds from its shipped assets to the oracle (and to the
// CUDA module, whose etxb_upload_scene takes exactly those bytes).  Only what has no Linux implementation in the reference is supplied here:
//   TaskScheduler        (render/host/tasks.cxx sizes its pimpl for MSVC; tasks.hxx:24-49 is the interface) -> serial execution
//   log::set_console_color (core/windows.cxx:85 / macos.cxx)                                                -> no-op
//   Film::generate_filter_image (render/host/film.cxx:123-135; film.cxx itself needs the OIDN headers)      -> restated
// Build: oracle/build_oracle.sh -> oracle/_ref/libreference_loader.so.
#include <etx/core/core.hxx>
#include <etx/core/environment.hxx>
#include <etx/core/log.hxx>
#include <etx/render/host/film.hxx>
#include <etx/render/host/scene_representation.hxx>
#include <etx/render/host/tasks.hxx>
#include <etx/render/shared/ior_database.hxx>

#include <cstring>
#include <memory>
#include <string>

namespace etx {

// ---- TaskScheduler (tasks.hxx:24-49): the loader only needs the blocking forms; one thread, whole range in one call ----
TaskScheduler::TaskScheduler() {}
TaskScheduler::~TaskScheduler() {}
uint32_t TaskScheduler::max_thread_count() { return 1u; }
void TaskScheduler::register_thread() {}
Task::Handle TaskScheduler::schedule(uint32_t range, Task* t) {
  if (t && range) t->execute_range(0u, range, 0u);
  return {};
}
Task::Handle TaskScheduler::schedule(uint32_t range, std::function<void(uint32_t, uint32_t, uint32_t)> func) {
  if (func && range) func(0u, range, 0u);
  return {};
}
void TaskScheduler::execute(uint32_t range, Task* t) {
  if (t && range) t->execute_range(0u, range, 0u);
}
void TaskScheduler::execute(uint32_t range, std::function<void(uint32_t, uint32_t, uint32_t)> func) {
  if (func && range) func(0u, range, 0u);
}
void TaskScheduler::execute_linear(uint32_t range, std::function<void(uint32_t, uint32_t, uint32_t)> func) {
  if (func && range) func(0u, range, 0u);
}
bool TaskScheduler::completed(Task::Handle) { return true; }
void TaskScheduler::wait(Task::Handle& h) { h = {}; }
void TaskScheduler::restart(Task::Handle) {}

void log::set_console_color(log::Color) {}

// Film::generate_filter_image (film.cxx:123-135) with filter_blackman_harris (film.cxx:63-67)
static float loader_filter_blackman_harris(const float2& p, float radius) {
  float sample_distance = sqrtf(p.x * p.x + p.y * p.y);
  float r = kDoublePi * saturate(0.5f + sample_distance / (2.0f * radius));
  return 0.35875f - 0.48829f * cosf(r) + 0.14128f * cosf(2.0f * r) - 0.01168f * cosf(3.0f * r);
}
void Film::generate_filter_image(uint32_t, std::vector<float4>& data) {
  constexpr float2 center = {float(PixelFilterSize) * 0.5f, float(PixelFilterSize) * 0.5f};
  constexpr float radius = float(PixelFilterSize) * 0.5f;
  data.resize(PixelFilterSize * PixelFilterSize);
  for (uint32_t y = 0; y < PixelFilterSize; ++y) {
    for (uint32_t x = 0; x < PixelFilterSize; ++x) {
      float value = loader_filter_blackman_harris(float2{float(x), float(y)} - center, radius);
      data[x + y * PixelFilterSize] = {value, value, value, 1.0f};
    }
  }
}

}  // namespace etx

namespace {
struct LoadedScene {
  etx::TaskScheduler scheduler;
  etx::IORDatabase ior;
  std::unique_ptr<etx::SceneRepresentation> repr;
};
}  // namespace

extern "C" {

// data_folder: the reference's `bin/` (IOR database under bin/spectrum); scene_file: a scene .json / .obj the reference ships or accepts
void* refloader_load(const char* data_folder, const char* scene_file) {
  auto* ls = new LoadedScene();
  std::string spectrum = std::string(data_folder ? data_folder : ".") + "/spectrum/";
  ls->ior.load(spectrum.c_str());
  ls->repr = std::make_unique<etx::SceneRepresentation>(ls->scheduler, ls->ior);
  if (!ls->repr->load_from_file(scene_file, etx::SceneRepresentation::LoadEverything) || !ls->repr->valid()) {
    delete ls;
    return nullptr;
  }
  return ls;
}
void refloader_free(void* h) { delete static_cast<LoadedScene*>(h); }
const void* refloader_scene(void* h, uint64_t* bytes) {
  *bytes = sizeof(etx::Scene);
  return &static_cast<LoadedScene*>(h)->repr->scene();
}
const void* refloader_camera(void* h, uint64_t* bytes) {
  *bytes = sizeof(etx::Camera);
  return &static_cast<LoadedScene*>(h)->repr->camera();
}
uint32_t refloader_material_index(void* h, const char* name) {
  const auto& m = static_cast<LoadedScene*>(h)->repr->material_mapping();
  auto it = m.find(name);
  return it == m.end() ? ~0u : it->second;
}

}  // extern "C"
