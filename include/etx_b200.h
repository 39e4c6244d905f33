/* etx_b200.h — C ABI of the B200-native wavefront VCM module (libetx_b200.so).
 *
 * This is the drop-in boundary for ONE path of serhii-rieznik/etx-tracer: everything the reference's
 * CPU VCM integrator does between `CPUVCM::run()` and the film (sources/etx/rt/integrators/vcm_cpu.cxx:81-241,
 * vcm_shared.cxx:49-152, sources/etx/rt/shared/vcm_shared.hxx, sources/etx/rt/rt.cxx:327-579).
 * A host adapter `GPUVCM : Integrator` (etx_tracer_b200/host/gpu_vcm.hpp) forwards the reference's
 * Integrator vtable (sources/etx/rt/integrators/integrator.hxx:12-98) to these calls; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions: plain C, no C++/torch types; all functions return 0 on success or a negative etxb_error;
 * the caller owns every host buffer it passes, the module owns all device memory; one context drives one
 * CUDA device (one process per GPU; multi-GPU sharding is expressed with etxb_set_partition + the
 * exchange buffers below, the collective itself is issued by the host with NCCL).
 *
 * The scene is passed in the reference's OWN in-memory layout (struct etx::Scene, 528 B, and
 * struct etx::Camera, 176 B — sources/etx/render/shared/scene.hxx:22-65, camera.hxx:8-39), so the
 * reference-side call is `etxb_upload_scene(ctx, &rt.scene(), sizeof(Scene), &rt.camera(), sizeof(Camera))`.
 * The etxb_* structs below are byte-compatible C mirrors of those PODs for callers that are not the
 * reference (tests, bench); oracle/layout_check.cxx static_asserts every offset against the real headers.
 */
#ifndef ETX_B200_H
#define ETX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ETXB_INVALID_INDEX 0xffffffffu

typedef enum etxb_error {
  ETXB_OK = 0,
  ETXB_ERR_INVALID_ARGUMENT = -1,
  ETXB_ERR_NO_DEVICE = -2,
  ETXB_ERR_CUDA = -3,
  ETXB_ERR_UNSUPPORTED = -4,
  ETXB_ERR_NOT_READY = -5,
  ETXB_ERR_OUT_OF_MEMORY = -6,
  ETXB_ERR_OVERFLOW = -7
} etxb_error;

/* ---- byte-compatible mirrors of the reference PODs ------------------------------------------------ */

/* etx::ArrayView<T> (base.hxx:52-56) */
typedef struct etxb_array_view {
  const void* a;
  uint64_t count;
} etxb_array_view;

/* etx::Vertex (math.hxx:599), 56 B */
typedef struct etxb_vertex {
  float pos[3], nrm[3], tan[3], btn[3], tex[2];
} etxb_vertex;

/* etx::Triangle (math.hxx:607), 32 B */
typedef struct etxb_triangle {
  uint32_t i[3];
  uint32_t material_index;
  float geo_n[3];
  float pad;
} etxb_triangle;

/* etx::SpectralImage / SampledImage / RefractiveIndex / Thinfilm / SubsurfaceMaterial (material.hxx:8-50) */
typedef struct etxb_spectral_image {
  uint32_t spectrum_index, image_index;
} etxb_spectral_image;
typedef struct etxb_sampled_image {
  float value[4];
  uint32_t image_index, channel;
} etxb_sampled_image;
typedef struct etxb_refractive_index {
  uint32_t cls, eta_index, k_index;
} etxb_refractive_index;
typedef struct etxb_thinfilm {
  etxb_refractive_index ior;
  uint32_t thickness_image;
  float min_thickness, max_thickness, pad;
} etxb_thinfilm;
typedef struct etxb_subsurface {
  uint32_t spectrum_index, image_index, cls, path;
} etxb_subsurface;

/* etx::Material::Class (material.hxx:53-68) */
enum {
  ETXB_MAT_DIFFUSE = 0,
  ETXB_MAT_TRANSLUCENT = 1,
  ETXB_MAT_PLASTIC = 2,
  ETXB_MAT_CONDUCTOR = 3,
  ETXB_MAT_DIELECTRIC = 4,
  ETXB_MAT_THINFILM = 5,
  ETXB_MAT_MIRROR = 6,
  ETXB_MAT_BOUNDARY = 7,
  ETXB_MAT_VELVET = 8,
  ETXB_MAT_PRINCIPLED = 9,
  ETXB_MAT_VOID = 10
};

/* etx::Material (material.hxx:52-97), 200 B */
typedef struct etxb_material {
  etxb_spectral_image reflectance, scattering, emission;
  etxb_sampled_image roughness, metalness, transmission;
  etxb_subsurface subsurface;
  etxb_thinfilm thinfilm;
  etxb_refractive_index ext_ior, int_ior;
  uint32_t cls, int_medium, ext_medium, normal_image_index, diffuse_variation, two_sided;
  float normal_scale, opacity, emission_collimation;
} etxb_material;

/* etx::EmitterProfile (emitter.hxx:7-42), 48 B; Class: 0 Area, 1 Environment, 2 Directional */
typedef struct etxb_emitter_profile {
  etxb_spectral_image emission;
  float direction[3];
  uint32_t cls;
  float angular_size, equivalent_disk_size, angular_size_cosine, pad0, pad1, pad2;
} etxb_emitter_profile;

/* etx::Emitter (emitter.hxx:44-71), 32 B */
typedef struct etxb_emitter {
  uint32_t cls, profile, triangle_index;
  float spectrum_weight, additional_weight, triangle_area, pad0, pad1;
} etxb_emitter;

/* etx::SpectralDistribution (spectrum.hxx:449-), 3552 B */
typedef struct etxb_spectrum {
  struct {
    float wavelength, power;
  } entries[441];
  uint32_t entry_count;
  float integrated[3];
  uint32_t pad[2];
} etxb_spectrum;

/* etx::Distribution::Entry / Distribution (distribution.hxx:7-15) */
typedef struct etxb_distribution_entry {
  float value, pdf, cdf;
} etxb_distribution_entry;
typedef struct etxb_distribution {
  etxb_array_view values; /* etxb_distribution_entry[count] */
  float total_weight;
  uint32_t pad[3];
} etxb_distribution;

/* etx::Image (image.hxx:8-50), 112 B; format 1 = RGBA32F, 2 = RGBA8 */
typedef struct etxb_image {
  etxb_array_view pixels;
  etxb_array_view x_distributions; /* etxb_distribution[isize.y] */
  etxb_distribution y_distribution;
  float fsize[2], offset[2], scale[2];
  uint32_t isize[2];
  float normalization;
  uint32_t options, format, data_size;
} etxb_image;

/* etx::Medium (medium.hxx:8-47), 80 B */
typedef struct etxb_medium {
  etxb_array_view density; /* float[dx*dy*dz], normalised to max 1 */
  float bounds_min[3], bounds_pad0, bounds_max[3], bounds_pad1;
  uint16_t cls, enable_explicit_connections;
  uint32_t absorption_index, scattering_index;
  float phase_function_g, max_sigma;
  uint32_t dimensions[3];
} etxb_medium;

/* etx::Camera (camera.hxx:8-39), 176 B; cls 0 = Perspective, 1 = Equirectangular */
typedef struct etxb_camera {
  float view_proj[16]; /* column-major float4 col[4] */
  float position[3];
  uint32_t cls;
  float target[3], tan_half_fov;
  float side[3], aspect;
  float up[3], area;
  float direction[3], image_plane;
  uint32_t film_size[2];
  float lens_radius, focal_distance;
  float clip_near, clip_far;
  uint32_t lens_image, medium_index;
} etxb_camera;

/* etx::Scene (scene.hxx:22-65), 528 B */
#define ETXB_SCENE_COMMITTED 1u
#define ETXB_SCENE_SPECTRAL 2u
typedef struct etxb_scene {
  etxb_array_view vertices, triangles, triangle_to_emitter, materials, emitter_profiles, emitter_instances, images, mediums, spectrums;
  etxb_distribution emitters_distribution;
  uint32_t environment_emitters[63];
  uint32_t environment_emitter_count;
  float bounding_sphere_center[3], bounding_sphere_radius;
  uint32_t pixel_sampler_image;
  float pixel_sampler_radius;
  uint32_t min_path_length, max_path_length, samples, random_path_termination;
  float noise_threshold, radiance_clamp;
  uint32_t black_spectrum, white_spectrum, rayleigh_spectrum, mie_spectrum, ozone_spectrum;
  uint32_t subsurface_scatter_material, subsurface_exit_material;
  uint32_t default_dielectric_eta, default_conductor_eta, default_conductor_k;
  uint32_t flags;
  uint32_t pad;
} etxb_scene;

/* ---- options / status -------------------------------------------------------------------------- */

/* VCMOptions bit flags (vcm_shared.hxx:24-37) */
#define ETXB_VCM_CONNECT_TO_CAMERA (1u << 0)
#define ETXB_VCM_DIRECT_HIT (1u << 1)
#define ETXB_VCM_CONNECT_TO_LIGHT (1u << 2)
#define ETXB_VCM_CONNECT_VERTICES (1u << 3)
#define ETXB_VCM_MERGE_VERTICES (1u << 4)
#define ETXB_VCM_ENABLE_MIS (1u << 5)
#define ETXB_VCM_ENABLE_MERGING (1u << 6)
#define ETXB_VCM_CONNECT_ONLY (ETXB_VCM_DIRECT_HIT | ETXB_VCM_CONNECT_TO_LIGHT | ETXB_VCM_CONNECT_TO_CAMERA | ETXB_VCM_CONNECT_VERTICES | ETXB_VCM_ENABLE_MIS)
#define ETXB_VCM_FULL (ETXB_VCM_CONNECT_ONLY | ETXB_VCM_ENABLE_MERGING | ETXB_VCM_MERGE_VERTICES)

/* Mirrors VCMOptions (vcm_shared.hxx:12-72); defaults = VCMOptions::default_values (vcm_shared.cxx:6-13).
 * The option keys the reference UI uses ("vcm-initial_radius", "vcm-radius_decay", "vcm-blue_noise", "vcm-kernel",
 * "vcm-direct_hit", ... vcm_shared.cxx:15-28) map 1:1 onto these fields; see etxb_options_set_key. */
typedef struct etxb_vcm_options {
  uint32_t options;      /* ETXB_VCM_* bits, default ETXB_VCM_FULL */
  uint32_t radius_decay; /* default 256 */
  uint32_t kernel;       /* 0 Tophat, 1 Epanechnikov (default) */
  float initial_radius;  /* 0 => 5 * R_scene / max(W,H) (vcm_cpu.cxx:102-106) */
  uint32_t blue_noise;   /* default 1 */
} etxb_vcm_options;

/* Mirrors Integrator::Status (integrator.hxx:24-37) + device counters. */
typedef struct etxb_status {
  double last_iteration_time; /* seconds, device time of the last finished iteration */
  double total_time;          /* seconds, sum over finished iterations */
  uint32_t completed_iterations;
  uint32_t current_iteration;
  uint32_t iteration_in_flight; /* 1 while an enqueued iteration has not finished */
  uint32_t light_vertices;      /* vertices stored by the last light pass */
  uint32_t overflow;            /* non-zero if a fixed-capacity pool overflowed (results invalid) */
  uint32_t pad;
} etxb_status;

/* Per-iteration event counters (feed the algorithmic-bytes formula). */
typedef struct etxb_counters {
  uint64_t rays_closest;    /* closest-hit queries */
  uint64_t rays_shadow;     /* transmittance (any-hit) queries */
  uint64_t nodes_visited;   /* only counted when built with ETXB_COUNT_TRAVERSAL */
  uint64_t tris_tested;     /* idem */
  uint64_t bounces_light;   /* light-subpath surface events */
  uint64_t bounces_camera;  /* camera-subpath surface events */
  uint64_t light_vertices;  /* stored light vertices */
  uint64_t connections;     /* camera-vertex x light-vertex connection attempts */
  uint64_t merge_queries;   /* hash-grid gathers */
  uint64_t merge_candidates;
  uint64_t merge_accepts;
  uint64_t splats;
  uint64_t kernel_launches; /* kernels of this module launched */
  uint64_t nodes_closest;   /* the part of nodes_visited / tris_tested spent by the closest-hit kernel (k_trace_closest) */
  uint64_t tris_closest;
} etxb_counters;

typedef struct etxb_device_config {
  int32_t device_index;        /* CUDA device ordinal */
  uint32_t max_light_vertices; /* pool capacity; 0 => 16 per pixel */
  uint32_t flags;              /* reserved */
  uint32_t pad;
} etxb_device_config;

/* Film layers readable with etxb_read_film: values follow Film::layer (film.cxx:381-418). */
enum {
  ETXB_FILM_RESULT = 0,      /* max(0, camera + light) as float4 */
  ETXB_FILM_CAMERA = 1,      /* camera image (running mean over iterations) */
  ETXB_FILM_LIGHT = 2,       /* light image (running mean over iterations) */
  ETXB_FILM_LIGHT_ITERATION = 3,
  /* layers the path tracer fills (Film::Normals shown as n * 0.5 + 0.5, Film::Albedo, Film::CameraAdaptive; film.cxx:406-414) */
  ETXB_FILM_NORMALS = 4,
  ETXB_FILM_ALBEDO = 5,
  ETXB_FILM_CAMERA_ADAPTIVE = 6
};

/* Named device buffers for etxb_read_buffer / etxb_device_pointer (parity tests + multi-GPU exchange). */
enum {
  ETXB_BUF_LIGHT_PATH_COUNT = 0,   /* uint32[N]  vertices stored per light path (VCMLightPath::count) */
  ETXB_BUF_LIGHT_PATH_OFFSET = 1,  /* uint32[N]  first vertex of the path in the pool (VCMLightPath::index) */
  ETXB_BUF_LIGHT_PATH_WAVELENGTH = 2, /* float[N] */
  ETXB_BUF_LIGHT_SAMPLER = 3,      /* uint32[N]  sampler seed at the end of the light subpath */
  ETXB_BUF_CAMERA_SAMPLER = 4,     /* uint32[N]  sampler seed at the end of the camera subpath */
  ETXB_BUF_LV_POS = 5,             /* float[3*P] light vertex positions, path-major order */
  ETXB_BUF_LV_THROUGHPUT = 6,      /* float[3*P] */
  ETXB_BUF_LV_MIS = 7,             /* float[3*P] d_vcm, d_vc, d_vm */
  ETXB_BUF_FILM_LIGHT_ITERATION = 8, /* float4[N] per-iteration light splats (all-reduced across ranks) */
  ETXB_BUF_FILM_CAMERA = 9,        /* float4[N] */
  ETXB_BUF_FILM_LIGHT = 10,        /* float4[N] */
  ETXB_BUF_PHOTON_RECORDS = 11,    /* packed photon records of this rank (all-gathered across ranks) */
  ETXB_BUF_CAMERA_GATHERED = 12,   /* float[3*N] last iteration's camera contribution per path */
  ETXB_BUF_PIXEL_INFO = 13,        /* uint32[N]  Film's InternalData in film storage order (film.cxx:27-32): sample_count bits 0-29, converged bit 30, tmp bit 31 */
  ETXB_BUF_PIXEL_ERROR = 14,       /* float[N]   InternalData::error_level */
  ETXB_BUF_COUNT
};

typedef struct etxb_ctx etxb_ctx;

/* ---- entry points -------------------------------------------------------------------------------- */

/* Library identity: "fast" (product build) or "parity" (strict-IEEE build used by the bit-exact tests). */
const char* etxb_build_flavor(void);
/* Number of CUDA devices visible, or a negative etxb_error. */
int etxb_device_count(void);

/* Replaces Raytracing/Integrator construction (rt.cxx:27-36, vcm_cpu.cxx:62-66). */
int etxb_create(etxb_ctx** out_ctx, const etxb_device_config* cfg);
void etxb_destroy(etxb_ctx* ctx);
const char* etxb_last_error(const etxb_ctx* ctx);

/* Replaces Raytracing::commit_changes (rt.cxx:58-64,323): copies the scene to HBM, builds the BVH,
 * allocates film + queues for camera->film_size.  `scene`/`camera` use the reference layouts. */
int etxb_upload_scene(etxb_ctx* ctx, const void* scene, uint64_t scene_bytes, const void* camera, uint64_t camera_bytes);

/* Blue-noise tables (thirdparty/bluenoise): sobol[256*256], and for the variant selected by
 * next_power(min(scene.samples,256)) (bluenoise.cxx:73-95) scrambling[128*128*8], ranking[128*128*8], as uint8. */
int etxb_upload_blue_noise(etxb_ctx* ctx, const uint8_t* sobol_256x256, const uint8_t* scrambling, const uint8_t* ranking);
/* CIE 2006 XYZ table, 441 x float3, 390..830 nm (spectrum.hxx:28 spectral_xyz) and the RGB response table
 * 391 x float3 (spectrum.cxx:399 rgb_response). */
int etxb_upload_color_tables(etxb_ctx* ctx, const float* xyz_441x3, const float* rgb_response_391x3);

/* VCMOptions::default_values / load (vcm_shared.cxx:6-28). */
void etxb_options_default(etxb_vcm_options* opt);
int etxb_options_set_key(etxb_vcm_options* opt, const char* key, double value);
int etxb_set_options(etxb_ctx* ctx, const etxb_vcm_options* opt);

/* Pixel-tile sharding across ranks (one process per GPU): rank r owns 32x32 tiles with (tile % world) == r.
 * world == 1 (default) renders the full frame. */
int etxb_set_partition(etxb_ctx* ctx, uint32_t rank, uint32_t world);

/* Iteration-interleaved multi-GPU runs: this context renders iterations first_iteration, first_iteration + stride, ...
 * (VCMIteration::iteration drives the merge radius, vcm_cpu.cxx:95-113, and the per-path sampler seeds, vcm_shared.hxx:286-300),
 * and its film is the mean over the iterations it rendered.  Default stride 1 = the reference's sequence. */
int etxb_set_iteration_stride(etxb_ctx* ctx, uint32_t stride);

/* Externally driven iteration index: the next enqueued iteration renders VCMIteration::iteration = `iteration`; the index no
 * longer advances by itself (used by etxb_group below). */
int etxb_set_next_iteration(etxb_ctx* ctx, uint32_t iteration);

/* CPUVCMImpl::start (vcm_cpu.cxx:81-93): clears film, sets iteration = first_iteration. */
int etxb_begin(etxb_ctx* ctx, uint32_t first_iteration);

/* ---- film export: RTApplication::on_save_image_selected (sources/raytracer/app.cxx:261-295) -------------------------
 * The reference saves the selected Film layer either as float OpenEXR (tinyexr SaveEXR) or tone-mapped to 8-bit PNG (1 - exp(-exposure c),
 * sRGB curve, stb_image_write).  etxb_read_film_ldr tone-maps on the device (4 bytes per pixel leave the GPU); the two writers and the host
 * tone map are plain host code (usable on any float4 / RGBA8 buffer, e.g. a multi-GPU frame from etxb_group_comm_reduce_film). */
#define ETXB_SAVE_EXR 0u
#define ETXB_SAVE_PNG_TONEMAPPED 1u
int etxb_read_film_ldr(etxb_ctx* ctx, uint32_t layer, float exposure, uint8_t* dst_rgba8, uint64_t dst_bytes);
int etxb_save_film(etxb_ctx* ctx, uint32_t layer, const char* file_name, uint32_t mode, float exposure);
int etxb_write_exr(const char* file_name, const float* rgba, uint32_t width, uint32_t height);
int etxb_write_png(const char* file_name, const uint8_t* rgba8, uint32_t width, uint32_t height);
int etxb_tonemap_rgba8(const float* rgba, uint64_t pixel_count, float exposure, uint8_t* out_rgba8);

/* ---- scene files: SceneRepresentation::load_from_file (render/host/scene_representation.cxx:679-838) ------------------
 * Reads the reference's `.json` + `.obj` + `.mtl` scene dialect (et::camera / et::medium / et::dir / et::env / et::spectrum blocks, the material
 * directives of parse_material :1682-2079, PNG / JPEG / TGA / BMP / EXR / HDR / PFM textures) into the Scene / Camera PODs etxb_create takes: host C++, no CUDA, no
 * third-party reader.  The object owns every array the PODs point into; keep it alive until etxb_create has returned (the module copies).
 * A file without an et::dir / et::env block gets the reference's default atmosphere (sun + sky images, generated here on the host threads).
 * `et::medium ... volume file.nvdb` becomes a heterogeneous medium with a dense grid (own NanoVDB 32.x reader, codecs none / ZIP; medium_pool.cxx:102-159).
 * Failure: ETXB_ERR_UNSUPPORTED with the reason in `err` (missing or corrupt file, BLOSC-compressed volume, glTF).
 * `data_folder` = where tables.bin lives; NULL = the `data/` folder beside the shared library. */
typedef struct etxb_scene_file etxb_scene_file;
int etxb_scene_file_load(const char* file_name, const char* data_folder, etxb_scene_file** out, char* err, uint64_t err_bytes);
const etxb_scene* etxb_scene_file_scene(const etxb_scene_file* sf);
const etxb_camera* etxb_scene_file_camera(const etxb_scene_file* sf);
uint32_t etxb_scene_file_warning_count(const etxb_scene_file* sf);
const char* etxb_scene_file_warning(const etxb_scene_file* sf, uint32_t index);
uint32_t etxb_scene_file_material_count(const etxb_scene_file* sf);
const char* etxb_scene_file_material_name(const etxb_scene_file* sf, uint32_t index);
void etxb_scene_file_free(etxb_scene_file* sf);
/* Scene::samples (the iteration count the integrators stop at, and what selects the blue-noise variant) — the application's "samples" setting */
void etxb_scene_file_set_samples(etxb_scene_file* sf, uint32_t samples);
/* One call from file to device: etxb_upload_color_tables + etxb_upload_blue_noise (the variant BNSampler picks for scene.samples,
 * thirdparty/bluenoise/bluenoise.cxx:73-95) + etxb_upload_scene, all from tables.bin and the loaded PODs. */
int etxb_scene_file_commit(etxb_ctx* ctx, const etxb_scene_file* sf);
/* A string value of the application's options file (util/options.cxx; ids "integrator" and "scene" are what RTApplication::init reads,
 * raytracer/app.cxx:88-105): copies it to `out` (NUL-terminated, truncated to out_bytes) and returns its length; 0 = id absent; < 0 = error. */
int etxb_options_file_string(const char* file_name, const char* id, char* out, uint64_t out_bytes);
/* The loader's image readers on their own — every PNG form stb_image decodes (bit depths 1-16, palette, colour keys, Adam7), baseline / progressive JPEG
 * (same bytes as stb_image's decoder), TGA, BMP, OpenEXR scan lines (none /
 * RLE / ZIPS / ZIP / PIZ; half and float), Radiance HDR, the reference's PFM variant — as ImagePool::load_data uses them (image_pool.cxx:271-383): rows in file order,
 * RGBA8 (*eight_bit = 1; the file's values, no sRGB step) or RGBA32F.  pixels = NULL queries the size. */
int etxb_image_file_read(const char* file_name, uint32_t* width, uint32_t* height, uint32_t* eight_bit, void* pixels, uint64_t capacity, char* err, uint64_t err_bytes);
/* Two parts of the loader on their own.  etxb_mesh_tangents: the tangent-space generator the reference calls for meshes with texture coordinates
 * (build_tangents, scene_representation.cxx:337-398) over etx::Vertex / etx::Triangle arrays, writing tan / btn of the vertices whose tangent is not valid
 * yet.  etxb_nvdb_density: the dense grid MediumPool::load_nvdb (medium_pool.cxx:102-159) builds from the first float grid of a NanoVDB file — call with
 * values = NULL for the dimensions (x, y, z; all 0 = the medium stays homogeneous), then with a buffer of x*y*z floats (x fastest; not yet normalised). */
int etxb_mesh_tangents(etxb_vertex* vertices, uint64_t vertex_count, const etxb_triangle* triangles, uint64_t triangle_count);
int etxb_nvdb_density(const char* file_name, uint32_t dimensions[3], float* values, uint64_t value_capacity, char* err, uint64_t err_bytes);
/* The procedural sun disk (128 x 128) and sky dome (sky_width x sky_height) images of an atmosphere block — what the loader generates for a
 * `newmtl et::atmosphere` block and for every scene file that declares no distant emitter (scene_representation.cxx:805-820, :1376-1495;
 * render/host/scattering.cxx) — for callers that assemble the emitters themselves.  parameters = altitude, anisotropy, rayleigh, mie, ozone
 * (scattering::Parameters); direction normalised; angular_size in radians; either output may be NULL. */
int etxb_atmosphere_images(const char* data_folder, const float direction[3], float angular_size, const float parameters[5], uint32_t sky_width, uint32_t sky_height,
                           float* sun_rgba_128x128, float* sky_rgba);
/* A table of tables.bin by name ("color_tables/xyz_441x3", "bluenoise/sobol", "spectra/gold.eta_power", ...); NULL when absent. */
const void* etxb_scene_file_table(const etxb_scene_file* sf, const char* name, uint64_t* bytes);

/* ---- the unidirectional path tracer on the same context -----------------------------------------
 * Replaces CPUPathTracing (rt/integrators/path_tracing.cxx:12-170) + run_path_iteration (rt/shared/path_tracing_shared.hxx:485-510)
 * + Film::accumulate_camera_image with the normal / albedo layers and Film::estimate_noise_levels (render/host/film.cxx:173-330).
 * etxb_set_integrator selects which algorithm etxb_begin / etxb_enqueue_iteration / etxb_poll / etxb_read_film drive: with
 * ETXB_INTEGRATOR_PT one iteration = one path per active pixel (CPUPathTracingImpl::execute_range) followed by update()'s
 * film.estimate_noise_levels(current_iteration, samples, noise_threshold); a pixel the estimate marks converged is skipped from
 * then on (Film::active_pixel).  Progressive preview (pixel_size > 1, film.cxx:440-449) is not part of it: its pixel choice is rand(). */
#define ETXB_INTEGRATOR_VCM 0u
#define ETXB_INTEGRATOR_PT 1u
/* PTOptions (path_tracing_shared.hxx:8-14); the reference's option ids are "nee", "direct", "mis", "bn" (path_tracing.cxx:36-39) */
typedef struct etxb_pt_options {
  uint32_t nee, direct, mis, blue_noise; /* all default 1 */
} etxb_pt_options;
typedef struct etxb_pt_status {
  uint32_t pixels_processed; /* CPUPathTracingImpl::pixels_processed of the last finished iteration; 0 => the reference stops the run (:90-92) */
  uint32_t active_pixels;    /* Film::active_pixel_count (film.cxx:430): pixel count after a clear, pixels converged in the last estimate after one */
  float noise_level;         /* Film::noise_level (film.cxx:461) */
  uint32_t max_sample_count; /* Scene::samples */
} etxb_pt_status;
void etxb_pt_options_default(etxb_pt_options* opt);
int etxb_pt_options_set_key(etxb_pt_options* opt, const char* key, double value);
int etxb_pt_set_options(etxb_ctx* ctx, const etxb_pt_options* opt);
int etxb_set_integrator(etxb_ctx* ctx, uint32_t integrator);
int etxb_pt_get_status(etxb_ctx* ctx, etxb_pt_status* out);
/* Scene::noise_threshold / Scene::radiance_clamp (scene.hxx:45-46) changed without a new upload */
int etxb_set_scene_settings(etxb_ctx* ctx, float noise_threshold, float radiance_clamp);

/* One VCM iteration = start_next_iteration + gather_light_vertices + complete_light_vertices +
 * gather_camera_vertices + complete_camera_vertices (vcm_cpu.cxx:95-241).  Asynchronous: returns after
 * the work is queued on the context's stream. */
int etxb_enqueue_iteration(etxb_ctx* ctx);
/* The same iteration in three phases for multi-GPU runs (exchange photons / light image in between). */
int etxb_enqueue_light_pass(etxb_ctx* ctx);
int etxb_enqueue_grid_build(etxb_ctx* ctx, const void* device_photon_records, uint64_t photon_count);
int etxb_enqueue_camera_pass(etxb_ctx* ctx);

/* CPUVCM::update (vcm_cpu.cxx:264-276): non-blocking; fills status. */
int etxb_poll(etxb_ctx* ctx, etxb_status* status);
/* Blocks until all queued work is done. */
int etxb_wait(etxb_ctx* ctx);
/* CPUVCM::stop (vcm_cpu.cxx:278-288). */
int etxb_stop(etxb_ctx* ctx, int wait_for_iteration);

/* Film::layer (film.cxx:381-418): row-major float4, y already flipped in storage like the reference. */
int etxb_read_film(etxb_ctx* ctx, uint32_t layer, float* dst_rgba, uint64_t dst_bytes);
int etxb_film_size(const etxb_ctx* ctx, uint32_t* width, uint32_t* height);

int etxb_get_counters(etxb_ctx* ctx, etxb_counters* out);
/* Per-kernel device time of the last iteration (CUDA events): names[i] -> ms[i]; returns count. */
int etxb_get_kernel_times(etxb_ctx* ctx, const char** names, float* ms, uint32_t* launches, uint32_t capacity);

/* Introspection for parity tests and NCCL exchange. */
int etxb_read_buffer(etxb_ctx* ctx, uint32_t buffer_id, void* dst, uint64_t dst_bytes, uint64_t* out_bytes);
int etxb_device_pointer(etxb_ctx* ctx, uint32_t buffer_id, void** out_ptr, uint64_t* out_bytes);
void* etxb_stream(etxb_ctx* ctx); /* cudaStream_t the module launches on */

/* Unit entry points used by the parity tests (same kernels the iteration uses). */
/* Closest-hit query for `count` rays (rt.cxx:428 Raytracing::trace): rays = float[8]*count {o,min_t,d,max_t},
 * seeds in/out = sampler state per ray; hits out = {u, v, tri, t} (IntersectionBase, math.hxx:666). */
int etxb_debug_trace(etxb_ctx* ctx, const float* rays, uint32_t* seeds, uint32_t count, float* hits_uv_t, uint32_t* hits_tri);
/* Which tree etxb_debug_trace walks: 0 = the BVH2 shared with the oracle, 1 = the product build's 4-wide quantised tree (returns 1 if it exists). */
int etxb_debug_select_tree(etxb_ctx* ctx, int wide);
/* Sampler / spectral KATs evaluated on the device (sampler.hxx:54,66; spectrum.hxx:219,234). */
int etxb_debug_sampler(etxb_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t count, uint32_t draws, uint32_t* out_seed, float* out_values);
int etxb_debug_math(etxb_ctx* ctx, uint32_t fn, const float* x, const float* y, uint32_t count, float* out);

/* ---- several iterations in flight on one device -----------------------------------------------------------------------
 * The reference runs ONE iteration at a time and spreads its paths over the CPU's threads (vcm_cpu.cxx:115-171, TaskScheduler).
 * A GPU wavefront iteration ends in a long, latency-bound tail (few paths, ~50 more bounces); an etxb_group keeps `lanes`
 * independent contexts on one device, one host thread each, pulling iteration indices from a shared counter, so that tail overlaps
 * other iterations.  Uploads/options go to every lane (etxb_group_lane); the film is the mean of the lanes' films weighted by the
 * iterations each finished — the same set of iterations as a single context running them in order.
 * Replaces the run()/update() pump of CPUVCM (vcm_cpu.cxx:247-276) for throughput rendering. */
typedef struct etxb_group etxb_group;
int etxb_group_create(etxb_group** out_group, const etxb_device_config* config, uint32_t lanes); /* 1..8 lanes */
void etxb_group_destroy(etxb_group* group);
uint32_t etxb_group_lanes(const etxb_group* group);
etxb_ctx* etxb_group_lane(etxb_group* group, uint32_t lane);
const char* etxb_group_last_error(const etxb_group* group);
int etxb_group_begin(etxb_group* group, uint32_t first_iteration);   /* Integrator::run: clears every lane's film */
int etxb_group_set_stride(etxb_group* group, uint32_t stride);       /* the k-th iteration handed out renders index first + k * stride
                                                                        (iteration-interleaved multi-GPU runs: first = rank, stride = ranks) */
int etxb_group_enqueue(etxb_group* group, uint32_t iterations);      /* asynchronous: queues `iterations` more iterations */
int etxb_group_wait(etxb_group* group);                              /* blocks until the queue has drained; returns the first error */
int etxb_group_poll(etxb_group* group, etxb_status* status);         /* completed = sum over lanes; total_time = wall time with work in flight */
int etxb_group_read_film(etxb_group* group, uint32_t layer, float* dst_rgba, uint64_t dst_bytes);
/* The same combined layer left on the device (float4[N]) + the number of iterations behind it: what a multi-GPU host reduces. */
int etxb_group_combine(etxb_group* group, uint32_t layer, void** device_ptr, uint64_t* bytes, uint32_t* completed_iterations);

/* ---- pixel-tile sharding over several GPUs (one process per GPU, NCCL over NVLink) ----------------------------------------
 * The reference renders one frame with one thread pool (vcm_cpu.cxx:115-241).  Here rank r owns the 32x32 pixel tiles with tile % world == r
 * (etxb_set_partition's rule) and runs the light and camera subpaths of its pixels; what is global is exchanged INSIDE etxb_enqueue_iteration:
 * all-reduce of the per-iteration light image (light-tracing splats land on any pixel, vcm_cpu.cxx:147-154 -> film.cxx:332-343) and an
 * all-gather of the photon records (merging queries every light path's vertices, vcm_cpu.cxx:219-221).  The host only distributes the NCCL id
 * (any transport: MPI, a file, torch.distributed) — rank 0 calls etxb_comm_unique_id, every rank calls etxb_comm_init with those bytes. */
#define ETXB_COMM_ID_BYTES 128
int etxb_comm_unique_id(void* out_id, uint64_t bytes);
int etxb_comm_init(etxb_ctx* ctx, uint32_t world, uint32_t rank, const void* id, uint64_t bytes); /* collective */
int etxb_comm_world(const etxb_ctx* ctx, uint32_t* world, uint32_t* rank);
/* Collective: sums the disjoint camera tiles on rank 0 (ncclReduce of the float4 film); dst_rgba is written on rank 0 only. */
int etxb_comm_reduce_film(etxb_ctx* ctx, uint32_t layer, float* dst_rgba, uint64_t dst_bytes);
/* The same with several iterations in flight per GPU: ids = (lanes + 1) x ETXB_COMM_ID_BYTES (one communicator per lane + one for the frame
 * reduce); lane l renders the iteration ordinals l, l + lanes, ... on every rank. */
int etxb_group_comm_init(etxb_group* group, uint32_t world, uint32_t rank, const void* ids, uint32_t id_count); /* collective */
/* Whole-frame iterations dealt to the ranks instead (the job's j-th iteration on rank j % world; etxb_group_enqueue then counts iterations of the
 * JOB): no collective inside an iteration, one count-weighted ncclReduce of the films per frame.  id = one ETXB_COMM_ID_BYTES id. */
int etxb_group_comm_init_replicas(etxb_group* group, uint32_t world, uint32_t rank, const void* id, uint64_t bytes); /* collective */
/* Optional, before etxb_group_comm_init_replicas: reserves the group's LAST lane for camera-split iterations.  When an etxb_group_enqueue(n >= world)
 * is not a multiple of the ranks, its last n % world iterations are then split over world / (n % world) ranks each: every part traces the whole
 * light pass itself (same photon map, no exchange) and the camera pass of its pixel tiles, and the parts meet in etxb_group_comm_reduce_film —
 * instead of leaving some ranks one whole iteration behind the others. */
int etxb_group_reserve_split_lane(etxb_group* group);
/* Test hook (needs no device): the (ordinal, part, parts) triples rank `rank` of `world` takes of etxb_group_enqueue(iterations) in replica mode. */
int etxb_debug_replica_plan(uint32_t world, uint32_t rank, uint32_t base, uint32_t iterations, int split_lane, uint32_t* out_triples, uint32_t capacity);
int etxb_group_comm_reduce_film(etxb_group* group, uint32_t layer, float* dst_rgba, uint64_t dst_bytes);        /* collective, both modes */

#ifdef __cplusplus
}
#endif
#endif /* ETX_B200_H */
