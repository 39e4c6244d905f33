// TEST INFRASTRUCTURE — compile-time proof that the C mirrors in include/etx_b200.h are byte-compatible
// with the reference PODs they stand for (a mismatch fails the oracle build).
#include <etx/core/core.hxx>
#include <etx/rt/rt.hxx>
#include <etx/render/shared/scene_camera.hxx>
#include <cstddef>
#include "../include/etx_b200.h"

using namespace etx;
#define SAME_SIZE(A, B) static_assert(sizeof(A) == sizeof(B), "size mismatch: " #A " vs " #B)
#define SAME_OFF(A, FA, B, FB) static_assert(offsetof(A, FA) == offsetof(B, FB), "offset mismatch: " #A "." #FA)

SAME_SIZE(etxb_vertex, Vertex);
SAME_SIZE(etxb_triangle, Triangle);
SAME_OFF(etxb_triangle, material_index, Triangle, material_index);
SAME_OFF(etxb_triangle, geo_n, Triangle, geo_n);
SAME_SIZE(etxb_material, Material);
SAME_OFF(etxb_material, scattering, Material, scattering);
SAME_OFF(etxb_material, emission, Material, emission);
SAME_OFF(etxb_material, roughness, Material, roughness);
SAME_OFF(etxb_material, metalness, Material, metalness);
SAME_OFF(etxb_material, transmission, Material, transmission);
SAME_OFF(etxb_material, subsurface, Material, subsurface);
SAME_OFF(etxb_material, thinfilm, Material, thinfilm);
SAME_OFF(etxb_material, ext_ior, Material, ext_ior);
SAME_OFF(etxb_material, int_ior, Material, int_ior);
SAME_OFF(etxb_material, cls, Material, cls);
SAME_OFF(etxb_material, int_medium, Material, int_medium);
SAME_OFF(etxb_material, normal_image_index, Material, normal_image_index);
SAME_OFF(etxb_material, two_sided, Material, two_sided);
SAME_OFF(etxb_material, opacity, Material, opacity);
SAME_OFF(etxb_material, emission_collimation, Material, emission_collimation);
SAME_SIZE(etxb_thinfilm, Thinfilm);
SAME_OFF(etxb_thinfilm, min_thickness, Thinfilm, min_thickness);
SAME_SIZE(etxb_emitter_profile, EmitterProfile);
SAME_OFF(etxb_emitter_profile, direction, EmitterProfile, direction);
SAME_OFF(etxb_emitter_profile, cls, EmitterProfile, cls);
SAME_OFF(etxb_emitter_profile, angular_size_cosine, EmitterProfile, angular_size_cosine);
SAME_SIZE(etxb_emitter, Emitter);
SAME_OFF(etxb_emitter, triangle_area, Emitter, triangle_area);
SAME_SIZE(etxb_spectrum, SpectralDistribution);
SAME_OFF(etxb_spectrum, entry_count, SpectralDistribution, spectral_entry_count);
SAME_SIZE(etxb_distribution, Distribution);
SAME_OFF(etxb_distribution, total_weight, Distribution, total_weight);
SAME_SIZE(etxb_image, Image);
SAME_OFF(etxb_image, x_distributions, Image, x_distributions);
SAME_OFF(etxb_image, y_distribution, Image, y_distribution);
SAME_OFF(etxb_image, fsize, Image, fsize);
SAME_OFF(etxb_image, isize, Image, isize);
SAME_OFF(etxb_image, normalization, Image, normalization);
SAME_OFF(etxb_image, format, Image, format);
SAME_SIZE(etxb_medium, Medium);
SAME_OFF(etxb_medium, bounds_min, Medium, bounds);
SAME_OFF(etxb_medium, cls, Medium, cls);
SAME_OFF(etxb_medium, absorption_index, Medium, absorption_index);
SAME_OFF(etxb_medium, max_sigma, Medium, max_sigma);
SAME_OFF(etxb_medium, dimensions, Medium, dimensions);
SAME_SIZE(etxb_camera, Camera);
SAME_OFF(etxb_camera, position, Camera, position);
SAME_OFF(etxb_camera, cls, Camera, cls);
SAME_OFF(etxb_camera, tan_half_fov, Camera, tan_half_fov);
SAME_OFF(etxb_camera, aspect, Camera, aspect);
SAME_OFF(etxb_camera, area, Camera, area);
SAME_OFF(etxb_camera, direction, Camera, direction);
SAME_OFF(etxb_camera, film_size, Camera, film_size);
SAME_OFF(etxb_camera, lens_radius, Camera, lens_radius);
SAME_OFF(etxb_camera, clip_near, Camera, clip_near);
SAME_OFF(etxb_camera, medium_index, Camera, medium_index);
SAME_SIZE(etxb_scene, Scene);
SAME_OFF(etxb_scene, spectrums, Scene, spectrums);
SAME_OFF(etxb_scene, emitters_distribution, Scene, emitters_distribution);
SAME_OFF(etxb_scene, environment_emitters, Scene, environment_emitters);
SAME_OFF(etxb_scene, bounding_sphere_center, Scene, bounding_sphere_center);
SAME_OFF(etxb_scene, bounding_sphere_radius, Scene, bounding_sphere_radius);
SAME_OFF(etxb_scene, pixel_sampler_image, Scene, pixel_sampler);
SAME_OFF(etxb_scene, min_path_length, Scene, min_path_length);
SAME_OFF(etxb_scene, samples, Scene, samples);
SAME_OFF(etxb_scene, random_path_termination, Scene, random_path_termination);
SAME_OFF(etxb_scene, black_spectrum, Scene, black_spectrum);
SAME_OFF(etxb_scene, subsurface_exit_material, Scene, subsurface_exit_material);
SAME_OFF(etxb_scene, flags, Scene, flags);
