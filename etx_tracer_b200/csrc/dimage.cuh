// dimage.cuh — textures in HBM: bilinear evaluate, importance sampling by row/column CDFs (render/shared/image.hxx:8-188,
// distribution.hxx).  Pixels stay in the reference's formats (RGBA32F float4 / RGBA8 ubyte4); the per-row distributions are
// flattened into one Entry array of H*(W+1) records + one of H+1.
#pragma once
#include "../../include/etx_b200.h"
#include "dcore.cuh"

namespace etxb {

enum : uint32_t {  // Image options (image.hxx:15-25)
  kImageRepeatU = 1u << 1u,
  kImageRepeatV = 1u << 2u,
  kImageHasAlpha = 1u << 4u,
  kImageUniformSamplingTable = 1u << 5u,
};

struct DImage {
  const float4* pixels_f32;
  const uchar4* pixels_u8;
  const etxb_distribution_entry* x_dist;  // isize.y rows of (isize.x + 1) entries
  const etxb_distribution_entry* y_dist;  // isize.y + 1 entries
  float fsize_x, fsize_y, offset_x, offset_y, scale_x, scale_y;
  uint32_t isize_x, isize_y;
  float normalization;
  uint32_t options, format;
  uint32_t has_distribution;
};

struct F4v {
  float x, y, z, w;
};
DEV F4v operator*(F4v a, float b) { return {a.x * b, a.y * b, a.z * b, a.w * b}; }
DEV F4v operator+(F4v a, F4v b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

DEV F4v image_pixel(const DImage& img, uint32_t x, uint32_t y) {
  uint32_t i = umin(x + y * img.isize_x, img.isize_x * img.isize_y - 1u);
  if (img.format == 2u) {  // RGBA8: to_float4(ubyte4) (math.hxx:709-711)
    uchar4 p = __ldg(&img.pixels_u8[i]);
    return {float(p.x) / 255.0f, float(p.y) / 255.0f, float(p.z) / 255.0f, float(p.w) / 255.0f};
  }
  float4 p = __ldg(&img.pixels_f32[i]);
  return {p.x, p.y, p.z, p.w};
}
DEV float clampf(float v, float lo, float hi) { return (v < lo) ? lo : (v > hi ? hi : v); }
DEV uint32_t clampu(uint32_t v, uint32_t lo, uint32_t hi) { return (v < lo) ? lo : (v > hi ? hi : v); }
DEV float tex_coord(float u, float size, bool repeat) {
  if (repeat) {
    float x = fmodf(u, size);
    return x < 0.0f ? (x + size) : x;
  }
  return clampf(u, 0.0f, nextafterf(size, 0.0f));
}

struct ImageGather {
  F4v p00, p01, p10, p11;
};
// Image::gather (image.hxx:52-74)
DEV ImageGather image_gather(const DImage& img, V2 in_uv) {
  V2 uv = {in_uv.x * img.fsize_x, in_uv.y * img.fsize_y};
  float x0 = tex_coord(uv.x, img.fsize_x, img.options & kImageRepeatU);
  float y0 = tex_coord(uv.y, img.fsize_y, img.options & kImageRepeatV);
  float dx = x0 - floorf(x0);
  float dy = y0 - floorf(y0);
  uint32_t row_0 = clampu(static_cast<uint32_t>(y0), 0u, img.isize_y - 1u);
  uint32_t row_1 = clampu(row_0 + 1u, 0u, img.isize_y - 1u);
  uint32_t col_0 = clampu(static_cast<uint32_t>(x0), 0u, img.isize_x - 1u);
  uint32_t col_1 = clampu(col_0 + 1u, 0u, img.isize_x - 1u);
  ImageGather g;
  g.p00 = image_pixel(img, col_0, row_0) * (1.0f - dx) * (1.0f - dy);
  g.p01 = image_pixel(img, col_1, row_0) * (dx) * (1.0f - dy);
  g.p10 = image_pixel(img, col_0, row_1) * (1.0f - dx) * (dy);
  g.p11 = image_pixel(img, col_1, row_1) * (dx) * (dy);
  return g;
}
// Image::evaluate (image.hxx:76-90); out of line: every material parameter lookup lands here
DEVN F4v image_evaluate(const DImage& img, V2 in_uv, float* pdf) {
  ImageGather g = image_gather(img, in_uv);
  if (pdf) {
    bool flat = (img.options & kImageUniformSamplingTable) || (img.fsize_y == 1.0f);
    float s_t = flat ? 1.0f : tmax(0.0f, m_sin(kPi * saturatef(in_uv.y + 0.0f / img.fsize_y)));
    F4v top = g.p00 + g.p01;
    float t = luminance({top.x, top.y, top.z}) * s_t;
    float s_b = flat ? 1.0f : tmax(0.0f, m_sin(kPi * saturatef(in_uv.y + 1.0f / img.fsize_y)));
    F4v bot = g.p10 + g.p11;
    float b = luminance({bot.x, bot.y, bot.z}) * s_b;
    *pdf = (t + b) / img.normalization;
  }
  return g.p00 + g.p01 + g.p10 + g.p11;
}
DEV float image_evaluate_alpha(const DImage& img, V2 in_uv) { return image_evaluate(img, in_uv, nullptr).w; }

// Distribution::sample (distribution.hxx:16-35) over `count` entries
DEV uint32_t dist_sample(const etxb_distribution_entry* values, uint32_t count, float rnd) {
  uint32_t b = 0, e = count;
  do {
    uint32_t m = b + (e - b) / 2;
    if (__ldg(&values[m].cdf) >= rnd) {
      e = m;
    } else {
      b = m;
    }
  } while ((e - b) > 1);
  return b;
}

// Image::sample (image.hxx:119-150)
DEV V2 image_sample(const DImage& img, V2 rnd, float& image_pdf, F4v& eval) {
  uint32_t ny = img.isize_y + 1u, nx = img.isize_x + 1u;
  uint32_t ly = dist_sample(img.y_dist, ny, rnd.y);
  const etxb_distribution_entry* xd = img.x_dist + size_t(ly) * nx;
  uint32_t lx = dist_sample(xd, nx, rnd.x);
  float x0c = __ldg(&xd[lx].cdf), x1c = __ldg(&xd[umin(lx + 1u, nx - 1u)].cdf);
  float dx = (rnd.x - x0c);
  if (x1c - x0c > 0.0f) dx /= (x1c - x0c);
  float y0c = __ldg(&img.y_dist[ly].cdf), y1c = __ldg(&img.y_dist[umin(ly + 1u, ny - 1u)].cdf);
  float dy = (rnd.y - y0c);
  if (y1c - y0c > 0.0f) dy /= (y1c - y0c);
  V2 uv = {(float(lx) + dx) / img.fsize_x, (float(ly) + dy) / img.fsize_y};
  eval = image_evaluate(img, uv, &image_pdf);
  return uv;
}

}  // namespace etxb
