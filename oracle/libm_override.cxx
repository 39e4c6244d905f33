// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product library.
//
// Re-exports the portable transcendentals (etx_tracer_b200/csrc/portable_math.h) under the libm names.
// The oracle shared object is linked with -Wl,-Bsymbolic, so every sinf/powf/... call made by the
// reference's headers (compiled in place from /root/reference/sources) binds to these definitions instead
// of glibc's.  The parity build of the CUDA module calls the same pm:: functions on the device, which is
// what makes sampler-state and film comparisons bit-exact.  Exact libm functions (sqrtf, floorf, fmodf,
// nextafterf, ...) are left to glibc: IEEE-754 fixes their results.
#include <complex.h>

#include "portable_math.h"

extern "C" {

float sinf(float x) { return pm::sinf_(x); }
float cosf(float x) { return pm::cosf_(x); }
void sincosf(float x, float* s, float* c) {
  *s = pm::sinf_(x);
  *c = pm::cosf_(x);
}
float tanf(float x) { return pm::tanf_(x); }
float acosf(float x) { return pm::acosf_(x); }
float asinf(float x) { return pm::asinf_(x); }
float atanf(float x) { return pm::atanf_(x); }
float atan2f(float y, float x) { return pm::atan2f_(y, x); }
float expf(float x) { return pm::expf_(x); }
float logf(float x) { return pm::logf_(x); }
float powf(float x, float y) { return pm::powf_(x, y); }
float coshf(float x) { return pm::coshf_(x); }
float sinhf(float x) { return pm::sinhf_(x); }
float tanhf(float x) { return pm::tanhf_(x); }
float atanhf(float x) { return pm::atanhf_(x); }

float _Complex csqrtf(float _Complex z) {
  float re, im;
  pm::csqrtf_(__real__ z, __imag__ z, re, im);
  float _Complex r;
  __real__ r = re;
  __imag__ r = im;
  return r;
}
float _Complex cexpf(float _Complex z) {
  float re, im;
  pm::cexpf_(__real__ z, __imag__ z, re, im);
  float _Complex r;
  __real__ r = re;
  __imag__ r = im;
  return r;
}
float cabsf(float _Complex z) { return pm::cabsf_(__real__ z, __imag__ z); }

// libgcc's float-complex division helper; defined here so the result does not depend on the libgcc build.
float _Complex __divsc3(float a, float b, float c, float d) {
  float re, im;
  pm::cdivf_(a, b, c, d, re, im);
  float _Complex r;
  __real__ r = re;
  __imag__ r = im;
  return r;
}

}  // extern "C"
