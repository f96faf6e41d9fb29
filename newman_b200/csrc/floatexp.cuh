// floatexp: double mantissa + int exponent ("a double with unbounded exponent range").
// Used by the series phase (K2) when the descended coefficients A/B/C no longer fit a double
// (|C| > 1.8e308 from pixel pitch ~ 1e-97 on, where the reference itself dies with SIGFPE —
// SURVEY.md finding 3) or eps^3 would underflow.
//
// Contract: every operation performs ONE IEEE double operation on exactly scaled operands, so
// whenever the same computation in plain double would neither overflow nor underflow, the floatexp
// result (converted back) is bit-identical to it. That is what lets the deep-range path be pinned
// against the double path (and through it against the compiled reference) on views both can handle.
// Mirrored line for line by oracle/oracle_p.c (fe_* functions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nm {

struct fe {
  double m;  // 0, or 1 <= |m| < 2
  int e;     // value = m * 2^e
};

__host__ __device__ __forceinline__ fe fe_norm(double m, int e) {
  fe r;
  if (m == 0.0 || m != m) { r.m = m; r.e = 0; return r; }
#ifdef __CUDA_ARCH__
  int hi = __double2hiint(m);
  int ex = (hi >> 20) & 0x7ff;
  if (ex == 0 || ex == 0x7ff) {  // denormal mantissa (cannot arise from scaled ops) or inf: use frexp
    int k;
    double f = frexp(m, &k);     // f in [0.5, 1)
    r.m = f * 2.0; r.e = e + k - 1;
    return r;
  }
  r.m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(m));
  r.e = e + ex - 1023;
#else
  int k;
  double f = frexp(m, &k);
  r.m = f * 2.0; r.e = e + k - 1;
#endif
  return r;
}

__host__ __device__ __forceinline__ fe fe_from_double(double x) { return fe_norm(x, 0); }
// mantissa/exponent pair as the host descends an mpf: x = m * 2^e with 0.5 <= |m| < 1 (mpf_get_d_2exp)
__host__ __device__ __forceinline__ fe fe_from_parts(double m, int e) { return fe_norm(m, e); }

__host__ __device__ __forceinline__ double fe_scale(double m, int k) {  // m * 2^k, k <= 0, exact unless it underflows
  if (k < -1080) return 0.0 * m;  // keeps the sign of zero like a gradual underflow to zero would
#ifdef __CUDA_ARCH__
  return scalbn(m, k);
#else
  return ldexp(m, k);
#endif
}

__host__ __device__ __forceinline__ double fe_to_double(fe a) {
  if (a.m == 0.0) return a.m;
  if (a.e > 1100) return a.m > 0 ? (double)INFINITY : -(double)INFINITY;
  if (a.e < -1200) return 0.0 * a.m;
#ifdef __CUDA_ARCH__
  return scalbn(a.m, a.e);
#else
  return ldexp(a.m, a.e);
#endif
}

__host__ __device__ __forceinline__ fe fe_mul(fe a, fe b) { return fe_norm(a.m * b.m, a.e + b.e); }
__host__ __device__ __forceinline__ fe fe_neg(fe a) { a.m = -a.m; return a; }

__host__ __device__ __forceinline__ fe fe_add(fe a, fe b) {
  if (a.m == 0.0) return b;
  if (b.m == 0.0) return a;
  if (a.e >= b.e) return fe_norm(a.m + fe_scale(b.m, b.e - a.e), a.e);
  return fe_norm(fe_scale(a.m, a.e - b.e) + b.m, b.e);
}
__host__ __device__ __forceinline__ fe fe_sub(fe a, fe b) { return fe_add(a, fe_neg(b)); }

// a < b for non-negative a, b
__host__ __device__ __forceinline__ bool fe_lt_nonneg(fe a, fe b) {
  if (b.m == 0.0) return false;
  if (a.m == 0.0) return true;
  if (a.e != b.e) return a.e < b.e;
  return a.m < b.m;
}

struct fec { fe re, im; };

// LPComplex operator* (complex.h:29-31): (ar*br - ai*bi, ar*bi + ai*br)
__host__ __device__ __forceinline__ fec fec_mul(fec a, fec b) {
  fec r;
  r.re = fe_sub(fe_mul(a.re, b.re), fe_mul(a.im, b.im));
  r.im = fe_add(fe_mul(a.re, b.im), fe_mul(a.im, b.re));
  return r;
}
// sq (complex.h:19-21): (re*re - im*im, 2.0*re*im)
__host__ __device__ __forceinline__ fec fec_sq(fec a) {
  fec r;
  r.re = fe_sub(fe_mul(a.re, a.re), fe_mul(a.im, a.im));
  fe two_re = a.re; two_re.e += (a.re.m != 0.0);  // 2.0 * re is exact
  r.im = fe_mul(two_re, a.im);
  return r;
}
__host__ __device__ __forceinline__ fec fec_add(fec a, fec b) {
  fec r; r.re = fe_add(a.re, b.re); r.im = fe_add(a.im, b.im); return r;
}
// sqMag (complex.h:23): re*re + im*im
__host__ __device__ __forceinline__ fe fec_sqmag(fec a) { return fe_add(fe_mul(a.re, a.re), fe_mul(a.im, a.im)); }

__host__ __device__ __forceinline__ double fe_log2_abs(fe a) {
  if (a.m == 0.0) return -(double)INFINITY;
  return log2(fabs(a.m)) + (double)a.e;
}

}  // namespace nm
