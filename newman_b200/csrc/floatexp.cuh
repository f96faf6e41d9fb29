// floatexp: double mantissa + int exponent ("a double with unbounded exponent range").
// Used by the series phase (K2) when the descended coefficients A/B/C no longer fit a double
// (|C| > 1.8e308 from pixel pitch ~ 1e-97 on, where the reference itself dies with SIGFPE —
// SURVEY.md finding 3) or eps^3 would underflow.
//
// Contract: every operation performs ONE IEEE double operation on exactly scaled operands, so
// whenever the same computation in plain double would neither overflow nor underflow, the floatexp
// result (converted back) is bit-identical to it. That is what lets the deep-range path be pinned
// against the double path (and through it against the compiled reference) on views both can handle.
// Restated by oracle/oracle_p.c (fe_* functions, pstate).
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

// 2^e as a double; flushed to zero below the normal range (no denormal scale factors anywhere)
__host__ __device__ __forceinline__ double pow2d(int e) {
  if (e < -1022) return 0.0;
  if (e > 1023) return (double)INFINITY;
#ifdef __CUDA_ARCH__
  return __hiloint2double((e + 1023) << 20, 0);
#else
  return ldexp(1.0, e);
#endif
}

// m * 2^k for k <= 0 and a normal m: exact; flushed to (signed) zero below 2^-1000 (far below half an
// ulp of anything it is added to, so sums are unaffected)
__host__ __device__ __forceinline__ double fe_scale(double m, int k) {
  if (k < -1000) return 0.0 * m;
  return m * pow2d(k);
}

// value as a double: exact inside the normal range, flushed to (signed) zero below it (no denormals)
__host__ __device__ __forceinline__ double fe_to_double(fe a) {
  if (a.m == 0.0) return a.m;
  if (a.e > 1023) return a.m > 0 ? (double)INFINITY : -(double)INFINITY;
  if (a.e < -1022) return 0.0 * a.m;
  return a.m * pow2d(a.e);
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

// ---- scaled perturbation state (restated by oracle/oracle_p.c: pstate) -----------------------------
// delta = (dr, di) * 2^e. e == 0: a plain state, iterated exactly as k3_perturb.cuh always did.
// e != 0: a "scaled" state with max(|dr|, |di|) in [1, 2) at every re-normalisation point; used while
// |delta| is below what a double product can hold (delta*delta underflows from |delta| < 1e-154 on,
// delta itself from 1e-308 on: views deeper than ~1e-150). One step is the same expression for both
// (S = 2^e; S == 1 for plain states, where every line is bit-identical to the plain kernel's):
//      w  = fma(S, d, 2*Z[j])        d' = fma(-+di, wi, fma(dr, wr, eps/2^e))        z = fma(S, d', Z[j+1])
// States are re-normalised when they are created and before the step from every index j = 0 (mod 64)
// (worst-case growth in between: 4^64 = 2^128; worst-case shrink: one passage of Z near zero).
constexpr int E_TO_PLAIN = -300;   // a normalised state with exponent above this becomes plain
constexpr int E_TO_SCALED = -400;  // a plain state whose larger component is below 2^this becomes scaled
constexpr int RENORM_MASK = 63;

struct pstate { double dr, di; int e; };

__host__ __device__ __forceinline__ pstate state_from_fec(fec d) {
  pstate s;
  if (d.re.m == 0.0 && d.im.m == 0.0) { s.dr = d.re.m; s.di = d.im.m; s.e = 0; return s; }
  const int E = d.re.m == 0.0 ? d.im.e : (d.im.m == 0.0 ? d.re.e : (d.re.e > d.im.e ? d.re.e : d.im.e));
  if (E > E_TO_PLAIN) { s.dr = fe_to_double(d.re); s.di = fe_to_double(d.im); s.e = 0; return s; }
  s.dr = d.re.m == 0.0 ? d.re.m : fe_scale(d.re.m, d.re.e - E);
  s.di = d.im.m == 0.0 ? d.im.m : fe_scale(d.im.m, d.im.e - E);
  s.e = E;
  return s;
}

__host__ __device__ __forceinline__ void state_renorm(pstate& s) {
  if (s.e == 0) {
    const double m = fmax(fabs(s.dr), fabs(s.di));
    if (!(m < pow2d(E_TO_SCALED)) || m == 0.0) return;
  }
  fec d; d.re = fe_norm(s.dr, s.e); d.im = fe_norm(s.di, s.e);
  s = state_from_fec(d);
}

// eps / 2^e as a double; eps0 = the value a plain state uses
__host__ __device__ __forceinline__ double eps_scaled(fe eps, double eps0, int e) {
  if (e == 0) return eps0;
  if (eps.m == 0.0) return eps.m;
  const int k = eps.e - e;
  if (k < -1000) return 0.0 * eps.m;
  if (k > 1000) return eps.m > 0 ? (double)INFINITY : -(double)INFINITY;
  return eps.m * pow2d(k);
}

}  // namespace nm
