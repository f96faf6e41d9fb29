// K3 candidate filter — integer-only tests on the high words of delta that decide, per iteration,
// whether the exact glitch / escape comparisons of k3_checked.cuh COULD fire. k3_fast runs the delta
// recurrence alone (6 FP64 instructions per iteration: 2 DADD + 4 DFMA) and never forms z = Z + delta;
// a block in which a filter fires is rolled back and replayed by k3_finish<.., EVENTS> with the exact doubles, so
// the decisions (and therefore the raster) are those of k3_perturb.cuh / the oracle. What has to hold is
// only that the filters have NO FALSE NEGATIVES; tests/test_k3_filter.py attacks exactly that on the
// host build of the functions below (nm_k3_filter_entry / nm_k3_filter_fires).
//
// Glitch (k3_checked.cuh: zmag = fma(zi, zi, zr*zr) < gb[j], zr = fl(Zr + dr), zi = fl(Zi + di)):
//   fl(zr*zr) <= zmag (the fma adds a non-negative term and rounding is monotone), and likewise
//   fl(zi*zi) <= zmag, so a glitch implies |zr|, |zi| < sqrt(gb)(1 + 2^-52), hence for the real sums
//   |Zc + dc| < g := sqrt(gb)(1 + 2^-30) in both components c (fl(x + y) is within 2^-53 relative of
//   x + y; gb below 2^-900, where squares could underflow, is treated as "always a candidate").
//   Where |Zc| > g the interval (-Zc - g, -Zc + g) does not contain 0: dc has the sign of -Zc and
//   |dc| lies in [(|Zc| - g)(1 - 2^-30), (|Zc| + g)(1 + 2^-30)]. Same-sign doubles order like their high
//   words, so   (uint32)(hi(dc) - lo_c) <= w_c   with lo_c = sign | hi(lower bound), w_c = hi(upper) -
//   hi(lower) is implied. A component with |Zc| <= g(1 + 2^-10) is not tested (lo = 0, w = 2^32 - 1).
//   The test is the AND of both components: false alarms need delta within ~sqrt(glitch_tol)|Z| of -Z
//   in a square instead of a disc, i.e. 4/pi times the true glitch rate.
//   gb == 0 (index 0, the escaped iterate, table padding, underflow) can never satisfy zmag < gb:
//   lo = 2^32 - 1, w = 0 (only a NaN high word equals it, and a NaN is no glitch either).
// Scaled states (floatexp.cuh: delta = d * 2^e, e <= -300 at a re-normalisation point, at most 64 steps
//   and a factor (2|Z| + |delta|) <= 4.01 per step later, so |delta| < 2^-170 while |Z| <= 2, and the few
//   orbit entries with |Z| > 2 right before the reference escapes have |Zc| > 1): their d is not delta,
//   so the kernel replaces their difference word by all-ones — they are candidates exactly at the
//   entries that test nothing. Entries with max|Zc| < 2^-100 are made such entries.
// Escape (zmag > 2^20), tested once per block of 4 on the block's last iterate like before (|z| > 1024
//   keeps growing and cannot overflow within 3 more steps): zmag <= (zr^2 + zi^2)(1 + 2^-52)^2, so
//   |Z + delta| > 1024(1 - 2^-51) and max(|dr|, |di|) > (1023.99 - |Z|) * 0.7071067 =: room. Candidate
//   iff hi(|dr|) >= hi(room) or hi(|di|) >= hi(room); room <= 0 (the escaped iterate) => always. Scaled
//   states (|delta| < 2^-127 everywhere) are candidates only where room <= 0.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace nm {

struct __align__(16) K3Filt { uint32_t lo_r, w_r, lo_i, w_i; };

__host__ __device__ __forceinline__ uint32_t k3f_hi32(double x) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__double2hiint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u >> 32);
#endif
}

__host__ __device__ __forceinline__ void k3f_component(double zc, double g, uint32_t* lo, uint32_t* w) {
  const double a = fabs(zc);
  if (a > g * 1.0009765625) {
    const double lo_mag = (a - g) * (1.0 - 9.313225746154785e-10);  // 2^-30
    const double hi_mag = (a + g) * (1.0 + 9.313225746154785e-10);
    const uint32_t l = k3f_hi32(lo_mag), h = k3f_hi32(hi_mag);
    *lo = (zc > 0.0 ? 0x80000000u : 0u) | l;   // delta_c has the sign of -Z_c
    *w = h - l;
  } else {
    *lo = 0u; *w = 0xffffffffu;
  }
}

// Table entry for index j: (zr, zi) = Z[j], gb = the glitch bound the exact test uses at j.
__host__ __device__ __forceinline__ void k3_filter_entry(double zr, double zi, double gb, K3Filt* f, int32_t* esc_hi) {
  const double az = sqrt(zr * zr + zi * zi) * (1.0 + 9.313225746154785e-10);
  const double room = (1023.99 - az) * 0.7071067;
  *esc_hi = room > 0.0 ? (int32_t)k3f_hi32(room) : 0;   // NaN / inf |Z| => 0 (always a candidate)
  if (!(gb > 0.0)) { f->lo_r = f->lo_i = 0xffffffffu; f->w_r = f->w_i = 0u; return; }
  const double amax = fmax(fabs(zr), fabs(zi));
  if (gb < 1.1832913578315177e-271 /* 2^-900 */ || !(amax >= 7.888609052210118e-31 /* 2^-100 */)) {
    f->lo_r = f->lo_i = 0u; f->w_r = f->w_i = 0xffffffffu; return;
  }
  const double g = sqrt(gb) * (1.0 + 9.313225746154785e-10);
  k3f_component(zr, g, &f->lo_r, &f->w_r);
  k3f_component(zi, g, &f->lo_i, &f->w_i);
}

// Quiet segments. k3_fast advances in segments of 16 iterations from indices j0 = 0 (mod 16). Whether the glitch
// test can fire ANYWHERE in a segment is decided once, at its start, from the size of delta alone:
//   the kernel's step  w = fl(2Z_j + delta), delta' = (fma(-di, wi, fma(dr, wr, er)), fma(di, wr, fma(dr, wi, ei)))
//   is the complex delta*w + eps with |w| <= (2|Z_j| + |delta|)(1 + u) and a componentwise error of at most
//   u|result| + u(1 + u)(|delta||w| + |eps|)  (u = 2^-53), so in 2-norms
//       |delta'| <= (|delta| (2|Z_j| + |delta|) + |eps|)(1 + 2^-40)  (+ 1e-290 for results that underflow).
//   A glitch at index k implies |Z_c + delta_c| < g_k = sqrt(gb_k)(1 + 2^-30) in both components (above), hence
//   |delta_k| > |Z_k| - sqrt(2) g_k =: L_k. So with D_0 >= |delta_j0| and D_{i+1} = the bound above, no glitch is
//   possible in the segment if D_i <= L_{j0+i} for i = 1..16. The bound is monotone in D_0; k3_seg_bound searches the
//   largest high word T such that every state with hi(|dr|) < T and hi(|di|) < T (=> |delta| < sqrt(2) double(T, 0))
//   and every |eps| <= e_max (the frame's largest pixel offset) passes. The kernel then runs such a segment without the
//   per-iteration filter (6 FP64 instructions per iteration and nothing else). |delta w| = |delta||w| is exact for
//   complex numbers, so the bound loses only the sqrt(2) of the component test: a segment is "loud" for a state only
//   where its |delta| really comes within a small factor of some |Z_k| (an approach of the reference to 0, or the
//   state's last iterations before it escapes).
//   T = 0 (never quiet): a segment that reaches beyond the table, contains an entry the filter treats as "always a
//   candidate" (gb < 2^-900 or max|Z_c| < 2^-100), or has L_k < 2^-160 somewhere. Scaled states (|delta| < 2^-170,
//   see above) are therefore quiet exactly where T > 0; the kernel masks their high words to 0.
//   gb_k == 0 (never a glitch) puts no constraint on D_k.
__host__ __device__ __forceinline__ double k3f_from_hi(uint32_t hi) {
  const uint64_t u = (uint64_t)hi << 32;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}

// Z[j] = (zr[j], zi[j]) and gb[j] for j <= jmax (the last valid table index); e_max >= |eps| of every sample.
// NS = steps the bound covers (16: k3_fast's segments).
template <int NS>
__host__ __device__ inline int32_t k3_seg_bound_n(const double* zr, const double* zi, const double* gb, int zstride, int j0, int jmax,
                                                  double e_max) {
  if (j0 + NS > jmax || !(e_max >= 0.0)) return 0;
  const double up = 1.0 + 9.094947017729282e-13;   // 2^-40
  double g2[NS], L[NS];                              // 2|Z_{j0+i}| (rounded up), L_{j0+i+1}
  for (int i = 0; i < NS; ++i) {
    const double ar = zr[(size_t)(j0 + i) * zstride], ai = zi[(size_t)(j0 + i) * zstride];
    g2[i] = 2.0 * sqrt(ar * ar + ai * ai) * up;
    if (!(g2[i] < 1e300)) return 0;
    const int k = j0 + i + 1;
    const double kr = zr[(size_t)k * zstride], ki = zi[(size_t)k * zstride], b = gb[k];
    if (!(b > 0.0)) { L[i] = 1e308; continue; }     // never a glitch at k
    if (b < 1.1832913578315177e-271 || !(fmax(fabs(kr), fabs(ki)) >= 7.888609052210118e-31)) return 0;
    const double l = sqrt(kr * kr + ki * ki) * (1.0 - 9.094947017729282e-13) - 1.4142135623730951 * sqrt(b) * (1.0 + 1.862645149230957e-09);
    if (!(l >= 6.842277657836021e-49 /* 2^-160 */)) return 0;
    L[i] = l;
  }
  uint32_t lo = 0u, hi = 0x7fe00001u;               // lo passes (or is the sentinel 0), hi fails
  while (hi - lo > 1u) {
    const uint32_t mid = lo + (hi - lo) / 2u;
    double D = 1.4142135623730951 * k3f_from_hi(mid) * up;
    bool ok = true;
    for (int i = 0; i < NS && ok; ++i) {
      D = (D * (g2[i] + D) + e_max) * up + 1e-290;
      ok = D <= L[i];                                // false for NaN / inf as well
    }
    if (ok) lo = mid; else hi = mid;
  }
  return (int32_t)lo;
}
__host__ __device__ inline int32_t k3_seg_bound(const double* zr, const double* zi, const double* gb, int zstride, int j0, int jmax,
                                                double e_max) {
  return k3_seg_bound_n<16>(zr, zi, gb, zstride, j0, jmax, e_max);
}

// The tests as k3_fast evaluates them (scaled_mask: 0 for a plain state, 0xffffffff for a scaled one).
__host__ __device__ __forceinline__ bool k3_filter_glitch(const K3Filt& f, double dr, double di, uint32_t scaled_mask) {
  const uint32_t a = (k3f_hi32(dr) - f.lo_r) | scaled_mask;
  const uint32_t b = (k3f_hi32(di) - f.lo_i) | scaled_mask;
  return (a <= f.w_r) & (b <= f.w_i);
}
__host__ __device__ __forceinline__ bool k3_filter_escape(int32_t esc_hi, double dr, double di, uint32_t scaled_mask) {
  const int32_t keep = (int32_t)(0x7fffffffu & ~scaled_mask);
  return (((int32_t)k3f_hi32(dr) & keep) >= esc_hi) | (((int32_t)k3f_hi32(di) & keep) >= esc_hi);
}

}  // namespace nm
