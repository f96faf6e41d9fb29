// K3 candidate filter — integer-only tests on the high words of delta that decide, per iteration,
// whether the exact glitch / escape comparisons of k3_checked.cuh COULD fire. k3_fast runs the delta
// recurrence alone (6 FP64 instructions per iteration: 2 DADD + 4 DFMA) and never forms z = Z + delta;
// a block in which a filter fires is rolled back and replayed by k3_events with the exact doubles, so
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
