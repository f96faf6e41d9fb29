// K1 — plain-double escape kernel.
// Replaces Mandelbrot::getIterationsHW (reference mandelbrot.cpp:231-254) + inCardioid (63-71)
// + getSmoothingMagnitude (133-136), one grid sample per lane, for the whole raster.
//
// Arithmetic contract (bit-exact escape counts): the reference iterates
//     Z = sq(Z) + Z0;  sq(a) = (a.re*a.re - a.im*a.im, 2.0*a.re*a.im)   (complex.h:19-21)
//     if (Z.re*Z.re + Z.im*Z.im > 2^20) break;                          (complex.h:23)
// compiled without FMA (reference makefile:3, x86-64 baseline). We issue exactly those IEEE ops in
// that order: t=rr-ii; zr'=t+cr; u=zr+zr (== 2.0*zr exactly); v=u*zi; zi'=v+ci; rr'=zr'^2;
// ii'=zi'^2; mag=rr'+ii'  -> 8 FP64 ops per iteration (rr', ii' are reused by the next iteration,
// the same CSE the host compiler does). The build uses -fmad=false so nothing is contracted.
//
// Scheduling: escape times differ by 1000x between neighbouring samples, so lanes are never tied
// to a fixed pixel. Every warp is persistent: idle lanes are found with a ballot, one lane reserves
// that many samples from a global cursor with a single atomicAdd, and each idle lane takes the
// sample at its rank (popc of the lower idle lanes). Lanes run in bursts of K1_BURST iterations
// between re-deals.
#pragma once
#include "nm_common.cuh"

namespace nm {

struct K1Params {
  const double* c_re;
  const double* c_im;
  int nr, nc, N;
  nm_escape* out;
  unsigned long long* ctr;
  int32_t* ambig;
  unsigned long long ambig_cap;
  FixupRec* fix;
  unsigned long long fix_cap;
  double log_bailout;
};

constexpr int K1_THREADS = 256;
constexpr int K1_BURST = 64;

// Cardioid / period-2 bulb test in double (reference does it in mpf, mandelbrot.cpp:63-71).
// Returns 1 inside, 0 outside, 2 when the double evaluation cannot decide: the margin is within the
// first-order error bound of (a) truncating the mpf coordinate to double and (b) double rounding.
__device__ __forceinline__ int cardioid_class(double x, double y) {
  const double K = 1.5e-14;  // 64 x 2^-52: generous multiple of the unit round-off
  double xmf = x - 0.25;
  double y2 = y * y;
  double q = xmf * xmf + y2;
  double lhs = q * (q + xmf);
  double rhs = 0.25 * y2;
  double gx = 2.0 * xmf * (2.0 * q + xmf) + q;
  double gy = 2.0 * y * (2.0 * q + xmf) - 0.5 * y;
  double thr = K * (fabs(gx) * fabs(x) + fabs(gy) * fabs(y) + fabs(lhs) + fabs(rhs));
  double m1 = lhs - rhs;
  if (fabs(m1) <= thr) return 2;
  if (m1 < 0.0) return 1;
  double q2 = x + 1.0;
  double t = q2 * q2 + y2;
  double thr2 = K * (fabs(2.0 * q2) * fabs(x) + 2.0 * y2 + t + 0.0625);
  double m2 = t - 0.0625;
  if (fabs(m2) <= thr2) return 2;
  return m2 < 0.0 ? 1 : 0;
}

__global__ void __launch_bounds__(K1_THREADS) k1_escape(K1Params p) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long total = (long long)p.nr * p.nc;

  bool active = false;
  bool drained = false;
  double zr = 0, zi = 0, cr = 0, ci = 0, rr = 0, ii = 0;
  int it = 0, pix = 0;
  unsigned long long executed = 0, skipped = 0;

  for (;;) {
    // ---- re-deal: hand fresh samples to idle lanes -------------------------------------------
    while (!drained) {
      unsigned idle = __ballot_sync(FULL_MASK, !active);
      if (!idle) break;
      unsigned long long base = 0;
      if (lane == 0) {
        if (((volatile unsigned long long*)p.ctr)[CTR_CANCEL]) base = (unsigned long long)total;
        else base = atomicAdd(&p.ctr[CTR_NEXT], (unsigned long long)__popc(idle));
      }
      base = __shfl_sync(FULL_MASK, base, 0);
      if ((long long)base >= total) { drained = true; break; }
      if (!active) {
        long long idx = (long long)base + __popc(idle & lt_mask);
        if (idx < total) {
          pix = (int)idx;
          int r = pix / p.nc, c = pix - r * p.nc;
          cr = p.c_re[c];
          ci = p.c_im[r];
          int cls = cardioid_class(cr, ci);
          if (cls == 1) {
            p.out[pix].iterations = p.N;
            p.out[pix].smoothing = 0.0f;
            skipped++;
          } else {
            if (cls == 2) {
              unsigned long long k = atomicAdd(&p.ctr[CTR_AMBIG], 1ULL);
              if (k < p.ambig_cap) p.ambig[k] = pix;
            }
            zr = cr; zi = ci;
            rr = zr * zr; ii = zi * zi;
            it = 0;
            active = true;
          }
        }
      }
    }
    if (!__any_sync(FULL_MASK, active)) break;

    // ---- burst of iterations -------------------------------------------------------------------
    if (active) {
      int n = p.N - it;
      if (n > K1_BURST) n = K1_BURST;
      int k = 0;
      bool esc = false;
      double mag = 0.0;
#pragma unroll 4
      for (; k < n; ++k) {
        double t = rr - ii;
        double nzr = t + cr;
        double u = zr + zr;
        double v = u * zi;
        double nzi = v + ci;
        zr = nzr; zi = nzi;
        rr = zr * zr; ii = zi * zi;
        mag = rr + ii;
        if (mag > BAILOUT2) { esc = true; break; }
      }
      it += k;
      executed += (unsigned long long)k + (esc ? 1ULL : 0ULL);
      if (esc) {
        bool unc;
        float s = smoothing_f32(mag, p.log_bailout, &unc);
        p.out[pix].iterations = it;
        p.out[pix].smoothing = s;
        if (unc) push_fixup(p.ctr, p.fix, p.fix_cap, pix, mag);
        active = false;
      } else if (it >= p.N) {
        p.out[pix].iterations = p.N;
        p.out[pix].smoothing = 0.0f;
        active = false;
      }
    }
  }

  // warp-reduce the statistics, one atomic per warp
  for (int o = 16; o; o >>= 1) {
    executed += __shfl_xor_sync(FULL_MASK, executed, o);
    skipped += __shfl_xor_sync(FULL_MASK, skipped, o);
  }
  if (lane == 0) {
    if (executed) atomicAdd(&p.ctr[CTR_EXECUTED], executed);
    if (skipped) atomicAdd(&p.ctr[CTR_SKIPPED], skipped);
  }
}

}  // namespace nm
