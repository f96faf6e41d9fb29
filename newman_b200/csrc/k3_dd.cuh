// K3 (double-double form) — the "exact mode" refinement. FP64 perturbation carries a relative error of ~5e-15 in delta
// after a few thousand iterations; a sample whose last few hundred iterations are chaotic amplifies that past 1 and its
// escape count moves (0.5 % of cfg2's samples, on all of which the reference — whose own arbitrary
// precision continuation is converged there — is right). For the samples listed, phase 3 is repeated from K2's hand-over
// with delta, eps and the orbit in double-double (~106 bits):
//        w = 2 Z[j] + delta      delta' = delta * w + eps      z' = Z[j+1] + delta'
// escape test on the high parts of z', rebasing onto Z[0] = 0 when the sample outlives the orbit, no glitch rule (the
// cancellation a glitch stands for costs ~20 of 106 bits). Z[j] = x_hi[j-1] + x_lo[j-1]: the orbit as the host descended
// it, exact to 106 bits. One thread per sample, run to completion (the list is ~1 % of a frame). Every operation is a
// plain IEEE double operation or an explicit fma, in the order oracle/oracle_p.c: oraclep_refine_dd defines; the build
// has -fmad=false, so nothing is contracted.
#pragma once
#include "k3_checked.cuh"

namespace nm {

struct dd_t { double hi, lo; };
__device__ __forceinline__ dd_t dd_fast2sum(double a, double b) { dd_t r; r.hi = a + b; r.lo = b - (r.hi - a); return r; }
__device__ __forceinline__ dd_t dd_2sum(double a, double b) {
  dd_t r; r.hi = a + b; const double bb = r.hi - a; r.lo = (a - (r.hi - bb)) + (b - bb); return r;
}
__device__ __forceinline__ dd_t dd_add(dd_t x, dd_t y) {
  dd_t s = dd_2sum(x.hi, y.hi);
  const dd_t t = dd_2sum(x.lo, y.lo);
  s.lo += t.hi;
  s = dd_fast2sum(s.hi, s.lo);
  s.lo += t.lo;
  return dd_fast2sum(s.hi, s.lo);
}
__device__ __forceinline__ dd_t dd_mul(dd_t x, dd_t y) {
  dd_t p; p.hi = x.hi * y.hi; p.lo = __fma_rn(x.hi, y.hi, -p.hi);
  p.lo += x.hi * y.lo;
  p.lo += x.lo * y.hi;
  return dd_fast2sum(p.hi, p.lo);
}
__device__ __forceinline__ dd_t dd_neg(dd_t x) { x.hi = -x.hi; x.lo = -x.lo; return x; }
__device__ __forceinline__ dd_t dd_dbl(dd_t x) { x.hi *= 2.0; x.lo *= 2.0; return x; }   // exact

struct DDParams {
  const double2* Xhi;     // [M + has_escape] truncated doubles of X[i]
  const double2* Xlo;     // [M] their low parts
  int M, Jmax, N, nc;
  const double* eps_re;   // [nc], [nr]: the pixel offsets as doubles ...
  const double* eps_im;
  const double* eps_re_lo;  // ... and their low parts (nullptr: zero)
  const double* eps_im_lo;
  nm_escape* out;
  unsigned long long* ctr;
  FixupRec* fix;
  unsigned long long fix_cap;
  double log_bailout;
};

constexpr int K3_DD_THREADS = 128;

__global__ void __launch_bounds__(K3_DD_THREADS) k3_dd(DDParams p, FreshArrays f, long long n) {
  unsigned long long executed = 0, rebased = 0;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (long long)gridDim.x * blockDim.x) {
    int j = f.j[w];
    if (j < 0) continue;   // K2 finished this sample itself (phase 2, cardioid, L >= N)
    const int pix = f.pix[w];
    int off = f.off[w];
    const double2 d0 = f.d[w];
    dd_t dr = {d0.x, 0.0}, di = {d0.y, 0.0};
    const int r = pix / p.nc, c = pix - r * p.nc;
    const dd_t er = {p.eps_re[c], p.eps_re_lo ? p.eps_re_lo[c] : 0.0}, ei = {p.eps_im[r], p.eps_im_lo ? p.eps_im_lo[r] : 0.0};
    for (unsigned step = 0;; ++step) {
      if ((step & 1023u) == 1023u && ((volatile unsigned long long*)p.ctr)[CTR_CANCEL]) break;
      dd_t xr = {0.0, 0.0}, xi = {0.0, 0.0};
      if (j >= 1) {
        const double2 h = p.Xhi[j - 1];
        xr.hi = h.x; xi.hi = h.y;
        if (j - 1 < p.M) { const double2 l = p.Xlo[j - 1]; xr.lo = l.x; xi.lo = l.y; }
      }
      const dd_t wr = dd_add(dd_dbl(xr), dr), wi = dd_add(dd_dbl(xi), di);
      const dd_t ndr = dd_add(dd_add(dd_mul(dr, wr), dd_neg(dd_mul(di, wi))), er);
      const dd_t ndi = dd_add(dd_add(dd_mul(dr, wi), dd_mul(di, wr)), ei);
      dr = ndr; di = ndi;
      ++j;
      ++executed;
      const double2 yh = p.Xhi[j - 1];
      dd_t yr = {yh.x, 0.0}, yi = {yh.y, 0.0};
      if (j - 1 < p.M) { const double2 l = p.Xlo[j - 1]; yr.lo = l.x; yi.lo = l.y; }
      const dd_t zr = dd_add(yr, dr), zi = dd_add(yi, di);
      const double zmag = __fma_rn(zi.hi, zi.hi, zr.hi * zr.hi);
      if (zmag > BAILOUT2) {
        const double r2 = zr.hi * zr.hi + zi.hi * zi.hi;
        bool unc;
        const float sm = smoothing_f32(r2, p.log_bailout, &unc);
        nm_escape v; v.iterations = j + off; v.smoothing = sm;
        p.out[pix] = v;
        if (unc) push_fixup(p.ctr, p.fix, p.fix_cap, pix, r2);
        break;
      }
      if (j + off + 1 >= p.N) {
        nm_escape v; v.iterations = p.N; v.smoothing = 0.0f;
        p.out[pix] = v;
        break;
      }
      if (j == p.Jmax) { ++rebased; off = j + off; j = 0; dr = zr; di = zi; }
    }
  }
  for (int o = 16; o; o >>= 1) {
    executed += __shfl_xor_sync(FULL_MASK, executed, o);
    rebased += __shfl_xor_sync(FULL_MASK, rebased, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (executed) { atomicAdd(&p.ctr[CTR_EXECUTED], executed); atomicAdd(&p.ctr[CTR_CHECKED], executed); }
    if (rebased) atomicAdd(&p.ctr[CTR_REBASED], rebased);
  }
}

// Which samples of two rasters of the same frame carry different escape counts (the exact mode's sensitivity probe: the
// frame iterated against the orbit rounded to nearest and against the truncated one).
__global__ void k_raster_diff(const nm_escape* a, const nm_escape* b, long long n, int32_t* list, unsigned long long* count) {
  for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
    const long long i = base + threadIdx.x;      // (whole warps stay in the loop together: warp_reserve votes)
    const bool differ = i < n && a[i].iterations != b[i].iterations;
    const unsigned long long slot = warp_reserve(count, differ);
    if (differ) list[slot] = (int32_t)i;
  }
}

}  // namespace nm
