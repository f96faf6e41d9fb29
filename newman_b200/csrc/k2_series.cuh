// K2 — series-approximation skip: phases 1 and 2 of Mandelbrot::getIterations
// (reference mandelbrot.cpp:144-207), one grid sample per thread.
//
// Bit-exactness by construction. Per sample the reference evaluates, for i = 1, 2, ...
//     b = B[i]*eps^2, c = C[i]*eps^3            (LPComplex operator*, complex.h:29-31)
//     unstable  <=>  |b|^2 * tol < |c|^2        (isUnstable, mandelbrot.cpp:138-142)
// with A/B/C descended to double by truncation (166-168), eps from the mpf subtraction (155-159),
// eps^2 = sq(eps), eps^3 = eps*eps^2 (160-161), all without FMA. Neither the test nor
// d[i] = (a+b)+c (180) depends on the previous index, so we run the test for i = 1.. until it first
// fires and then evaluate d[] in closed form only where the reference reads it: at found = L-1 and
// at the O(log L) probes of the phase-2 binary search (186-200). The ops and their order are the
// reference's; the build uses -fmad=false.
//
// Phase 2 forms X[found] + d[found] in mpf and truncates to double (184-186, 61). We carry X as
// hi + lo (hi = trunc53(X), lo = trunc53(X - hi)) and form trunc53(hi + lo + d) with error-free
// transforms and one round-toward-zero add; it can differ from the mpf value only if the exact sum
// lies within 2^-106 (relative) of a double, where the mpf digits below lo would decide.
#pragma once
#include "nm_common.cuh"

namespace nm {

struct K2Params {
  const double2* A;
  const double2* B;
  const double2* C;
  const double2* Z;     // Z[0]=0, Z[j]=X[j-1] (truncated doubles)
  const double2* Xlo;   // [M] low parts of X[i]
  int M, N;
  double tol;
  const double* eps_re;
  const double* eps_im;
  int nc;
  const int32_t* pix_list;  // nullptr: work index == pixel id
  long long W;              // number of work items
  int cardioid_mode;
  const uint8_t* mask;
  double2* init_d;          // [W] delta at j0
  int32_t* init_j;          // [W] table index j0 = L of the first K3 state, or -1 if finished here
  unsigned* hist;           // [K+1] fresh samples per orbit chunk
  int CH;
  nm_escape* out;
  unsigned long long* ctr;
  FixupRec* fix;
  unsigned long long fix_cap;
  double log_bailout;
};

constexpr int K2_THREADS = 256;

struct Cplx { double re, im; };

__device__ __forceinline__ Cplx cmul(double ar, double ai, double br, double bi) {
  Cplx r;
  r.re = ar * br - ai * bi;
  r.im = ar * bi + ai * br;
  return r;
}

// trunc53(hi + lo + d): see header comment.
__device__ __forceinline__ double trunc_add3(double hi, double lo, double d) {
  double s = hi + d;
  double bb = s - hi;
  double e = (hi - (s - bb)) + (d - bb);  // s + e == hi + d exactly
  double t = e + lo;
  double h = s + t;
  double b2 = h - s;
  double l = (s - (h - b2)) + (t - b2);   // h + l == s + t exactly
  return __dadd_rz(h, l);
}

struct SeriesEval {
  double er, ei, e2r, e2i, e3r, e3i;
  const double2 *A, *B, *C;
  __device__ __forceinline__ Cplx d_at(int j) const {
    Cplx r;
    if (j == 0) { r.re = er; r.im = ei; return r; }
    double2 a = A[j], b = B[j], c = C[j];
    Cplx ta = cmul(a.x, a.y, er, ei);
    Cplx tb = cmul(b.x, b.y, e2r, e2i);
    Cplx tc = cmul(c.x, c.y, e3r, e3i);
    r.re = (ta.re + tb.re) + tc.re;
    r.im = (ta.im + tb.im) + tc.im;
    return r;
  }
};

__global__ void __launch_bounds__(K2_THREADS) k2_series(K2Params p) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long evals = 0, skipped = 0;

  for (; w < p.W; w += stride) {
    int pix = p.pix_list ? p.pix_list[w] : (int)w;
    if (p.cardioid_mode == NM_CARDIOID_ALL || (p.cardioid_mode == NM_CARDIOID_MASK && p.mask[pix])) {
      p.out[pix].iterations = p.N;
      p.out[pix].smoothing = 0.0f;
      p.init_j[w] = -1;
      skipped++;
      continue;
    }
    int r = pix / p.nc, c = pix - r * p.nc;
    SeriesEval se;
    se.A = p.A; se.B = p.B; se.C = p.C;
    se.er = p.eps_re[c];
    se.ei = p.eps_im[r];
    // eps2 = sq(eps): (re*re - im*im, 2.0*re*im); eps3 = eps*eps2
    se.e2r = se.er * se.er - se.ei * se.ei;
    se.e2i = (2.0 * se.er) * se.ei;
    Cplx e3 = cmul(se.er, se.ei, se.e2r, se.e2i);
    se.e3r = e3.re; se.e3i = e3.im;

    // ---- phase 1: first unstable index ---------------------------------------------------------
    int L = p.M;
    for (int i = 1; i < p.M; ++i) {
      double2 b = p.B[i], cc = p.C[i];
      Cplx tb = cmul(b.x, b.y, se.e2r, se.e2i);
      Cplx tc = cmul(cc.x, cc.y, se.e3r, se.e3i);
      double bmag = tb.re * tb.re + tb.im * tb.im;
      double cmag = tc.re * tc.re + tc.im * tc.im;
      evals++;
      if (bmag * p.tol < cmag) {
        int good = i - 3;
        if (good < 1) good = 1;
        L = good;
        break;
      }
    }

    // ---- phase 2: did the series-approximated point already escape? ---------------------------
    int found = L - 1;
    Cplx d = se.d_at(found);
    double2 xh = p.Z[found + 1], xl = p.Xlo[found];
    double yr = trunc_add3(xh.x, xl.x, d.re);
    double yi = trunc_add3(xh.y, xl.y, d.im);
    double mag = yr * yr + yi * yi;
    if (mag > BAILOUT2) {
      int low = 0, high = L - 1, mid = L / 2;
      while (low <= high) {
        Cplx dm = se.d_at(mid);
        double2 mh = p.Z[mid + 1], ml = p.Xlo[mid];
        double mr = trunc_add3(mh.x, ml.x, dm.re);
        double mi = trunc_add3(mh.y, ml.y, dm.im);
        double mm = mr * mr + mi * mi;
        if (!(mm > BAILOUT2)) low = mid + 1;
        else { high = mid - 1; found = mid; }
        mid = (low + high) / 2;
      }
      d = se.d_at(found);
      xh = p.Z[found + 1]; xl = p.Xlo[found];
      yr = trunc_add3(xh.x, xl.x, d.re);
      yi = trunc_add3(xh.y, xl.y, d.im);
      mag = yr * yr + yi * yi;
      bool unc;
      float s = smoothing_f32(mag, p.log_bailout, &unc);
      p.out[pix].iterations = found;
      p.out[pix].smoothing = s;
      if (unc) push_fixup(p.ctr, p.fix, p.fix_cap, pix, mag);
      p.init_j[w] = -1;
      continue;
    }

    // ---- hand over to K3 ------------------------------------------------------------------------
    if (L >= p.N) {  // loop 'for (i = d.size(); i < N; i++)' (212) is empty
      p.out[pix].iterations = p.N;
      p.out[pix].smoothing = 0.0f;
      p.init_j[w] = -1;
      continue;
    }
    p.init_d[w] = make_double2(d.re, d.im);
    p.init_j[w] = L;  // state: delta paired with Z[L] = X[L-1]
    atomicAdd(&p.hist[L / p.CH], 1u);
  }

  for (int o = 16; o; o >>= 1) {
    evals += __shfl_xor_sync(FULL_MASK, evals, o);
    skipped += __shfl_xor_sync(FULL_MASK, skipped, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (evals) atomicAdd(&p.ctr[CTR_SERIES], evals);
    if (skipped) atomicAdd(&p.ctr[CTR_SKIPPED], skipped);
  }
}

// Exclusive scan of the per-chunk histogram (K+1 <= a few thousand entries): one CTA.
__global__ void k2_scan(const unsigned* hist, unsigned* offs, unsigned* cursor, int K1n) {
  __shared__ unsigned carry;
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    for (int i = 0; i < K1n; ++i) { offs[i] = acc; cursor[i] = acc; acc += hist[i]; }
    offs[K1n] = acc;
    carry = acc;
  }
  __syncthreads();
}

// Scatter fresh work indices into chunk-sorted order.
__global__ void __launch_bounds__(256) k2_scatter(const int32_t* init_j, long long W, int CH,
                                                  unsigned* cursor, int32_t* fresh_ids) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; w < W; w += stride) {
    int j = init_j[w];
    if (j >= 0) {
      unsigned slot = atomicAdd(&cursor[j / CH], 1u);
      fresh_ids[slot] = (int32_t)w;
    }
  }
}

}  // namespace nm
