// K2 — series-approximation skip: phases 1 and 2 of Mandelbrot::getIterations
// (reference mandelbrot.cpp:144-207), one grid sample per thread.
//
// Bit-exactness by construction. Per sample the reference evaluates, for i = 1, 2, ...
//     b = B[i]*eps^2, c = C[i]*eps^3            (LPComplex operator*, complex.h:29-31)
//     unstable  <=>  |b|^2 * tol < |c|^2        (isUnstable, mandelbrot.cpp:138-142)
// with A/B/C descended to double by truncation (166-168), eps from the mpf subtraction (155-159),
// eps^2 = sq(eps), eps^3 = eps*eps^2 (160-161), all without FMA. Neither the test nor
// d[i] = (a+b)+c (180) depends on the previous index, so only the FIRST index at which the test
// fires matters, and d[] is needed only where the reference reads it: at found = L-1 and at the
// O(log L) probes of the phase-2 binary search (186-200). Every test that is evaluated uses the
// reference's ops in the reference's order; the build uses -fmad=false.
//
// Two ways to find the first firing index (template LITERAL):
//   LITERAL   the reference's scan: test i = 1, 2, ... (19 FP64 instructions per index per sample —
//             twice a perturbation iteration; this would make the "skip" the dominant cost)
//   filtered  a pixel-independent per-index filter says which indices CAN fire for a sample with
//             log2|eps| = le; only those are tested, in ascending order, with the exact test.
//             The filter has no false negatives (derivation below), so L is the same integer.
//
// Filter. With p = log2|B_i| + 2 le (= log2|b|), q = log2|C_i| + 3 le (= log2|c|), lt = log2 tol:
//   (1) q <= -545: both components of c are below 2^-539, their squares round to +0, |c|^2 == 0 and
//       `x < 0` is false for every x >= 0: the test cannot fire.
//   (2) else if |b|^2*tol is a normal double with room to spare (2p + min(lt,0) >= -990) and nothing
//       overflows (2p + max(lt,0) <= 1000, 2q <= 1000): every intermediate that matters is normal, so
//       computed |b|^2*tol and |c|^2 carry relative errors of a few dozen ulp (or |c|^2 < 2^-999 is
//       far below |b|^2*tol and the test is false). Then "fires" implies
//       tol*|B_i|^2/|C_i|^2 < |eps|^2 (1 + 1e-14), i.e.   rlog_i := lt/2 + log2|B_i| - log2|C_i| <= le + ETA
//       with ETA = 1e-6 >> 1e-14: index i is a *candidate*.
//   (3) otherwise (denormal |b|^2*tol, or overflow) anything can happen in floating point: *unsafe*,
//       always tested.
// Samples with |eps| < 2^-333 (eps^3 denormal: beyond the reference's own depth limit) or non-finite
// le use the literal scan. Prefix minima/maxima of the per-index thresholds make "first index that
// must be tested" three binary searches; from there the scan walks forward testing only must-test
// indices. On every known-answer view the first candidate fires (1.00 exact tests per sample).
//
// Phase 2 forms X[found] + d[found] in mpf and truncates to double (184-186, 61). We carry X as
// hi + lo (hi = trunc53(X), lo = trunc53(X - hi)) and form trunc53(hi + lo + d) with error-free
// transforms and one round-toward-zero add; it can differ from the mpf value only if the exact sum
// lies within 2^-106 (relative) of a double, where the mpf digits below lo would decide.
#pragma once
#include "k3_checked.cuh"

namespace nm {

struct K2Filter {  // per-index thresholds in units of log2|eps| (length M each)
  double* rlog;    // candidate iff rlog[i] <= le + ETA
  double* a;       // |c|^2 can be nonzero iff le > a[i]
  double* b;       // |b|^2*tol may be denormal iff le < b[i]
  double* ov;      // something may overflow iff le > ov[i]
  double* pmin_rlog;
  double* pmin_ov;
  double* pmin_a;
  double* pmax_b;
};

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
  FreshArrays fresh;        // [W] K3 start state of work item w (j = -1 if the sample finished here)
  int align4;               // 1: take the <= 3 checked steps that make j a multiple of 4 (k3_fast)
  CheckedParams ck;         // tables/lists for those steps
  unsigned* hist;           // [Jmax+2] fresh samples per exact start index L
  nm_escape* out;
  unsigned long long* ctr;
  FixupRec* fix;
  unsigned long long fix_cap;
  double log_bailout;
  K2Filter f;
};

constexpr int K2_THREADS = 256;
constexpr double K2_ETA = 1e-6;
constexpr double K2_LE_MIN = -333.0;

struct Cplx { double re, im; };

__device__ __forceinline__ Cplx cmul(double ar, double ai, double br, double bi) {
  Cplx r;
  r.re = ar * br - ai * bi;
  r.im = ar * bi + ai * br;
  return r;
}

// log2 of the modulus without overflow/underflow of the squares; -inf for 0.
__device__ __forceinline__ double log2_abs(double x, double y) {
  double ax = fabs(x), ay = fabs(y);
  double m = fmax(ax, ay), n = fmin(ax, ay);
  if (m == 0.0) return -INFINITY;
  double r = n / m;
  return log2(m) + 0.5 * log2(1.0 + r * r);
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
  // the reference's isUnstable at index i (mandelbrot.cpp:166-173, 138-142)
  __device__ __forceinline__ bool unstable(int i, double tol) const {
    double2 b = B[i], c = C[i];
    Cplx tb = cmul(b.x, b.y, e2r, e2i);
    Cplx tc = cmul(c.x, c.y, e3r, e3i);
    double bmag = tb.re * tb.re + tb.im * tb.im;
    double cmag = tc.re * tc.re + tc.im * tc.im;
    return bmag * tol < cmag;
  }
};

// first i in [1, M) with arr[i] <= key (arr non-increasing), else M
__device__ __forceinline__ int first_le(const double* arr, int M, double key) {
  int lo = 1, hi = M;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (arr[mid] <= key) hi = mid; else lo = mid + 1;
  }
  return lo;
}
__device__ __forceinline__ int first_lt(const double* arr, int M, double key) {
  int lo = 1, hi = M;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (arr[mid] < key) hi = mid; else lo = mid + 1;
  }
  return lo;
}
// first i in [1, M) with arr[i] > key (arr non-decreasing), else M
__device__ __forceinline__ int first_gt(const double* arr, int M, double key) {
  int lo = 1, hi = M;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (arr[mid] > key) hi = mid; else lo = mid + 1;
  }
  return lo;
}

template <bool LITERAL>
__global__ void __launch_bounds__(K2_THREADS) k2_series(K2Params p) {
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long evals = 0, skipped = 0, aligned_steps = 0;

  // warp-uniform trip count: the tail of the loop body is warp-aggregated
  for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < p.W; base += stride) {
    const long long w = base + lane;
    const bool valid = w < p.W;
    int handoff_L = -1;  // >= 0: this sample continues in K3 from table index L

    if (valid) {
      int pix = p.pix_list ? p.pix_list[w] : (int)w;
      if (p.cardioid_mode == NM_CARDIOID_ALL || (p.cardioid_mode == NM_CARDIOID_MASK && p.mask[pix])) {
        p.out[pix].iterations = p.N;
        p.out[pix].smoothing = 0.0f;
        p.fresh.j[w] = -1;
        skipped++;
      } else {
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

        // ---- phase 1: first unstable index --------------------------------------------------------
        int first = p.M;  // index at which the test first fires (M: never)
        bool literal = LITERAL;
        double le = 0.0;
        if (!LITERAL) {
          if (se.er == 0.0 && se.ei == 0.0) {
            literal = false;  // b = c = 0 at every index: `0*tol < 0` never fires
          } else {
            le = log2_abs(se.er, se.ei);
            literal = !(le >= K2_LE_MIN) || !isfinite(le);
            if (!literal) {
              int i0 = first_le(p.f.pmin_rlog, p.M, le + K2_ETA);
              int i1 = first_lt(p.f.pmin_ov, p.M, le);
              int ia = first_lt(p.f.pmin_a, p.M, le), ib = first_gt(p.f.pmax_b, p.M, le);
              int i2 = ia > ib ? ia : ib;
              int i = i0 < i1 ? i0 : i1;
              if (i2 < i) i = i2;
              for (; i < p.M; ++i) {
                bool must = (p.f.rlog[i] <= le + K2_ETA) || (le > p.f.a[i] && le < p.f.b[i]) || (le > p.f.ov[i]);
                if (must) {
                  evals++;
                  if (se.unstable(i, p.tol)) { first = i; break; }
                }
              }
            }
          }
        }
        if (literal) {
          for (int i = 1; i < p.M; ++i) {
            evals++;
            if (se.unstable(i, p.tol)) { first = i; break; }
          }
        }
        int L = p.M;
        if (first < p.M) {  // d.resize(max(i - 3, 1)), mandelbrot.cpp:173-177
          L = first - 3;
          if (L < 1) L = 1;
        }

        // ---- phase 2: did the series-approximated point already escape? -------------------------
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
          p.fresh.j[w] = -1;
        } else if (L >= p.N) {  // loop 'for (i = d.size(); i < N; i++)' (212) is empty
          p.out[pix].iterations = p.N;
          p.out[pix].smoothing = 0.0f;
          p.fresh.j[w] = -1;
        } else {
          // ---- hand over to K3: delta paired with Z[L] = X[L-1] ---------------------------------
          double dr = d.re, di = d.im;
          int j = L, off = -1;
          bool cont = true;
          if (p.align4) {  // k3_fast works on indices that are multiples of 4: take up to 3 exact steps here
            int steps = 0;
            cont = advance_checked(p.ck, pix, se.er, se.ei, off, dr, di, j, (4 - (L & 3)) & 3, &steps);
            aligned_steps += (unsigned long long)steps;
          }
          if (cont) {
            p.fresh.d[w] = make_double2(dr, di);
            p.fresh.j[w] = j;
            p.fresh.off[w] = off;
            p.fresh.pix[w] = pix;
            handoff_L = j;
          } else {
            p.fresh.j[w] = -1;
          }
        }
      }
    }
    // warp-aggregated histogram of the start index (neighbouring samples share L almost always)
    {
      int bin = handoff_L;
      unsigned peers = __match_any_sync(FULL_MASK, bin);
      if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(&p.hist[bin], (unsigned)__popc(peers));
    }
  }

  for (int o = 16; o; o >>= 1) {
    evals += __shfl_xor_sync(FULL_MASK, evals, o);
    skipped += __shfl_xor_sync(FULL_MASK, skipped, o);
    aligned_steps += __shfl_xor_sync(FULL_MASK, aligned_steps, o);
  }
  if (lane == 0) {
    if (aligned_steps) atomicAdd(&p.ctr[CTR_EXECUTED], aligned_steps);
    if (evals) atomicAdd(&p.ctr[CTR_SERIES], evals);
    if (skipped) atomicAdd(&p.ctr[CTR_SKIPPED], skipped);
  }
}

// Per-index filter thresholds (pixel-independent, O(M) once per frame): any grid.
__global__ void __launch_bounds__(256) k2_prepare(const double2* B, const double2* C, int M, double tol, K2Filter f) {
  const double lt = log2(tol);
  const double lt_neg = fmin(lt, 0.0), lt_pos = fmax(lt, 0.0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    double rlog = INFINITY, a = INFINITY, b = -INFINITY, ov = INFINITY;
    if (i >= 1) {
      double2 Bi = B[i], Ci = C[i];
      double lb = log2_abs(Bi.x, Bi.y), lc = log2_abs(Ci.x, Ci.y);
      if (lc != -INFINITY) {            // C_i == 0: |c|^2 == 0, the test cannot fire
        rlog = (lb == -INFINITY) ? -INFINITY : 0.5 * lt + lb - lc;
        a = (-545.0 - lc) / 3.0;
        b = (lb == -INFINITY) ? INFINITY : (-990.0 - lt_neg - 2.0 * lb) / 4.0;
        double c1 = (lb == -INFINITY) ? INFINITY : (1000.0 - lt_pos - 2.0 * lb) / 4.0;
        double c2 = (1000.0 - 2.0 * lc) / 6.0;
        ov = fmin(c1, c2);
      }
    }
    f.rlog[i] = rlog; f.a[i] = a; f.b[i] = b; f.ov[i] = ov;
  }
}

// Their prefix minima/maxima: one CTA, each thread owns a contiguous segment, segment totals are
// combined with a warp-shuffle scan.
__global__ void __launch_bounds__(1024) k2_prefix(int M, K2Filter f) {
  __shared__ double s_rlog[32], s_ov[32], s_a[32], s_b[32];
  const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, wid = t >> 5;
  const int seg = (M + nt - 1) / nt;
  const int i0 = min(M, t * seg), i1 = min(M, i0 + seg);
  double m_rlog = INFINITY, m_ov = INFINITY, m_a = INFINITY, m_b = -INFINITY;
  for (int i = i0; i < i1; ++i) {
    m_rlog = fmin(m_rlog, f.rlog[i]); m_ov = fmin(m_ov, f.ov[i]); m_a = fmin(m_a, f.a[i]); m_b = fmax(m_b, f.b[i]);
    f.pmin_rlog[i] = m_rlog; f.pmin_ov[i] = m_ov; f.pmin_a[i] = m_a; f.pmax_b[i] = m_b;
  }
  // inclusive scan of the segment totals inside each warp, then across warps
  double w_rlog = m_rlog, w_ov = m_ov, w_a = m_a, w_b = m_b;
  for (int o = 1; o < 32; o <<= 1) {
    double r = __shfl_up_sync(FULL_MASK, w_rlog, o), v = __shfl_up_sync(FULL_MASK, w_ov, o);
    double a = __shfl_up_sync(FULL_MASK, w_a, o), b = __shfl_up_sync(FULL_MASK, w_b, o);
    if (lane >= o) { w_rlog = fmin(w_rlog, r); w_ov = fmin(w_ov, v); w_a = fmin(w_a, a); w_b = fmax(w_b, b); }
  }
  if (lane == 31) { s_rlog[wid] = w_rlog; s_ov[wid] = w_ov; s_a[wid] = w_a; s_b[wid] = w_b; }
  __syncthreads();
  // carry-in of this thread = totals of all earlier threads
  double c_rlog = INFINITY, c_ov = INFINITY, c_a = INFINITY, c_b = -INFINITY;
  for (int w = 0; w < wid; ++w) {
    c_rlog = fmin(c_rlog, s_rlog[w]); c_ov = fmin(c_ov, s_ov[w]); c_a = fmin(c_a, s_a[w]); c_b = fmax(c_b, s_b[w]);
  }
  double e_rlog = __shfl_up_sync(FULL_MASK, w_rlog, 1), e_ov = __shfl_up_sync(FULL_MASK, w_ov, 1);
  double e_a = __shfl_up_sync(FULL_MASK, w_a, 1), e_b = __shfl_up_sync(FULL_MASK, w_b, 1);
  if (lane > 0) { c_rlog = fmin(c_rlog, e_rlog); c_ov = fmin(c_ov, e_ov); c_a = fmin(c_a, e_a); c_b = fmax(c_b, e_b); }
  for (int i = i0; i < i1; ++i) {
    f.pmin_rlog[i] = fmin(f.pmin_rlog[i], c_rlog);
    f.pmin_ov[i] = fmin(f.pmin_ov[i], c_ov);
    f.pmin_a[i] = fmin(f.pmin_a[i], c_a);
    f.pmax_b[i] = fmax(f.pmax_b[i], c_b);
  }
}

// Exclusive scan of the start-index histogram with every run rounded up to a multiple of G (the
// per-lane pixel group of k3_fast): offs[L] .. offs[L+1] holds the samples that start at L, padded
// with -1. One CTA, each thread owns a contiguous segment of the n bins.
__global__ void __launch_bounds__(1024) k2_scan(const unsigned* hist, unsigned* offs, unsigned* cursor, int n, unsigned G) {
  __shared__ unsigned s_tot[32];
  const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, wid = t >> 5;
  const int seg = (n + nt - 1) / nt;
  const int i0 = min(n, t * seg), i1 = min(n, i0 + seg);
  unsigned acc = 0;
  for (int i = i0; i < i1; ++i) acc += (hist[i] + G - 1) / G * G;
  unsigned incl = acc;  // inclusive scan over the warp
  for (int o = 1; o < 32; o <<= 1) {
    unsigned v = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_tot[wid] = incl;
  __syncthreads();
  unsigned base = incl - acc;
  for (int w = 0; w < wid; ++w) base += s_tot[w];
  for (int i = i0; i < i1; ++i) {
    offs[i] = base; cursor[i] = base;
    base += (hist[i] + G - 1) / G * G;
  }
  if (i1 == n) offs[n] = base;  // last owner and any empty tail thread agree on the grand total
}

// Scatter fresh work indices into start-index-sorted order (warp-aggregated slot reservation).
// fresh_ids is pre-filled with -1 so the padding of each run reads as "no sample".
__global__ void __launch_bounds__(256) k2_scatter(const int32_t* init_j, long long W_fixed,
                                                  const unsigned long long* W_ptr, unsigned* cursor,
                                                  int32_t* fresh_ids) {
  const long long W = W_ptr ? (long long)*W_ptr : W_fixed;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < W; base += stride) {
    const long long w = base + lane;
    int bin = w < W ? init_j[w] : -1;
    unsigned peers = __match_any_sync(FULL_MASK, bin);
    unsigned slot0 = 0;
    int leader = __ffs(peers) - 1;
    if (bin >= 0 && lane == leader) slot0 = atomicAdd(&cursor[bin], (unsigned)__popc(peers));
    slot0 = __shfl_sync(FULL_MASK, slot0, leader);
    if (bin >= 0) fresh_ids[slot0 + __popc(peers & ((1u << lane) - 1u))] = (int32_t)w;
  }
}

}  // namespace nm
