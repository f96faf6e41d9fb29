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
#include "floatexp.cuh"
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
  const double2* A;     // descended coefficients; in floatexp mode: their mantissas (0.5 <= |m| < 1) ...
  const double2* B;
  const double2* C;
  const int2* Ae;       // ... and binary exponents (floatexp mode only)
  const int2* Be;
  const int2* Ce;
  const double2* Xhi;   // [M] X[i] descended (truncated doubles, complex.h:33-35): what phase 2 adds d to
  const double2* Xlo;   // [M] low parts of X[i]
  int M, N;
  double tol;
  EpsTab eps;           // per-column / per-row pixel offsets (mantissa + exponent in scaled frames)
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
  __device__ __forceinline__ void init(double r, double i) {
    er = r; ei = i;
    // eps2 = sq(eps): (re*re - im*im, 2.0*re*im); eps3 = eps*eps2   (mandelbrot.cpp:160-161)
    e2r = er * er - ei * ei;
    e2i = (2.0 * er) * ei;
    Cplx e3 = cmul(er, ei, e2r, e2i);
    e3r = e3.re; e3i = e3.im;
  }
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

// The same evaluator with A/B/C, eps^2, eps^3 and every intermediate in floatexp: bit-identical to
// SeriesEval whenever the double computation neither overflows nor underflows, and defined beyond
// (pixel pitch < 1e-97, where the reference's doubles overflow and it dies with SIGFPE).
struct SeriesEvalFE {
  fec eps, e2, e3;
  const double2 *A, *B, *C;
  const int2 *Ae, *Be, *Ce;
  __device__ __forceinline__ void init(double er, double ei) {
    eps.re = fe_from_double(er); eps.im = fe_from_double(ei);
    e2 = fec_sq(eps);
    e3 = fec_mul(eps, e2);
  }
  __device__ __forceinline__ void init_fe(fe er, fe ei) {
    eps.re = er; eps.im = ei;
    e2 = fec_sq(eps);
    e3 = fec_mul(eps, e2);
  }
  __device__ __forceinline__ fec d_fe(int j) const {
    if (j == 0) return eps;
    fec ta = fec_mul(load(A, Ae, j), eps);
    fec tb = fec_mul(load(B, Be, j), e2);
    fec tc = fec_mul(load(C, Ce, j), e3);
    return fec_add(fec_add(ta, tb), tc);
  }
  __device__ __forceinline__ static fec load(const double2* m, const int2* e, int i) {
    double2 mm = m[i]; int2 ee = e[i];
    fec r; r.re = fe_from_parts(mm.x, ee.x); r.im = fe_from_parts(mm.y, ee.y);
    return r;
  }
  __device__ __forceinline__ Cplx d_at(int j) const {
    Cplx r;
    if (j == 0) { r.re = fe_to_double(eps.re); r.im = fe_to_double(eps.im); return r; }
    fec ta = fec_mul(load(A, Ae, j), eps);
    fec tb = fec_mul(load(B, Be, j), e2);
    fec tc = fec_mul(load(C, Ce, j), e3);
    fec d = fec_add(fec_add(ta, tb), tc);
    r.re = fe_to_double(d.re); r.im = fe_to_double(d.im);
    return r;
  }
  __device__ __forceinline__ bool unstable(int i, double tol) const {
    fe bmag = fec_sqmag(fec_mul(load(B, Be, i), e2));
    fe cmag = fec_sqmag(fec_mul(load(C, Ce, i), e3));
    return fe_lt_nonneg(fe_mul(bmag, fe_from_double(tol)), cmag);
  }
};

template <bool FE> struct SeriesSelect;
template <> struct SeriesSelect<false> {
  typedef SeriesEval type;
  __device__ __forceinline__ static void bind(SeriesEval&, const int2*, const int2*, const int2*) {}
};
template <> struct SeriesSelect<true> {
  typedef SeriesEvalFE type;
  __device__ __forceinline__ static void bind(SeriesEvalFE& s, const int2* a, const int2* b, const int2* c) { s.Ae = a; s.Be = b; s.Ce = c; }
};

// eps-dependent set-up of the evaluator for plain (double eps) and scaled (floatexp eps) frames
template <bool SCALED> struct K2Init;
template <> struct K2Init<false> {
  template <class SE> __device__ __forceinline__ static void init(SE& se, const EpsVal<false>& e) { se.init(e.r0, e.i0); }
  __device__ __forceinline__ static bool is_zero(const EpsVal<false>& e) { return e.r0 == 0.0 && e.i0 == 0.0; }
  __device__ __forceinline__ static double log2_eps(const EpsVal<false>& e) { return log2_abs(e.r0, e.i0); }
  template <class SE> __device__ __forceinline__ static pstate state(const SE&, int) { pstate s; s.dr = s.di = 0.0; s.e = 0; return s; }
};
template <> struct K2Init<true> {
  __device__ __forceinline__ static void init(SeriesEvalFE& se, const EpsVal<true>& e) { se.init_fe(e.r, e.i); }
  __device__ __forceinline__ static bool is_zero(const EpsVal<true>& e) { return e.r.m == 0.0 && e.i.m == 0.0; }
  // log2|eps| of a floatexp eps: bring both components to the larger exponent first
  __device__ __forceinline__ static double log2_eps(const EpsVal<true>& e) {
    const int E = e.r.m == 0.0 ? e.i.e : (e.i.m == 0.0 ? e.r.e : (e.r.e > e.i.e ? e.r.e : e.i.e));
    const double a = e.r.m == 0.0 ? 0.0 : fe_scale(e.r.m, e.r.e - E), b = e.i.m == 0.0 ? 0.0 : fe_scale(e.i.m, e.i.e - E);
    return log2_abs(a, b) + (double)E;
  }
  __device__ __forceinline__ static pstate state(const SeriesEvalFE& se, int j) { return state_from_fec(se.d_fe(j)); }
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

// FEM: 0 = double series, 1 = floatexp series (double eps, plain K3 states), 2 = floatexp series +
// floatexp eps + scaled K3 states.
template <bool LITERAL, int FEM>
__global__ void __launch_bounds__(K2_THREADS) k2_series(K2Params p) {
  constexpr bool FE = FEM != 0;
  constexpr bool SCALED = FEM == 2;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long evals = 0, skipped = 0, aligned_steps = 0;
  int min_j = 0x7fffffff;

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
        EpsVal<SCALED> epsv;
        epsv.load(p.eps, pix);
        const double eps_r = epsv.r0, eps_i = epsv.i0;  // eps as doubles (flushed to 0 below the normal range when SCALED)
        typename SeriesSelect<FE>::type se;
        se.A = p.A; se.B = p.B; se.C = p.C;
        SeriesSelect<FE>::bind(se, p.Ae, p.Be, p.Ce);
        K2Init<SCALED>::init(se, epsv);

        // ---- phase 1: first unstable index --------------------------------------------------------
        int first = p.M;  // index at which the test first fires (M: never)
        bool literal = LITERAL;
        double le = 0.0;
        if (!LITERAL) {
          if (K2Init<SCALED>::is_zero(epsv)) {
            literal = false;  // b = c = 0 at every index: `0*tol < 0` never fires
          } else {
            le = K2Init<SCALED>::log2_eps(epsv);
            literal = (!FE && !(le >= K2_LE_MIN)) || !isfinite(le);
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
        double2 xh = p.Xhi[found], xl = p.Xlo[found];
        double yr = trunc_add3(xh.x, xl.x, d.re);
        double yi = trunc_add3(xh.y, xl.y, d.im);
        double mag = yr * yr + yi * yi;
        if (mag > BAILOUT2) {
          int low = 0, high = L - 1, mid = L / 2;
          while (low <= high) {
            Cplx dm = se.d_at(mid);
            double2 mh = p.Xhi[mid], ml = p.Xlo[mid];
            double mr = trunc_add3(mh.x, ml.x, dm.re);
            double mi = trunc_add3(mh.y, ml.y, dm.im);
            double mm = mr * mr + mi * mi;
            if (!(mm > BAILOUT2)) low = mid + 1;
            else { high = mid - 1; found = mid; }
            mid = (low + high) / 2;
          }
          d = se.d_at(found);
          xh = p.Xhi[found]; xl = p.Xlo[found];
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
          int j = L, off = -1, e = 0;
          if (SCALED) {  // the floatexp value of d[found]: normalised (dr, di) * 2^e
            pstate ps = K2Init<SCALED>::state(se, found);
            dr = ps.dr; di = ps.di; e = ps.e;
          }
          bool cont = true;
          if (p.align4) {  // k3_fast works on indices that are multiples of 4: take up to 3 exact steps here
            int steps = 0;
            cont = advance_checked<SCALED>(p.ck, pix, epsv, off, dr, di, e, j, (4 - (L & 3)) & 3, &steps);
            aligned_steps += (unsigned long long)steps;
          }
          if (cont) {
            p.fresh.d[w] = make_double2(dr, di);
            p.fresh.j[w] = j;
            p.fresh.off[w] = off;
            p.fresh.pix[w] = pix;
            if (SCALED) p.fresh.e[w] = e;
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
      if (bin >= 0 && bin < min_j) min_j = bin;
    }
  }
  for (int o = 16; o; o >>= 1) min_j = min(min_j, __shfl_xor_sync(FULL_MASK, min_j, o));
  if (lane == 0 && min_j != 0x7fffffff) atomicMin(&p.ctr[CTR_MINJ], (unsigned long long)min_j);

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
template <bool FE>
__global__ void __launch_bounds__(256) k2_prepare(const double2* B, const double2* C, const int2* Be, const int2* Ce, int M,
                                                  double tol, K2Filter f) {
  const double lt = log2(tol);
  const double lt_neg = fmin(lt, 0.0), lt_pos = fmax(lt, 0.0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    double rlog = INFINITY, a = INFINITY, b = -INFINITY, ov = INFINITY;
    if (i >= 1) {
      double2 Bi = B[i], Ci = C[i];
      double lb, lc;
      if (FE) {  // |m| * 2^e with per-component exponents: bring both components to the larger exponent
        int2 be = Be[i], ce = Ce[i];
        int eb = max(Bi.x != 0.0 ? be.x : INT_MIN / 2, Bi.y != 0.0 ? be.y : INT_MIN / 2);
        int ec = max(Ci.x != 0.0 ? ce.x : INT_MIN / 2, Ci.y != 0.0 ? ce.y : INT_MIN / 2);
        lb = (Bi.x == 0.0 && Bi.y == 0.0) ? -INFINITY
                                          : log2_abs(fe_scale(Bi.x, max(be.x - eb, -2000)), fe_scale(Bi.y, max(be.y - eb, -2000))) + eb;
        lc = (Ci.x == 0.0 && Ci.y == 0.0) ? -INFINITY
                                          : log2_abs(fe_scale(Ci.x, max(ce.x - ec, -2000)), fe_scale(Ci.y, max(ce.y - ec, -2000))) + ec;
      } else {
        lb = log2_abs(Bi.x, Bi.y); lc = log2_abs(Ci.x, Ci.y);
      }
      if (lc != -INFINITY) {            // C_i == 0: |c|^2 == 0, the test cannot fire
        rlog = (lb == -INFINITY) ? -INFINITY : 0.5 * lt + lb - lc;
        if (!FE) {                      // floatexp has no denormal/overflow ranges: candidates only
          a = (-545.0 - lc) / 3.0;
          b = (lb == -INFINITY) ? INFINITY : (-990.0 - lt_neg - 2.0 * lb) / 4.0;
          double c1 = (lb == -INFINITY) ? INFINITY : (1000.0 - lt_pos - 2.0 * lb) / 4.0;
          double c2 = (1000.0 - 2.0 * lc) / 6.0;
          ov = fmin(c1, c2);
          // The "denormal |b|^2*tol while |c|^2 != 0" window a < le < b is empty at almost every index
          // (it needs |C| huge next to a tiny |B|). Record empty windows as such, so that their a[] / b[]
          // do not drag the prefix minima/maxima down: with those, "first index whose window may
          // contain le" fell tens of thousands of indices before the first candidate on deep views and
          // every sample walked that gap index by index (173 ms instead of 0.5 ms on a 1280x720 frame).
          if (!(a < b)) { a = INFINITY; b = -INFINITY; }
        }
      }
    }
    f.rlog[i] = rlog; f.a[i] = a; f.b[i] = b; f.ov[i] = ov;
  }
}

// Their prefix minima/maxima over the M entries, in three small launches (a single CTA walking the
// whole table took 0.48 ms at M = 61 513, twice per frame):
//   k2_prefix_tiles   one CTA per tile of 1024: inclusive scan inside the tile (warp shuffles + one
//                     shared-memory hop), tile totals to tile_tot[4][n_tiles]
//   k2_prefix_carry   one CTA: exclusive scan of the tile totals (n_tiles <= 4096 here: M <= 2^22)
//   k2_prefix_apply   fold each tile's carry into its entries
__device__ __forceinline__ void k2_scan4(double& v0, double& v1, double& v2, double& v3, int lane) {
  for (int o = 1; o < 32; o <<= 1) {
    double u0 = __shfl_up_sync(FULL_MASK, v0, o), u1 = __shfl_up_sync(FULL_MASK, v1, o);
    double u2 = __shfl_up_sync(FULL_MASK, v2, o), u3 = __shfl_up_sync(FULL_MASK, v3, o);
    if (lane >= o) { v0 = fmin(v0, u0); v1 = fmin(v1, u1); v2 = fmin(v2, u2); v3 = fmax(v3, u3); }
  }
}

__global__ void __launch_bounds__(1024) k2_prefix_tiles(int M, K2Filter f, double* tile_tot, int n_tiles) {
  __shared__ double s_w[4][32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int i = blockIdx.x * 1024 + t;
  double v0 = INFINITY, v1 = INFINITY, v2 = INFINITY, v3 = -INFINITY;
  if (i < M) { v0 = f.rlog[i]; v1 = f.ov[i]; v2 = f.a[i]; v3 = f.b[i]; }
  k2_scan4(v0, v1, v2, v3, lane);
  if (lane == 31) { s_w[0][wid] = v0; s_w[1][wid] = v1; s_w[2][wid] = v2; s_w[3][wid] = v3; }
  __syncthreads();
  double c0 = INFINITY, c1 = INFINITY, c2 = INFINITY, c3 = -INFINITY;
  for (int w = 0; w < wid; ++w) { c0 = fmin(c0, s_w[0][w]); c1 = fmin(c1, s_w[1][w]); c2 = fmin(c2, s_w[2][w]); c3 = fmax(c3, s_w[3][w]); }
  v0 = fmin(v0, c0); v1 = fmin(v1, c1); v2 = fmin(v2, c2); v3 = fmax(v3, c3);
  if (i < M) { f.pmin_rlog[i] = v0; f.pmin_ov[i] = v1; f.pmin_a[i] = v2; f.pmax_b[i] = v3; }
  if (t == 1023) {
    tile_tot[blockIdx.x] = v0; tile_tot[n_tiles + blockIdx.x] = v1;
    tile_tot[2 * n_tiles + blockIdx.x] = v2; tile_tot[3 * n_tiles + blockIdx.x] = v3;
  }
}

// exclusive scan of the tile totals, in place; each thread owns a contiguous run of tiles
__global__ void __launch_bounds__(1024) k2_prefix_carry(double* tile_tot, int n_tiles) {
  __shared__ double s_w[4][32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (n_tiles + 1023) / 1024;
  const int b0 = min(n_tiles, t * per), b1 = min(n_tiles, b0 + per);
  double v0 = INFINITY, v1 = INFINITY, v2 = INFINITY, v3 = -INFINITY;
  for (int b = b0; b < b1; ++b) {
    v0 = fmin(v0, tile_tot[b]); v1 = fmin(v1, tile_tot[n_tiles + b]);
    v2 = fmin(v2, tile_tot[2 * n_tiles + b]); v3 = fmax(v3, tile_tot[3 * n_tiles + b]);
  }
  double r0 = v0, r1 = v1, r2 = v2, r3 = v3;  // inclusive over this thread's run -> inclusive over threads
  k2_scan4(r0, r1, r2, r3, lane);
  if (lane == 31) { s_w[0][wid] = r0; s_w[1][wid] = r1; s_w[2][wid] = r2; s_w[3][wid] = r3; }
  __syncthreads();
  double c0 = INFINITY, c1 = INFINITY, c2 = INFINITY, c3 = -INFINITY;  // everything before this thread's run
  for (int w = 0; w < wid; ++w) { c0 = fmin(c0, s_w[0][w]); c1 = fmin(c1, s_w[1][w]); c2 = fmin(c2, s_w[2][w]); c3 = fmax(c3, s_w[3][w]); }
  double e0 = __shfl_up_sync(FULL_MASK, r0, 1), e1 = __shfl_up_sync(FULL_MASK, r1, 1);
  double e2 = __shfl_up_sync(FULL_MASK, r2, 1), e3 = __shfl_up_sync(FULL_MASK, r3, 1);
  if (lane > 0) { c0 = fmin(c0, e0); c1 = fmin(c1, e1); c2 = fmin(c2, e2); c3 = fmax(c3, e3); }
  __syncthreads();
  for (int b = b0; b < b1; ++b) {  // exclusive carry of tile b, then fold the tile in
    const double t0 = tile_tot[b], t1 = tile_tot[n_tiles + b], t2 = tile_tot[2 * n_tiles + b], t3 = tile_tot[3 * n_tiles + b];
    tile_tot[b] = c0; tile_tot[n_tiles + b] = c1; tile_tot[2 * n_tiles + b] = c2; tile_tot[3 * n_tiles + b] = c3;
    c0 = fmin(c0, t0); c1 = fmin(c1, t1); c2 = fmin(c2, t2); c3 = fmax(c3, t3);
  }
}

__global__ void __launch_bounds__(1024) k2_prefix_apply(int M, K2Filter f, const double* tile_tot, int n_tiles) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i >= M || blockIdx.x == 0) return;
  const int b = blockIdx.x;
  f.pmin_rlog[i] = fmin(f.pmin_rlog[i], tile_tot[b]);
  f.pmin_ov[i] = fmin(f.pmin_ov[i], tile_tot[n_tiles + b]);
  f.pmin_a[i] = fmin(f.pmin_a[i], tile_tot[2 * n_tiles + b]);
  f.pmax_b[i] = fmax(f.pmax_b[i], tile_tot[3 * n_tiles + b]);
}

// Exclusive scan of the start-index histogram with every run rounded up to a multiple of G (the
// per-lane pixel group of k3_fast): offs[L] .. offs[L+1] holds the samples that start at L, padded
// with -1. One CTA, each thread owns a contiguous segment of the n bins.
__global__ void __launch_bounds__(1024) k2_scan(const unsigned* hist, unsigned* offs, unsigned* cursor, int n, unsigned G) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_carry;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (t == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {  // coalesced tiles of 1024 bins
    const int i = base + t;
    const unsigned v = i < n ? (hist[i] + G - 1) / G * G : 0u;
    unsigned incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned u = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    unsigned b = s_carry;
    for (int w = 0; w < wid; ++w) b += s_w[w];
    const unsigned excl = b + incl - v;
    if (i < n) { offs[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (t == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (t == 0) offs[n] = s_carry;
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
