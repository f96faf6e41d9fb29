// See hp_host.h. C++11, GMP C API only (explicit precisions => re-entrant, safe under std::thread).
#include "hp_host.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "../../include/newman_b200.h"

namespace newman_b200 {

namespace {

const double kBailout2 = 1024.0 * 1024.0;  // mandelbrot.cpp:58-59

struct Mp {  // RAII mpf at an explicit precision
  mpf_t v;
  explicit Mp(mp_bitcnt_t prec) { mpf_init2(v, prec); }
  ~Mp() { mpf_clear(v); }
  Mp(const Mp&) = delete;
  Mp& operator=(const Mp&) = delete;
};

// `k * s` as gmpxx evaluates int * mpf: mpf_mul_ui on |k|, then mpf_neg
inline void mul_int(mpf_ptr out, mpf_srcptr s, long k) {
  if (k >= 0) mpf_mul_ui(out, s, (unsigned long)k);
  else { mpf_mul_ui(out, s, 0UL - (unsigned long)k); mpf_neg(out, out); }
}

// Pixel map (mandelbrot.cpp:86-87, 271, 275): re = center.re + (c - cols/2)*sz.re,
// im = center.im + (rows/2 - r - 1)*sz.im. `tmp` and `out` at the working precision.
inline void pixel_re(const ViewHP& v, int c, mpf_ptr tmp, mpf_ptr out) {
  mul_int(tmp, v.sz_re, (long)(c - v.nc / 2));
  mpf_add(out, v.center_re, tmp);
}
inline void pixel_im(const ViewHP& v, int r, mpf_ptr tmp, mpf_ptr out) {
  mul_int(tmp, v.sz_im, (long)(v.nr / 2 - r - 1));
  mpf_add(out, v.center_im, tmp);
}

// One reference-orbit step with the reference's expression order (mandelbrot.cpp:102-103):
//   re' = re*re - im*im + x0.re ;  im' = 2.0*(re*im) + x0.im     (2.0 enters as a 64-bit mpf)
struct OrbitStep {
  Mp t1, t2, t3, two;
  explicit OrbitStep(mp_bitcnt_t prec) : t1(prec), t2(prec), t3(prec), two(64) { mpf_set_d(two.v, 2.0); }
  void operator()(mpf_ptr nre, mpf_ptr nim, mpf_srcptr re, mpf_srcptr im, mpf_srcptr x0re, mpf_srcptr x0im) {
    mpf_mul(t1.v, re, re);
    mpf_mul(t2.v, im, im);
    mpf_sub(t3.v, t1.v, t2.v);
    mpf_mul(t1.v, re, im);  // before nre is written: nre may alias re
    mpf_add(nre, t3.v, x0re);
    mpf_mul(t2.v, two.v, t1.v);
    mpf_add(nim, t2.v, x0im);
  }
};

inline bool bailed(mpf_srcptr re, mpf_srcptr im) {  // bailedOut, mandelbrot.cpp:61
  double r = mpf_get_d(re), i = mpf_get_d(im);
  return r * r + i * i > kBailout2;
}

// X.size() of computeOrbit (mandelbrot.cpp:97-110) without storing the orbit.
int orbit_length(const ViewHP& v, mpf_srcptr x0re, mpf_srcptr x0im) {
  Mp re(v.prec), im(v.prec), nre(v.prec), nim(v.prec);
  OrbitStep step(v.prec);
  mpf_set(re.v, x0re);
  mpf_set(im.v, x0im);
  for (int i = 1; i < v.N; i++) {
    step(nre.v, nim.v, re.v, im.v, x0re, x0im);
    if (bailed(nre.v, nim.v)) return i;
    mpf_swap(re.v, nre.v);
    mpf_swap(im.v, nim.v);
  }
  return v.N;
}

int pick_threads(int threads) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  return threads;
}

// inCardioid (mandelbrot.cpp:63-71) on an mpf point; optionally returns the two margins as doubles.
bool in_cardioid(mp_bitcnt_t prec, mpf_srcptr zre, mpf_srcptr zim, double* margin1, double* margin2) {
  Mp fourth(prec), xmf(prec), y2(prec), q(prec), t(prec), lhs(prec), rhs(prec), one(64);
  mpf_set_d(fourth.v, 0.25);
  mpf_set_d(one.v, 1.0);
  mpf_sub(xmf.v, zre, fourth.v);
  mpf_mul(y2.v, zim, zim);
  mpf_mul(t.v, xmf.v, xmf.v);
  mpf_add(q.v, t.v, y2.v);
  mpf_add(t.v, q.v, xmf.v);
  mpf_mul(lhs.v, q.v, t.v);
  mpf_mul(rhs.v, fourth.v, y2.v);
  bool in1 = mpf_cmp(lhs.v, rhs.v) < 0;
  if (margin1) { mpf_sub(t.v, lhs.v, rhs.v); *margin1 = mpf_get_d(t.v); }
  if (in1 && !margin2) return true;
  mpf_add(q.v, zre, one.v);
  mpf_mul(t.v, q.v, q.v);
  mpf_add(lhs.v, t.v, y2.v);
  mpf_mul(rhs.v, fourth.v, fourth.v);
  bool in2 = mpf_cmp(lhs.v, rhs.v) < 0;
  if (margin2) { mpf_sub(t.v, lhs.v, rhs.v); *margin2 = mpf_get_d(t.v); }
  return in1 || in2;
}

}  // namespace

int precision_bits_for(mpf_srcptr sz_re) {
  const double log_alpha = log2(1.0e-20);
  const int beta = 64;
  signed long int e;
  mpf_get_d_2exp(&e, sz_re);
  int bits = (int)(beta - e + log_alpha);
  if (bits < 64) bits = 64;
  return bits;
}

void pixel_coords(const ViewHP& v, std::vector<double>& c_re, std::vector<double>& c_im) {
  Mp tmp(v.prec), p(v.prec);
  c_re.resize(v.nc);
  c_im.resize(v.nr);
  for (int c = 0; c < v.nc; c++) { pixel_re(v, c, tmp.v, p.v); c_re[c] = mpf_get_d(p.v); }
  for (int r = 0; r < v.nr; r++) { pixel_im(v, r, tmp.v, p.v); c_im[r] = mpf_get_d(p.v); }
}

bool in_cardioid_pixel(const ViewHP& v, int r, int c) {
  Mp tmp(v.prec), re(v.prec), im(v.prec);
  pixel_re(v, c, tmp.v, re.v);
  pixel_im(v, r, tmp.v, im.v);
  return in_cardioid(v.prec, re.v, im.v, nullptr, nullptr);
}

int classify_cardioid(const ViewHP& v, int threads, std::vector<uint8_t>& mask) {
  // The two margins are polynomials in (x, y) with |grad| < 200 wherever they are small, so if both
  // exceed 200 * (view extent) at the centre sample every sample of the view decides like it.
  Mp tmp(v.prec), re(v.prec), im(v.prec);
  pixel_re(v, v.nc / 2, tmp.v, re.v);
  pixel_im(v, v.nr / 2, tmp.v, im.v);
  double m1 = 0, m2 = 0;
  bool inside = in_cardioid(v.prec, re.v, im.v, &m1, &m2);
  double extent = fabs(mpf_get_d(v.sz_re)) * v.nc + fabs(mpf_get_d(v.sz_im)) * v.nr;
  if (fabs(m1) > 200.0 * extent && fabs(m2) > 200.0 * extent) return inside ? NM_CARDIOID_ALL : NM_CARDIOID_NONE;

  mask.assign((size_t)v.nr * v.nc, 0);
  threads = pick_threads(threads);
  std::atomic<int> next(0);
  auto work = [&]() {
    Mp t(v.prec), pre(v.prec), pim(v.prec);
    for (;;) {
      int r = next.fetch_add(1);
      if (r >= v.nr) break;
      pixel_im(v, r, t.v, pim.v);
      for (int c = 0; c < v.nc; c++) {
        pixel_re(v, c, t.v, pre.v);
        mask[(size_t)r * v.nc + c] = in_cardioid(v.prec, pre.v, pim.v, nullptr, nullptr) ? 1 : 0;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int i = 1; i < threads; i++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return NM_CARDIOID_MASK;
}

void probe_candidates(const ViewHP& v, std::vector<std::pair<int, int> >& cand) {
  cand.clear();
  for (int c = 0; c < v.nc; c += 2) {
    cand.emplace_back(v.nr / 4, c);
    cand.emplace_back(v.nr / 2, c);
    cand.emplace_back(3 * v.nr / 4, c);
  }
  for (int r = 0; r < v.nr; r += 2) cand.emplace_back(r, v.nc / 2);
}

void probe_lengths(const ViewHP& v, const std::vector<std::pair<int, int> >& cand, const std::vector<int>& which,
                   int threads, std::vector<int>& len) {
  const int n = (int)which.size();
  len.assign(n, -1);
  std::atomic<int> next(0);
  threads = pick_threads(threads);
  if (threads > n) threads = n;
  auto work = [&]() {
    Mp tmp(v.prec), pre(v.prec), pim(v.prec);
    for (;;) {
      int k = next.fetch_add(1);
      if (k >= n) break;
      pixel_re(v, cand[which[k]].second, tmp.v, pre.v);
      pixel_im(v, cand[which[k]].first, tmp.v, pim.v);
      len[k] = orbit_length(v, pre.v, pim.v);
    }
  };
  std::vector<std::thread> pool;
  for (int i = 1; i < threads; i++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
}

void find_probe(const ViewHP& v, int threads, int& row, int& col, int& length) {
  // candidate order of mandelbrot.cpp:77-83
  std::vector<std::pair<int, int>> cand;
  probe_candidates(v, cand);
  const int n = (int)cand.size();
  std::vector<int> len(n, -1);
  // A probe that never escapes (length N) cannot be beaten by a later one (strict '>' at :90), so
  // candidates after the first such probe need not be evaluated.
  std::atomic<int> next(0), first_full(n);
  threads = pick_threads(threads);
  if (threads > n) threads = n;
  auto work = [&]() {
    Mp tmp(v.prec), pre(v.prec), pim(v.prec);
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= n) break;
      if (i > first_full.load()) continue;
      pixel_re(v, cand[i].second, tmp.v, pre.v);
      pixel_im(v, cand[i].first, tmp.v, pim.v);
      int L = orbit_length(v, pre.v, pim.v);
      len[i] = L;
      if (L >= v.N) {
        int cur = first_full.load();
        while (i < cur && !first_full.compare_exchange_weak(cur, i)) {}
      }
    }
  };
  std::vector<std::thread> pool;
  for (int i = 1; i < threads; i++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  int best = 0;
  for (int i = 1; i < n; i++)
    if (len[i] > len[best]) best = i;  // first longest wins
  row = cand[best].first;
  col = cand[best].second;
  length = len[best];
}

// Orbit and series one after the other on the calling thread (host_threads == 1, and the cross-check of the
// pipelined form below). Leaves X[0] in x0re / x0im.
static void tables_serial(const ViewHP& v, mpf_srcptr x0re_in, mpf_srcptr x0im_in, DeepTablesHost& out) {
  const mp_bitcnt_t P = v.prec;
  Mp x0re(P), x0im(P);
  mpf_set(x0re.v, x0re_in); mpf_set(x0im.v, x0im_in);

  // ---- computeOrbit (mandelbrot.cpp:97-110), keeping the escaped iterate the reference drops ----
  struct Pair { mpf_t re, im; };
  std::vector<Pair> X;
  X.reserve(1024);
  auto push = [&](mpf_srcptr re, mpf_srcptr im) {
    X.emplace_back();
    mpf_init2(X.back().re, P); mpf_init2(X.back().im, P);
    mpf_set(X.back().re, re); mpf_set(X.back().im, im);
  };
  push(x0re.v, x0im.v);
  {
    OrbitStep step(P);
    Mp nre(P), nim(P);
    out.has_escape = false;
    for (int i = 1; i < v.N; i++) {
      step(nre.v, nim.v, X[i - 1].re, X[i - 1].im, x0re.v, x0im.v);
      push(nre.v, nim.v);
      if (bailed(nre.v, nim.v)) { out.has_escape = true; break; }
    }
  }
  const int M = (int)X.size() - (out.has_escape ? 1 : 0);
  out.M = M;

  // ---- descend X: hi = trunc(X), lo = trunc(X - hi) -----------------------------------------------
  out.x_hi.resize(2 * (size_t)X.size());
  out.x_lo.resize(2 * (size_t)M);
  {
    Mp d64(64), rem(P);
    for (size_t i = 0; i < X.size(); i++) {
      double hr = mpf_get_d(X[i].re), hi = mpf_get_d(X[i].im);
      out.x_hi[2 * i] = hr;
      out.x_hi[2 * i + 1] = hi;
      if ((int)i < M) {
        mpf_set_d(d64.v, hr); mpf_sub(rem.v, X[i].re, d64.v); out.x_lo[2 * i] = mpf_get_d(rem.v);
        mpf_set_d(d64.v, hi); mpf_sub(rem.v, X[i].im, d64.v); out.x_lo[2 * i + 1] = mpf_get_d(rem.v);
      }
    }
  }

  // ---- computeSeries (mandelbrot.cpp:112-131), descending as we go --------------------------------
  // Note the reference's recurrences use the NEW A[i] in B[i] and the new A[i], B[i] in C[i].
  out.a.resize(2 * (size_t)M); out.b.resize(2 * (size_t)M); out.c.resize(2 * (size_t)M);
  out.a_m.resize(2 * (size_t)M); out.b_m.resize(2 * (size_t)M); out.c_m.resize(2 * (size_t)M);
  out.a_e.resize(2 * (size_t)M); out.b_e.resize(2 * (size_t)M); out.c_e.resize(2 * (size_t)M);
  {
    Mp ar(P), ai(P), br(P), bi(P), cr(P), ci(P), nar(P), nai(P), nbr(P), nbi(P), ncr(P), nci(P);
    Mp p1(P), p2(P), s(P), u(P), two(64), one(64);
    mpf_set_d(two.v, 2.0);
    mpf_set_d(one.v, 1.0);
    mpf_set_d(ar.v, 1.0);  // A[0] = 1, B[0] = C[0] = 0
    auto split = [](mpf_srcptr x, double& m, int32_t& e) {  // x = m * 2^e, 0.5 <= |m| < 1, truncating
      long ex = 0;
      m = mpf_get_d_2exp(&ex, x);
      e = (int32_t)ex;
    };
    auto store = [&](int i) {
      out.a[2 * i] = mpf_get_d(ar.v); out.a[2 * i + 1] = mpf_get_d(ai.v);
      out.b[2 * i] = mpf_get_d(br.v); out.b[2 * i + 1] = mpf_get_d(bi.v);
      out.c[2 * i] = mpf_get_d(cr.v); out.c[2 * i + 1] = mpf_get_d(ci.v);
      split(ar.v, out.a_m[2 * i], out.a_e[2 * i]); split(ai.v, out.a_m[2 * i + 1], out.a_e[2 * i + 1]);
      split(br.v, out.b_m[2 * i], out.b_e[2 * i]); split(bi.v, out.b_m[2 * i + 1], out.b_e[2 * i + 1]);
      split(cr.v, out.c_m[2 * i], out.c_e[2 * i]); split(ci.v, out.c_m[2 * i + 1], out.c_e[2 * i + 1]);
    };
    store(0);
    for (int i = 1; i < M; i++) {
      mpf_srcptr xr = X[i - 1].re, xi = X[i - 1].im;
      // A[i].re = 2.0 * (xr*ar - xi*ai) + 1.0
      mpf_mul(p1.v, xr, ar.v); mpf_mul(p2.v, xi, ai.v); mpf_sub(s.v, p1.v, p2.v);
      mpf_mul(u.v, two.v, s.v); mpf_add(nar.v, u.v, one.v);
      // A[i].im = 2.0 * (xr*ai + xi*ar)
      mpf_mul(p1.v, xr, ai.v); mpf_mul(p2.v, xi, ar.v); mpf_add(s.v, p1.v, p2.v);
      mpf_mul(nai.v, two.v, s.v);
      // B[i].re = 2.0 * (xr*br - xi*bi) + A.re*A.re - A.im*A.im        (new A)
      mpf_mul(p1.v, xr, br.v); mpf_mul(p2.v, xi, bi.v); mpf_sub(s.v, p1.v, p2.v);
      mpf_mul(u.v, two.v, s.v);
      mpf_mul(p1.v, nar.v, nar.v); mpf_add(s.v, u.v, p1.v);
      mpf_mul(p2.v, nai.v, nai.v); mpf_sub(nbr.v, s.v, p2.v);
      // B[i].im = 2.0 * (xr*bi + xi*br + A.re*A.im)
      mpf_mul(p1.v, xr, bi.v); mpf_mul(p2.v, xi, br.v); mpf_add(s.v, p1.v, p2.v);
      mpf_mul(p1.v, nar.v, nai.v); mpf_add(u.v, s.v, p1.v);
      mpf_mul(nbi.v, two.v, u.v);
      // C[i].re = 2.0 * (xr*cr - xi*ci + A.re*B.re - A.im*B.im)         (new A, new B)
      mpf_mul(p1.v, xr, cr.v); mpf_mul(p2.v, xi, ci.v); mpf_sub(s.v, p1.v, p2.v);
      mpf_mul(p1.v, nar.v, nbr.v); mpf_add(u.v, s.v, p1.v);
      mpf_mul(p2.v, nai.v, nbi.v); mpf_sub(s.v, u.v, p2.v);
      mpf_mul(ncr.v, two.v, s.v);
      // C[i].im = 2.0 * (xr*ci + xi*cr + A.re*B.im + A.im*B.re)
      mpf_mul(p1.v, xr, ci.v); mpf_mul(p2.v, xi, cr.v); mpf_add(s.v, p1.v, p2.v);
      mpf_mul(p1.v, nar.v, nbi.v); mpf_add(u.v, s.v, p1.v);
      mpf_mul(p2.v, nai.v, nbr.v); mpf_add(s.v, u.v, p2.v);
      mpf_mul(nci.v, two.v, s.v);
      mpf_swap(ar.v, nar.v); mpf_swap(ai.v, nai.v);
      mpf_swap(br.v, nbr.v); mpf_swap(bi.v, nbi.v);
      mpf_swap(cr.v, ncr.v); mpf_swap(ci.v, nci.v);
      store(i);
    }
  }
  for (auto& x : X) { mpf_clear(x.re); mpf_clear(x.im); }
}

// The same tables from a pipeline: the orbit X and the three coefficient recurrences each on their own thread.
// A[i] needs X[i-1] and A[i-1]; B[i] needs X[i-1], B[i-1] and the NEW A[i]; C[i] needs X[i-1], C[i-1] and the new
// A[i], B[i] (mandelbrot.cpp:118-128) — so stage S at index i only waits for stage S-1 at index i. Per index the four
// stages cost 3 / 6 / 9 / 10 multiplications, so the pipeline runs at the pace of C: ~2.7x faster than the serial form.
// With 6 or more threads the products that do not involve the stage's own previous value — A[i]^2 for B, A[i] B[i] for C
// — come from two more stages (AA, AB) that run ahead, and A, B, C are left with 6 multiplications each.
// Every stage issues exactly the mpf operations of tables_serial on operands of the same precision, and every sum is
// formed in the serial form's order, so the tables are bit-identical (tests/test_host_tables.py compares the forms with
// each other and with the compiled reference). This is what a deep frame spends most of its host time in: at 1e-400
// (M = 5e5, 1344 bits) one reference costs 6.6 s serially and a frame builds three.
// Memory: a stream gives a block of 4096 entries back once the last stage (C) is past it, so what is held is the lead of
// each stage over C — mostly the orbit's, which the serial form keeps in full as well.
namespace {

template <int K>
struct HpVals { mpf_t v[K]; };

// HpStream::append() lays mpf values out by hand (an mpf_init2 per value would be a malloc per value): {_mp_prec,
// _mp_size, _mp_exp, _mp_d} with prec + 1 limbs from a pool. That ties this file to GMP's documented-but-private mpf
// layout, so it is checked once per process against the running libgmp: a hand-built value must take the same results
// as an mpf_init2 value (same limbs, size, exponent) from mpf_set / mpf_mul / mpf_add at several precisions. If the
// check fails the pooled layout is not used: every entry falls back to mpf_init2 (slower, always right).
static_assert(sizeof(mp_limb_t) == 8, "hp_host.cpp assumes 64-bit limbs");
static_assert(sizeof(__mpf_struct) == 24, "unexpected __mpf_struct layout");
bool pooled_mpf_layout_ok() {
  static const bool ok = [] {
    for (mp_bitcnt_t bits : {64u, 173u, 341u, 1336u}) {
      mpf_t ref, a, b;
      mpf_init2(ref, bits); mpf_init2(a, bits); mpf_init2(b, bits);
      mpf_set_d(a, 1.2345678901234567); mpf_div_ui(a, a, 3UL);
      mpf_set_d(b, -0.9876543210987654); mpf_div_ui(b, b, 7UL);
      const int prec = ref->_mp_prec;
      std::vector<mp_limb_t> pool((size_t)prec + 1, ~(mp_limb_t)0);
      __mpf_struct hand;
      hand._mp_prec = prec; hand._mp_size = 0; hand._mp_exp = 0; hand._mp_d = pool.data();
      bool same = true;
      for (int op = 0; op < 3 && same; op++) {
        if (op == 0) { mpf_set(ref, a); mpf_set(&hand, a); }
        else if (op == 1) { mpf_mul(ref, a, b); mpf_mul(&hand, a, b); }
        else { mpf_add(ref, ref, b); mpf_add(&hand, &hand, b); }
        same = ref->_mp_size == hand._mp_size && ref->_mp_exp == hand._mp_exp && mpf_cmp(ref, &hand) == 0;
        const int n = ref->_mp_size < 0 ? -ref->_mp_size : ref->_mp_size;
        for (int i = 0; i < n && same; i++) same = ref->_mp_d[i] == hand._mp_d[i];
        same = same && n <= prec + 1;
      }
      mpf_clear(ref); mpf_clear(a); mpf_clear(b);
      if (!same) return false;
    }
    return true;
  }();
  return ok;
}

// Values a stage publishes for the next ones. Blocks of 4096 entries hang off a pointer table of fixed size, so a
// reader never sees storage move; `ready` = entries published (release / acquire).
template <int K>
struct HpStream {
  static const int kBlock = 4096;
  static const long kBatch = 32;        // entries per publication: the counter's cache line changes hands rarely
  std::vector<HpVals<K>*> blocks;
  std::vector<mp_limb_t*> limbs;        // one limb pool per block (an mpf_init2 per value would be a malloc per value)
  int prec_limbs;                        // _mp_prec of an mpf at the stream's precision
  mp_bitcnt_t prec_bits;
  bool pooled;                           // entries are laid out by hand in a limb pool (pooled_mpf_layout_ok)
  alignas(64) std::atomic<long> ready;
  std::atomic<bool> finished;
  alignas(64) long allocated;           // producer's side (the consumers keep their own copies of `ready`)
  long freed_blocks;
  const std::atomic<long>* dead_below;  // entries below this position are read by nobody any more (the last stage's progress)
  HpStream(long max_entries, mp_bitcnt_t p, const std::atomic<long>* dead)
      : blocks((size_t)(max_entries / kBlock + 2), nullptr), limbs((size_t)(max_entries / kBlock + 2), nullptr), ready(0),
        finished(false), allocated(0), freed_blocks(0), dead_below(dead) {
    mpf_t t;
    mpf_init2(t, p);
    prec_limbs = t->_mp_prec;            // what mpf_init2 derives from the bit count
    mpf_clear(t);
    prec_bits = p;
    pooled = pooled_mpf_layout_ok();
  }
  ~HpStream() {
    if (!pooled)
      for (size_t bi = 0; bi < blocks.size(); bi++)
        if (blocks[bi]) {
          const long n = std::min<long>(kBlock, allocated - (long)bi * kBlock);
          for (long i = 0; i < n; i++) for (int k = 0; k < K; k++) mpf_clear(blocks[bi][i].v[k]);
        }
    for (HpVals<K>* b : blocks) delete[] b;
    for (mp_limb_t* l : limbs) delete[] l;
  }
  HpVals<K>& at(long i) { return blocks[(size_t)(i / kBlock)][i % kBlock]; }
  // producer only: a fresh entry (values 0) laid out like mpf_init2 does it — prec + 1 limbs each — inside the block's
  // pool; such a value is written by mpf_set / mpf_mul and never passed to mpf_clear / mpf_set_prec
  HpVals<K>& append() {
    const long i = allocated++;
    const size_t bi = (size_t)(i / kBlock);
    const size_t per = (size_t)prec_limbs + 1;
    if (i % kBlock == 0) {
      // a new block: first give back the blocks every reader is past (producer-side only, so `blocks` never changes
      // under a reader that could still want the entry)
      const long dead = dead_below->load(std::memory_order_acquire);
      while ((freed_blocks + 1) * kBlock <= dead) {
        if (!pooled)
          for (long e = 0; e < kBlock; e++) for (int k = 0; k < K; k++) mpf_clear(blocks[(size_t)freed_blocks][e].v[k]);
        delete[] blocks[(size_t)freed_blocks]; blocks[(size_t)freed_blocks] = nullptr;
        delete[] limbs[(size_t)freed_blocks]; limbs[(size_t)freed_blocks] = nullptr;
        freed_blocks++;
      }
      blocks[bi] = new HpVals<K>[kBlock];
      if (pooled) limbs[bi] = new mp_limb_t[K * per * kBlock];
    }
    HpVals<K>& q = at(i);
    if (!pooled) {
      for (int k = 0; k < K; k++) mpf_init2(q.v[k], prec_bits);
      return q;
    }
    mp_limb_t* base = limbs[bi] + K * per * (size_t)(i % kBlock);
    for (int k = 0; k < K; k++) {
      q.v[k]->_mp_prec = prec_limbs; q.v[k]->_mp_size = 0; q.v[k]->_mp_exp = 0; q.v[k]->_mp_d = base + k * per;
    }
    return q;
  }
  void publish(bool force = false) {
    if (force || allocated - ready.load(std::memory_order_relaxed) >= kBatch) ready.store(allocated, std::memory_order_release);
  }
  void finish() {
    ready.store(allocated, std::memory_order_release);
    finished.store(true, std::memory_order_release);
  }
  // true once entry i is there; false if the producer finished without it. `have` is the caller's cached copy of
  // `ready`: the shared counter is only read when the consumer has caught up with what it last saw.
  bool wait_for(long i, long& have) {
    if (i < have) return true;
    for (int spin = 0;; ++spin) {
      have = ready.load(std::memory_order_acquire);
      if (have > i) return true;
      if (finished.load(std::memory_order_acquire)) {
        have = ready.load(std::memory_order_acquire);
        return have > i;
      }
      if (spin > 20000) std::this_thread::sleep_for(std::chrono::microseconds(50));   // oversubscribed host: stop burning a core
      else if (spin > 256) std::this_thread::yield();
    }
  }
};

struct CoefOut {  // one coefficient's descended tables
  std::vector<double>* d; std::vector<double>* m; std::vector<int32_t>* e;
  void push(mpf_srcptr re, mpf_srcptr im) {
    long ex = 0;
    d->push_back(mpf_get_d(re)); d->push_back(mpf_get_d(im));
    m->push_back(mpf_get_d_2exp(&ex, re)); e->push_back((int32_t)ex);
    m->push_back(mpf_get_d_2exp(&ex, im)); e->push_back((int32_t)ex);
  }
};

}  // namespace

// split_products: the A[i]^2 and A[i] B[i] products come from their own stages (6 threads in all) instead of from B and C.
static void tables_pipelined(const ViewHP& v, mpf_srcptr x0re, mpf_srcptr x0im, DeepTablesHost& out, bool split_products) {
  const mp_bitcnt_t P = v.prec;
  // X holds the non-escaped iterates only (what the series reads); A and B hold the new coefficients of index i at
  // entry i - 1 (index 0 is the constant start value); AA = (A.re^2, A.im^2, A.re A.im) and
  // AB = (A.re B.re, A.im B.im, A.re B.im, A.im B.re) of the same index at the same entry
  // C is the last stage: whatever position it is past, every other stage is past as well, so the producers give those
  // blocks back as they go (the orbit may still run far ahead of the series: that part stays, like in the serial form)
  std::atomic<long> c_done(0);
  HpStream<2> X(v.N, P, &c_done), A(v.N, P, &c_done), B(v.N, P, &c_done);
  HpStream<3> AA(split_products ? v.N : 1, P, &c_done);
  HpStream<4> AB(split_products ? v.N : 1, P, &c_done);
  out.a.clear(); out.b.clear(); out.c.clear(); out.a_m.clear(); out.b_m.clear(); out.c_m.clear();
  out.a_e.clear(); out.b_e.clear(); out.c_e.clear(); out.x_hi.clear(); out.x_lo.clear();
  out.has_escape = false;
  std::vector<std::thread> stages;

  stages.emplace_back([&]() {   // ---- A[i] = 2 X[i-1] A[i-1] + 1
    Mp ar(P), ai(P), nar(P), nai(P), p1(P), p2(P), s(P), u(P), two(64), one(64);
    mpf_set_d(two.v, 2.0); mpf_set_d(one.v, 1.0); mpf_set_d(ar.v, 1.0);
    CoefOut o = {&out.a, &out.a_m, &out.a_e};
    long have_x = 0;
    if (X.wait_for(0, have_x)) o.push(ar.v, ai.v);
    for (long i = 1; X.wait_for(i, have_x); i++) {   // index i exists iff X[i] is a non-escaped iterate
      mpf_srcptr xr = X.at(i - 1).v[0], xi = X.at(i - 1).v[1];
      mpf_mul(p1.v, xr, ar.v); mpf_mul(p2.v, xi, ai.v); mpf_sub(s.v, p1.v, p2.v);
      mpf_mul(u.v, two.v, s.v); mpf_add(nar.v, u.v, one.v);
      mpf_mul(p1.v, xr, ai.v); mpf_mul(p2.v, xi, ar.v); mpf_add(s.v, p1.v, p2.v);
      mpf_mul(nai.v, two.v, s.v);
      mpf_swap(ar.v, nar.v); mpf_swap(ai.v, nai.v);
      HpVals<2>& q = A.append();
      mpf_set(q.v[0], ar.v); mpf_set(q.v[1], ai.v);
      A.publish();
      o.push(ar.v, ai.v);
    }
    A.finish();
  });
  if (split_products)
    stages.emplace_back([&]() {   // ---- the products of the new A that B[i] adds
      long have_a = 0;
      for (long e = 0; A.wait_for(e, have_a); e++) {
        mpf_srcptr nar = A.at(e).v[0], nai = A.at(e).v[1];
        HpVals<3>& q = AA.append();
        mpf_mul(q.v[0], nar, nar); mpf_mul(q.v[1], nai, nai); mpf_mul(q.v[2], nar, nai);
        AA.publish();
      }
      AA.finish();
    });
  stages.emplace_back([&]() {   // ---- B[i] = 2 X[i-1] B[i-1] + A[i]^2
    Mp br(P), bi(P), nbr(P), nbi(P), p1(P), p2(P), s(P), u(P), q1(P), q2(P), q3(P), two(64);
    mpf_set_d(two.v, 2.0);
    CoefOut o = {&out.b, &out.b_m, &out.b_e};
    long have_x = 0, have_a = 0, have_aa = 0;
    if (X.wait_for(0, have_x)) o.push(br.v, bi.v);
    for (long i = 1; split_products ? AA.wait_for(i - 1, have_aa) : A.wait_for(i - 1, have_a); i++) {
      mpf_srcptr xr = X.at(i - 1).v[0], xi = X.at(i - 1).v[1];
      mpf_srcptr a2r, a2i, ari;   // A.re*A.re, A.im*A.im, A.re*A.im of the new A
      if (split_products) { a2r = AA.at(i - 1).v[0]; a2i = AA.at(i - 1).v[1]; ari = AA.at(i - 1).v[2]; }
      else {
        mpf_srcptr nar = A.at(i - 1).v[0], nai = A.at(i - 1).v[1];
        mpf_mul(q1.v, nar, nar); mpf_mul(q2.v, nai, nai); mpf_mul(q3.v, nar, nai);
        a2r = q1.v; a2i = q2.v; ari = q3.v;
      }
      mpf_mul(p1.v, xr, br.v); mpf_mul(p2.v, xi, bi.v); mpf_sub(s.v, p1.v, p2.v);
      mpf_mul(u.v, two.v, s.v);
      mpf_add(s.v, u.v, a2r);
      mpf_sub(nbr.v, s.v, a2i);
      mpf_mul(p1.v, xr, bi.v); mpf_mul(p2.v, xi, br.v); mpf_add(s.v, p1.v, p2.v);
      mpf_add(u.v, s.v, ari);
      mpf_mul(nbi.v, two.v, u.v);
      mpf_swap(br.v, nbr.v); mpf_swap(bi.v, nbi.v);
      HpVals<2>& q = B.append();
      mpf_set(q.v[0], br.v); mpf_set(q.v[1], bi.v);
      B.publish();
      o.push(br.v, bi.v);
    }
    B.finish();
  });
  if (split_products)
    stages.emplace_back([&]() {   // ---- the products of the new A and B that C[i] adds
      long have_b = 0;
      for (long e = 0; B.wait_for(e, have_b); e++) {   // B[e] published => A[e] is there as well
        mpf_srcptr nar = A.at(e).v[0], nai = A.at(e).v[1], nbr = B.at(e).v[0], nbi = B.at(e).v[1];
        HpVals<4>& q = AB.append();
        mpf_mul(q.v[0], nar, nbr); mpf_mul(q.v[1], nai, nbi); mpf_mul(q.v[2], nar, nbi); mpf_mul(q.v[3], nai, nbr);
        AB.publish();
      }
      AB.finish();
    });
  stages.emplace_back([&]() {   // ---- C[i] = 2 (X[i-1] C[i-1] + A[i] B[i])
    Mp cr(P), ci(P), ncr(P), nci(P), p1(P), p2(P), s(P), u(P), q1(P), q2(P), q3(P), q4(P), two(64);
    mpf_set_d(two.v, 2.0);
    CoefOut o = {&out.c, &out.c_m, &out.c_e};
    long have_x = 0, have_b = 0, have_ab = 0;
    if (X.wait_for(0, have_x)) o.push(cr.v, ci.v);
    for (long i = 1; split_products ? AB.wait_for(i - 1, have_ab) : B.wait_for(i - 1, have_b); i++) {
      mpf_srcptr xr = X.at(i - 1).v[0], xi = X.at(i - 1).v[1];
      mpf_srcptr rr, ii, ri, ir;   // A.re*B.re, A.im*B.im, A.re*B.im, A.im*B.re of the new A, B
      if (split_products) { rr = AB.at(i - 1).v[0]; ii = AB.at(i - 1).v[1]; ri = AB.at(i - 1).v[2]; ir = AB.at(i - 1).v[3]; }
      else {
        mpf_srcptr nar = A.at(i - 1).v[0], nai = A.at(i - 1).v[1], nbr = B.at(i - 1).v[0], nbi = B.at(i - 1).v[1];
        mpf_mul(q1.v, nar, nbr); mpf_mul(q2.v, nai, nbi); mpf_mul(q3.v, nar, nbi); mpf_mul(q4.v, nai, nbr);
        rr = q1.v; ii = q2.v; ri = q3.v; ir = q4.v;
      }
      mpf_mul(p1.v, xr, cr.v); mpf_mul(p2.v, xi, ci.v); mpf_sub(s.v, p1.v, p2.v);
      mpf_add(u.v, s.v, rr);
      mpf_sub(s.v, u.v, ii);
      mpf_mul(ncr.v, two.v, s.v);
      mpf_mul(p1.v, xr, ci.v); mpf_mul(p2.v, xi, cr.v); mpf_add(s.v, p1.v, p2.v);
      mpf_add(u.v, s.v, ri);
      mpf_add(s.v, u.v, ir);
      mpf_mul(nci.v, two.v, s.v);
      mpf_swap(cr.v, ncr.v); mpf_swap(ci.v, nci.v);
      o.push(cr.v, ci.v);
      // index i is done: C's next reads are X[i] and entry i of A, B, AB; positions below i are dead for everybody
      if ((i & 255) == 0) c_done.store(i, std::memory_order_release);
    }
  });

  // ---- stage X on the calling thread: computeOrbit + the descended hi / lo parts as it goes -------
  {
    OrbitStep step(P);
    Mp nre(P), nim(P), d64(64), rem(P);
    auto descend_pair = [&](mpf_srcptr re, mpf_srcptr im, bool with_lo) {
      const double hr = mpf_get_d(re), hi = mpf_get_d(im);
      out.x_hi.push_back(hr); out.x_hi.push_back(hi);
      if (with_lo) {
        mpf_set_d(d64.v, hr); mpf_sub(rem.v, re, d64.v); out.x_lo.push_back(mpf_get_d(rem.v));
        mpf_set_d(d64.v, hi); mpf_sub(rem.v, im, d64.v); out.x_lo.push_back(mpf_get_d(rem.v));
      }
    };
    HpVals<2>& q0 = X.append();
    mpf_set(q0.v[0], x0re); mpf_set(q0.v[1], x0im);
    descend_pair(q0.v[0], q0.v[1], true);
    X.publish(true);
    for (int i = 1; i < v.N; i++) {
      HpVals<2>& prev = X.at(i - 1);
      step(nre.v, nim.v, prev.v[0], prev.v[1], x0re, x0im);
      if (bailed(nre.v, nim.v)) {   // the escaped iterate: kept in x_hi only (the reference drops it)
        out.has_escape = true;
        descend_pair(nre.v, nim.v, false);
        break;
      }
      HpVals<2>& q = X.append();
      mpf_set(q.v[0], nre.v); mpf_set(q.v[1], nim.v);
      descend_pair(q.v[0], q.v[1], true);
      X.publish();
    }
    X.finish();
  }
  for (std::thread& t : stages) t.join();
  out.M = (int)X.allocated;
}

void build_tables(const ViewHP& v, int row, int col, DeepTablesHost& out, int threads) {
  const mp_bitcnt_t P = v.prec;
  out.probe_row = row;
  out.probe_col = col;
  Mp tmp(P), x0re(P), x0im(P);
  pixel_re(v, col, tmp.v, x0re.v);
  pixel_im(v, row, tmp.v, x0im.v);
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  static const bool debug = getenv("NM_DEBUG_TABLES") != nullptr;   // NM_DEBUG_TABLES=1: time of orbit + series per call
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  if (threads >= 4) tables_pipelined(v, x0re.v, x0im.v, out, threads >= 6);
  else tables_serial(v, x0re.v, x0im.v, out);
  if (debug)
    fprintf(stderr, "nm build_tables: M %d, %d bits, %s: %.3f s\n", out.M, (int)P, threads >= 6 ? "pipelined (6 stages)" : threads >= 4 ? "pipelined (4 stages)" : "serial",
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  out.finite = true;
  for (size_t i = 0; i < out.a.size() && out.finite; i++)
    if (!std::isfinite(out.a[i]) || !std::isfinite(out.b[i]) || !std::isfinite(out.c[i])) out.finite = false;

  // ---- eps arrays: trunc((pixel - X[0])) per column / per row (mandelbrot.cpp:155-159) ------------
  out.eps_re.resize(v.nc); out.eps_re_m.resize(v.nc); out.eps_re_e.resize(v.nc); out.eps_re_lo.resize(v.nc);
  out.eps_im.resize(v.nr); out.eps_im_m.resize(v.nr); out.eps_im_e.resize(v.nr); out.eps_im_lo.resize(v.nr);
  {
    Mp p(P), y(P), hi(64), lo(P);
    long ex = 0;
    for (int c = 0; c < v.nc; c++) {
      pixel_re(v, c, tmp.v, p.v);
      mpf_sub(y.v, p.v, x0re.v);
      out.eps_re[c] = mpf_get_d(y.v);
      out.eps_re_m[c] = mpf_get_d_2exp(&ex, y.v); out.eps_re_e[c] = (int32_t)ex;
      mpf_set_d(hi.v, std::isfinite(out.eps_re[c]) ? out.eps_re[c] : 0.0);
      mpf_sub(lo.v, y.v, hi.v);
      out.eps_re_lo[c] = mpf_get_d(lo.v);
    }
    for (int r = 0; r < v.nr; r++) {
      pixel_im(v, r, tmp.v, p.v);
      mpf_sub(y.v, p.v, x0im.v);
      out.eps_im[r] = mpf_get_d(y.v);
      out.eps_im_m[r] = mpf_get_d_2exp(&ex, y.v); out.eps_im_e[r] = (int32_t)ex;
      mpf_set_d(hi.v, std::isfinite(out.eps_im[r]) ? out.eps_im[r] : 0.0);
      mpf_sub(lo.v, y.v, hi.v);
      out.eps_im_lo[r] = mpf_get_d(lo.v);
    }
    mpf_get_d_2exp(&ex, v.sz_re);
    out.pitch_exp = (int)ex;
  }
}

bool host_mpf_layout_ok() { return pooled_mpf_layout_ok(); }
}  // namespace newman_b200
