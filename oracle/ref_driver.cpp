// Oracle-R: headless C-ABI driver around the UNMODIFIED reference translation unit
// /root/reference/mandelbrot.cpp (compiled where it lies, see oracle/Makefile).
//
// TEST INFRASTRUCTURE ONLY. Nothing under newman_b200/ may link or load this; it exists so that
// tests/ and bench.py's cpu_baseline / --impl reference legs can (a) pin the restated oracle
// (oracle/oracle_p.c) and the CUDA path against the reference itself, and (b) time the reference's
// own CPU render on the box's host cores.
//
// The subclass only *reads* protected state (X/A/B/C/grid, mandelbrot.h:9-10) and calls the
// reference's own member functions; no reference arithmetic is re-implemented here except the
// per-pixel epsilon dump, which repeats the expressions of mandelbrot.cpp:86-87,155-159,271,275
// through the same mpf_class operators.
#include "mandelbrot.h"  // resolved via -I/root/reference
#include <chrono>
#include <cstring>
#include <cstdlib>

#define ORACLE_API extern "C" __attribute__((visibility("default")))

namespace {
struct RefView : public Mandelbrot {
  RefView(int nr, int nc) : Mandelbrot(nr, nc) {}
  int orbit_len() const { return (int)X.size(); }
  const std::vector<HPComplex>& tab(int which) const {
    return which == 0 ? X : which == 1 ? A : which == 2 ? B : C;
  }
  const RenderGrid& raster() const { return grid; }
  bool cardioid(const HPComplex& p) { return inCardioid(p); }
  RenderGrid::EscapeValue one(const HPComplex& p) { return useHardware() ? getIterationsHW(p) : getIterations(p); }
  HPComplex pixel(int r, int c) const {
    HPComplex pt;
    pt.im = center.im + (rows() / 2 - r - 1) * sz.im;
    pt.re = center.re + (c - cols() / 2) * sz.re;
    return pt;
  }
  void orbit_from(const HPComplex& p) { computeOrbit(p); }
  void reprec_orbit(mp_bitcnt_t bits) {   // value-preserving when widening, and when narrowing back to the original precision
    for (size_t i = 0; i < X.size(); i++) { X[i].re.set_prec(bits); X[i].im.set_prec(bits); }
  }
  void series() { computeSeries(); }
  void clear_tabs() { X.clear(); A.clear(); B.clear(); C.clear(); }
};
}  // namespace

ORACLE_API void* ref_create(int nr, int nc) { return new RefView(nr, nc); }
ORACLE_API void ref_destroy(void* h) { delete (RefView*)h; }

// SURVEY.md §8d construction order: N; sz from decimal strings; zoom(1.0f) (-> setPrecision);
// THEN centre from decimal strings (parsed at the new precision); tolerance.
ORACLE_API void ref_set_view(void* h, int N, const char* sz_re, const char* sz_im,
                             const char* c_re, const char* c_im, double tol) {
  RefView* v = (RefView*)h;
  v->N = N;
  if (sz_re && sz_im) {
    v->sz.re = sz_re; v->sz.im = sz_im;
    v->zoom(1.0f);
  }
  if (c_re && c_im) { v->center.re = c_re; v->center.im = c_im; }
  v->error_tolerance = tol;
}

ORACLE_API int ref_rows(void* h) { return ((RefView*)h)->rows(); }
ORACLE_API int ref_cols(void* h) { return ((RefView*)h)->cols(); }
ORACLE_API int ref_use_hardware(void* h) { return ((RefView*)h)->useHardware() ? 1 : 0; }
ORACLE_API int ref_precision_bits(void* h) { return (int)((RefView*)h)->center.re.get_prec(); }
ORACLE_API double ref_precompute(void* h) {
  auto t0 = std::chrono::steady_clock::now();
  ((RefView*)h)->precompute();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// Force a specific reference point (pixel r,c) instead of findProbe: lets tests pin orbit/series
// arithmetic on big grids without paying for the reference's exhaustive probe scan.
ORACLE_API void ref_precompute_at(void* h, int r, int c) {
  RefView* v = (RefView*)h;
  v->clear_tabs();
  v->orbit_from(v->pixel(r, c));
  v->series();
}
ORACLE_API int ref_orbit_len(void* h) { return ((RefView*)h)->orbit_len(); }

// Truncating descend (complex.h:33-35) of table `which` (0=X,1=A,2=B,3=C) into out[2*M].
ORACLE_API void ref_dump_table(void* h, int which, double* out) {
  RefView* v = (RefView*)h;
  const std::vector<HPComplex>& t = v->tab(which);
  for (size_t i = 0; i < t.size(); i++) {
    out[2 * i] = t[i].re.get_d();
    out[2 * i + 1] = t[i].im.get_d();
  }
}
// The iterate the reference computes and then drops when its orbit escapes (mandelbrot.cpp:101-108):
// X[M] = X[M-1]^2 + X[0], same mpf expressions. Returns 0 if the orbit ran to N without escaping.
ORACLE_API int ref_orbit_escape(void* h, double* out) {
  RefView* v = (RefView*)h;
  const std::vector<HPComplex>& X = v->tab(0);
  if ((int)X.size() >= v->N) return 0;
  size_t i = X.size();
  HPComplex nx;
  nx.re = X[i - 1].re * X[i - 1].re - X[i - 1].im * X[i - 1].im + X[0].re;
  nx.im = 2.0 * (X[i - 1].re * X[i - 1].im) + X[0].im;
  out[0] = nx.re.get_d(); out[1] = nx.im.get_d();
  return 1;
}
// X[i] as hi + lo doubles (hi = truncation, lo = truncation of the remainder), out[4*M].
ORACLE_API void ref_dump_orbit_dd(void* h, double* out) {
  RefView* v = (RefView*)h;
  const std::vector<HPComplex>& t = v->tab(0);
  for (size_t i = 0; i < t.size(); i++) {
    double hr = t[i].re.get_d(), hi = t[i].im.get_d();
    mpf_class lr = t[i].re - hr, li = t[i].im - hi;
    out[4 * i] = hr; out[4 * i + 1] = lr.get_d();
    out[4 * i + 2] = hi; out[4 * i + 3] = li.get_d();
  }
}
// eps_re[c] / eps_im[r]: truncated doubles of (pixel - X[0]) exactly as mandelbrot.cpp:155-159.
ORACLE_API void ref_dump_eps(void* h, double* eps_re, double* eps_im) {
  RefView* v = (RefView*)h;
  const HPComplex& x0 = v->tab(0)[0];
  for (int c = 0; c < v->cols(); c++) {
    HPComplex p = v->pixel(0, c);
    mpf_class d = p.re - x0.re;
    eps_re[c] = d.get_d();
  }
  for (int r = 0; r < v->rows(); r++) {
    HPComplex p = v->pixel(r, 0);
    mpf_class d = p.im - x0.im;
    eps_im[r] = d.get_d();
  }
}
// Pixel coordinates as truncated doubles (what getIterationsHW iterates, mandelbrot.cpp:234).
ORACLE_API void ref_dump_coords(void* h, double* c_re, double* c_im) {
  RefView* v = (RefView*)h;
  for (int c = 0; c < v->cols(); c++) c_re[c] = v->pixel(0, c).re.get_d();
  for (int r = 0; r < v->rows(); r++) c_im[r] = v->pixel(r, 0).im.get_d();
}
ORACLE_API int ref_in_cardioid(void* h, int r, int c) {
  RefView* v = (RefView*)h;
  return v->cardioid(v->pixel(r, c)) ? 1 : 0;
}

// Render rows listed in rows[0..n) with the reference's computeRow (mandelbrot.cpp:269-283) and
// copy each finished row (8-byte {int32,float32} records, grid.h:8-16) to out + k*cols.
// Returns wall seconds spent inside computeRow.
ORACLE_API double ref_compute_rows(void* h, const int* rows, int n, void* out) {
  RefView* v = (RefView*)h;
  const int nc = v->cols();
  double secs = 0.0;
  for (int k = 0; k < n; k++) {
    auto t0 = std::chrono::steady_clock::now();
    v->computeRow(rows[k]);
    secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out)
      std::memcpy((char*)out + (size_t)k * nc * sizeof(RenderGrid::EscapeValue),
                  &v->raster().values[(size_t)rows[k] * nc], (size_t)nc * sizeof(RenderGrid::EscapeValue));
  }
  return secs;
}
// The per-pixel body of computeRow (mandelbrot.cpp:269-283) for an arbitrary list of pixel ids
// (r*cols+c): same pixel map, same getIterations/getIterationsHW. Lets the timed CPU baseline be a
// bounded, strided sample of a frame whose full render would take hours.
ORACLE_API double ref_compute_pixels(void* h, const int* pix, int n, void* out) {
  RefView* v = (RefView*)h;
  RenderGrid::EscapeValue* o = (RenderGrid::EscapeValue*)out;
  auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < n; k++) {
    int r = pix[k] / v->cols(), c = pix[k] % v->cols();
    o[k] = v->one(v->pixel(r, c));
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
ORACLE_API double ref_render_all(void* h, void* out) {
  RefView* v = (RefView*)h;
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < v->rows(); r++) v->computeRow(r);
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (out) std::memcpy(out, v->raster().values.data(), v->raster().values.size() * sizeof(RenderGrid::EscapeValue));
  return secs;
}
// Multisample rescale paths (mandelbrot.cpp:320-360) for the at()/scaleUp/scaleDown parity tests.
ORACLE_API void ref_scale(void* h, int sc, int up) {
  RefView* v = (RefView*)h;
  if (up) v->scaleUp(sc); else v->scaleDown(sc);
}
ORACLE_API void ref_read_grid(void* h, void* out) {
  RefView* v = (RefView*)h;
  std::memcpy(out, v->raster().values.data(), v->raster().values.size() * sizeof(RenderGrid::EscapeValue));
}
ORACLE_API void ref_write_grid(void* h, const void* in) {
  RefView* v = (RefView*)h;
  std::memcpy((void*)&v->raster().values[0], in, v->raster().values.size() * sizeof(RenderGrid::EscapeValue));
}
ORACLE_API void ref_zoom(void* h, float s) { ((RefView*)h)->zoom(s); }
ORACLE_API void ref_translate(void* h, int dr, int dc, int sc) { ((RefView*)h)->translate(dr, dc, sc); }
ORACLE_API void ref_zoom_at(void* h, float s, int r, int c, int sc) { ((RefView*)h)->zoomAt(s, r, c, sc); }
ORACLE_API int ref_load_legacy(void* h, const char* fn) { ((RefView*)h)->loadLegacy(fn); return 0; }
// Decimal dump of the view (same mpf_out_str form as viewer.cpp:15-21), for transform parity.
ORACLE_API int ref_view_string(void* h, int which, char* buf, int cap) {
  RefView* v = (RefView*)h;
  const mpf_class& f = which == 0 ? v->center.re : which == 1 ? v->center.im : which == 2 ? v->sz.re : v->sz.im;
  mp_exp_t e;
  char* s = mpf_get_str(NULL, &e, 10, 0, f.get_mpf_t());
  int n = snprintf(buf, cap, "%s@%ld", s, (long)e);
  free(s);
  return n;
}

// Brute-force truth for the listed pixels (ids r*cols+c): the pixel coordinate exactly as the reference forms
// it (pixel(), mandelbrot.cpp:271, 275, at the view's precision) is widened to `prec_bits` and iterated
// directly, z <- z^2 + c from z = c, all in mpf at that precision. out_it[k] = index of the first iterate with
// |z|^2 > 2^20 (the deep path's counting convention, mandelbrot.cpp:212-217: iterate 0 is the pixel) or N;
// out_r2[k] = |z|^2 of that iterate (truncated to double). This is NOT the reference's algorithm (its own
// continuation runs at the view's 64-192 bits with truncating operations, SURVEY.md finding 4): it is what the
// reference's and the CUDA path's escape counts are adjudicated against. tests/golden/make_k3_truth.py.
ORACLE_API double ref_truth_pixels(void* h, const int* pix, int n, int prec_bits, int* out_it, double* out_r2) {
  RefView* v = (RefView*)h;
  const mp_bitcnt_t saved = mpf_get_default_prec();
  auto t0 = std::chrono::steady_clock::now();
  mpf_t cr, ci, zr, zi, rr, ii, ri, m;
  mpf_init2(cr, prec_bits); mpf_init2(ci, prec_bits); mpf_init2(zr, prec_bits); mpf_init2(zi, prec_bits);
  mpf_init2(rr, prec_bits); mpf_init2(ii, prec_bits); mpf_init2(ri, prec_bits); mpf_init2(m, prec_bits);
  for (int k = 0; k < n; k++) {
    HPComplex p = v->pixel(pix[k] / v->cols(), pix[k] % v->cols());
    mpf_set(cr, p.re.get_mpf_t()); mpf_set(ci, p.im.get_mpf_t());
    mpf_set(zr, cr); mpf_set(zi, ci);
    int it = v->N;
    double r2 = 0.0;
    for (int i = 1; i < v->N; i++) {
      mpf_mul(rr, zr, zr); mpf_mul(ii, zi, zi); mpf_mul(ri, zr, zi);
      mpf_sub(zr, rr, ii); mpf_add(zr, zr, cr);
      mpf_mul_ui(ri, ri, 2UL); mpf_add(zi, ri, ci);
      // cheap pre-test in double, exact comparison only near the threshold
      const double dr = mpf_get_d(zr), di = mpf_get_d(zi);
      const double mag = dr * dr + di * di;
      if (mag > 1048570.0) {
        mpf_mul(rr, zr, zr); mpf_mul(ii, zi, zi); mpf_add(m, rr, ii);
        if (mpf_cmp_d(m, 1048576.0) > 0) { it = i; r2 = mpf_get_d(m); break; }
      }
    }
    out_it[k] = it;
    if (out_r2) out_r2[k] = r2;
  }
  mpf_clear(cr); mpf_clear(ci); mpf_clear(zr); mpf_clear(zi); mpf_clear(rr); mpf_clear(ii); mpf_clear(ri); mpf_clear(m);
  mpf_set_default_prec(saved);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// The reference's OWN per-pixel algorithm (getIterations, mandelbrot.cpp:144-229) with its phase-3 continuation run at
// `prec_bits` instead of the view's precision. The pixel coordinate is formed at the view's precision exactly as
// computeRow forms it, then widened (value unchanged); the orbit X is widened in place (values unchanged) and narrowed
// back afterwards; phases 1-2 are double arithmetic on descended values and do not depend on the precision. Phase 3
// (`Y = X[L-1] + d[L-1]; Yn = Y^2 + Y0`, 209-224) then runs through the same mpf_class operators at the global default
// precision `prec_bits`: the exact continuation of the reference's own phase-2 state, free of the truncation noise of
// its 64-192-bit arithmetic (SURVEY.md finding 4). Used to adjudicate reference-vs-CUDA escape-count differences.
ORACLE_API double ref_compute_pixels_wide(void* h, const int* pix, int n, int prec_bits, void* out) {
  RefView* v = (RefView*)h;
  RenderGrid::EscapeValue* o = (RenderGrid::EscapeValue*)out;
  const mp_bitcnt_t view_prec = mpf_get_default_prec();
  const mp_bitcnt_t orbit_prec = v->orbit_len() ? v->tab(0)[0].re.get_prec() : view_prec;
  auto t0 = std::chrono::steady_clock::now();
  v->reprec_orbit((mp_bitcnt_t)prec_bits);
  for (int k = 0; k < n; k++) {
    int r = pix[k] / v->cols(), c = pix[k] % v->cols();
    mpf_set_default_prec(view_prec);
    HPComplex p = v->pixel(r, c);
    mpf_set_default_prec((mp_bitcnt_t)prec_bits);
    HPComplex pw;
    pw.re = mpf_class(p.re, (mp_bitcnt_t)prec_bits);
    pw.im = mpf_class(p.im, (mp_bitcnt_t)prec_bits);
    o[k] = v->one(pw);
  }
  mpf_set_default_prec(view_prec);
  v->reprec_orbit(orbit_prec);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Floatexp form of the tables, for views whose coefficients leave double range (pixel pitch < ~1e-97, where the
// reference's own per-pixel code dies with SIGFPE but its orbit and series recurrences are fine): mantissa in
// [0.5, 1) and binary exponent of every component, truncating like descend() — mpf_get_d_2exp. which as ref_dump_table.
ORACLE_API void ref_dump_table_2exp(void* h, int which, double* mant, int* expo) {
  RefView* v = (RefView*)h;
  const std::vector<HPComplex>& t = v->tab(which);
  for (size_t i = 0; i < t.size(); i++) {
    long e;
    mant[2 * i] = mpf_get_d_2exp(&e, t[i].re.get_mpf_t()); expo[2 * i] = (int)e;
    mant[2 * i + 1] = mpf_get_d_2exp(&e, t[i].im.get_mpf_t()); expo[2 * i + 1] = (int)e;
  }
}
ORACLE_API void ref_dump_eps_2exp(void* h, double* mre, int* ere, double* mim, int* eim) {
  RefView* v = (RefView*)h;
  const HPComplex& x0 = v->tab(0)[0];
  long e;
  for (int c = 0; c < v->cols(); c++) {
    HPComplex p = v->pixel(0, c);
    mpf_class d = p.re - x0.re;
    mre[c] = mpf_get_d_2exp(&e, d.get_mpf_t()); ere[c] = (int)e;
  }
  for (int r = 0; r < v->rows(); r++) {
    HPComplex p = v->pixel(r, 0);
    mpf_class d = p.im - x0.im;
    mim[r] = mpf_get_d_2exp(&e, d.get_mpf_t()); eim[r] = (int)e;
  }
}
