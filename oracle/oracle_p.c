/* Oracle-P — see oracle_p.h. TEST INFRASTRUCTURE ONLY; parity PINNED against Oracle-R.
 * Build: gcc -std=c11 -O2 -ffp-contract=off (oracle/Makefile). No FMA is contracted implicitly; the
 * only fused operations are the explicit fma() calls of the perturbation step, which mirror the
 * explicit __fma_rn of newman_b200/csrc/k3_perturb.cuh one for one.
 */
#include "oracle_p.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BAILOUT 1024.0
#define BAILOUT2 (BAILOUT * BAILOUT) /* mandelbrot.cpp:58-59 */

/* mandelbrot.cpp:133-136; narrowed to float on store (grid.h:11) */
float oraclep_smoothing(double r2) { return (float)(1.0 - log2(0.5 * log(r2) / log(BAILOUT))); }

/* ---- plain-double path (mandelbrot.cpp:231-254, complex.h:19-23) ------------------------------ */
static int in_cardioid_ld(double x, double y) { /* mandelbrot.cpp:63-71 at 64-bit mantissa */
  long double X = x, Y = y, fourth = 0.25L;
  long double xmf = X - fourth, y2 = Y * Y, q = xmf * xmf + y2;
  if (q * (q + xmf) < fourth * y2) return 1;
  q = X + 1.0L;
  return q * q + y2 < fourth * fourth;
}

void oraclep_render_hw(const double* c_re, int nc, const double* c_im, int nr, int N,
                       const uint8_t* in_cardioid, op_escape* out, op_stats* st) {
  uint64_t executed = 0, skipped = 0;
  for (int r = 0; r < nr; r++)
    for (int c = 0; c < nc; c++) {
      op_escape* e = &out[(size_t)r * nc + c];
      double z0r = c_re[c], z0i = c_im[r];
      int inside = in_cardioid ? in_cardioid[(size_t)r * nc + c] : in_cardioid_ld(z0r, z0i);
      if (inside) { e->iterations = N; e->smoothing = 0.0f; skipped++; continue; }
      double zr = z0r, zi = z0i, mag = 0.0;
      int it;
      for (it = 0; it < N; it++) {
        double nr_ = zr * zr - zi * zi;      /* sq(): a.re*a.re - a.im*a.im */
        double ni_ = 2.0 * zr * zi;          /* 2.0 * a.re * a.im          */
        zr = nr_ + z0r; zi = ni_ + z0i;      /* operator+                  */
        mag = zr * zr + zi * zi;             /* sqMag                      */
        executed++;
        if (mag > BAILOUT2) break;
      }
      e->iterations = it;
      e->smoothing = it < N ? oraclep_smoothing(mag) : 0.0f;
    }
  if (st) { st->executed_iters += executed; st->skipped_pixels += skipped; }
}

/* ---- deep path ---------------------------------------------------------------------------------- */
typedef struct { double re, im; } cplx;
static inline cplx cmul(double ar, double ai, double br, double bi) { /* complex.h:29-31 */
  cplx r; r.re = ar * br - ai * bi; r.im = ar * bi + ai * br; return r;
}

/* h + l rounded toward zero (what CUDA's __dadd_rz returns) */
static double add_rz(double h, double l) {
  double s = h + l;
  double bb = s - h;
  double err = (h - (s - bb)) + (l - bb); /* h + l == s + err exactly */
  if ((s > 0.0 && err < 0.0) || (s < 0.0 && err > 0.0)) s = nextafter(s, 0.0);
  return s;
}

/* trunc53(hi + lo + d): the double view of `X[found] + d[found]` formed in mpf and descended
 * (mandelbrot.cpp:184-186, 61; complex.h:33-35), with X carried as hi + lo. */
double oraclep_trunc_add3(double hi, double lo, double d) {
  double s = hi + d;
  double bb = s - hi;
  double e = (hi - (s - bb)) + (d - bb);
  double t = e + lo;
  double h = s + t;
  double b2 = h - s;
  double l = (s - (h - b2)) + (t - b2);
  return add_rz(h, l);
}

/* ---- floatexp: double mantissa + int exponent (mirrors newman_b200/csrc/floatexp.cuh one for one) ----
 * Every operation is ONE IEEE double operation on exactly scaled operands, so wherever the same
 * computation in plain double neither overflows nor underflows the result is bit-identical to it. */
typedef struct { double m; int e; } fe_t;      /* value = m * 2^e; m == 0 or 1 <= |m| < 2 */
typedef struct { fe_t re, im; } fec_t;

static fe_t fe_norm(double m, int e) {
  fe_t r;
  if (m == 0.0 || m != m) { r.m = m; r.e = 0; return r; }
  int k;
  double f = frexp(m, &k); /* f in [0.5, 1) */
  r.m = f * 2.0; r.e = e + k - 1;
  return r;
}
static fe_t fe_from_double(double x) { return fe_norm(x, 0); }
static double fe_scale(double m, int k) { /* m * 2^k, k <= 0; flushed to (signed) zero below 2^-1000 */
  if (k < -1000) return 0.0 * m;
  return ldexp(m, k);
}
static double fe_to_double(fe_t a) { /* no denormal results: flushed to (signed) zero below 2^-1022 */
  if (a.m == 0.0) return a.m;
  if (a.e > 1023) return a.m > 0 ? INFINITY : -INFINITY;
  if (a.e < -1022) return 0.0 * a.m;
  return ldexp(a.m, a.e);
}
static fe_t fe_mul(fe_t a, fe_t b) { return fe_norm(a.m * b.m, a.e + b.e); }
static fe_t fe_neg(fe_t a) { a.m = -a.m; return a; }
static fe_t fe_add(fe_t a, fe_t b) {
  if (a.m == 0.0) return b;
  if (b.m == 0.0) return a;
  if (a.e >= b.e) return fe_norm(a.m + fe_scale(b.m, b.e - a.e), a.e);
  return fe_norm(fe_scale(a.m, a.e - b.e) + b.m, b.e);
}
static fe_t fe_sub(fe_t a, fe_t b) { return fe_add(a, fe_neg(b)); }
static int fe_lt_nonneg(fe_t a, fe_t b) {
  if (b.m == 0.0) return 0;
  if (a.m == 0.0) return 1;
  if (a.e != b.e) return a.e < b.e;
  return a.m < b.m;
}
static fec_t fec_mul(fec_t a, fec_t b) { /* complex.h:29-31 */
  fec_t r;
  r.re = fe_sub(fe_mul(a.re, b.re), fe_mul(a.im, b.im));
  r.im = fe_add(fe_mul(a.re, b.im), fe_mul(a.im, b.re));
  return r;
}
static fec_t fec_sq(fec_t a) { /* complex.h:19-21 */
  fec_t r;
  r.re = fe_sub(fe_mul(a.re, a.re), fe_mul(a.im, a.im));
  fe_t two_re = a.re; two_re.e += (a.re.m != 0.0);
  r.im = fe_mul(two_re, a.im);
  return r;
}
static fec_t fec_add(fec_t a, fec_t b) { fec_t r; r.re = fe_add(a.re, b.re); r.im = fe_add(a.im, b.im); return r; }
static fe_t fec_sqmag(fec_t a) { return fe_add(fe_mul(a.re, a.re), fe_mul(a.im, a.im)); }

/* ---- series phase ------------------------------------------------------------------------------- */
typedef struct {
  double er, ei, e2r, e2i, e3r, e3i; /* double mode */
  fec_t eps, e2, e3;                 /* floatexp mode */
  int fe;
  const op_tables* t;
} series_t;

static int tables_fe(const op_tables* t) { return t->a_exp && t->b_exp && t->c_exp; }

static fec_t load_fe(const double* m, const int32_t* e, int i) {
  fec_t r; r.re = fe_norm(m[2 * i], e[2 * i]); r.im = fe_norm(m[2 * i + 1], e[2 * i + 1]); return r;
}

/* eps: the pixel's offset as floatexp; (er, ei): the same as plain doubles (what the double mode uses) */
static void series_init(series_t* s, const op_tables* t, fec_t eps, double er, double ei) {
  s->t = t; s->fe = tables_fe(t);
  s->eps = eps;
  s->er = er; s->ei = ei;
  if (s->fe) {
    s->e2 = fec_sq(eps);
    s->e3 = fec_mul(eps, s->e2);
    return;
  }
  s->e2r = er * er - ei * ei;  /* eps2 = sq(eps), mandelbrot.cpp:160 */
  s->e2i = 2.0 * er * ei;
  cplx e3 = cmul(er, ei, s->e2r, s->e2i); /* eps3 = eps * eps2, :161 */
  s->e3r = e3.re; s->e3i = e3.im;
}

static fec_t series_d_fe(const series_t* s, int j) { /* d[j], mandelbrot.cpp:164, 166-172, 180 */
  const op_tables* t = s->t;
  fec_t r;
  if (j == 0) return s->eps;
  if (s->fe) {
    fec_t a = fec_mul(load_fe(t->a, t->a_exp, j), s->eps);
    fec_t b = fec_mul(load_fe(t->b, t->b_exp, j), s->e2);
    fec_t c = fec_mul(load_fe(t->c, t->c_exp, j), s->e3);
    return fec_add(fec_add(a, b), c);
  }
  cplx a = cmul(t->a[2 * j], t->a[2 * j + 1], s->er, s->ei);
  cplx b = cmul(t->b[2 * j], t->b[2 * j + 1], s->e2r, s->e2i);
  cplx c = cmul(t->c[2 * j], t->c[2 * j + 1], s->e3r, s->e3i);
  r.re = fe_from_double((a.re + b.re) + c.re);
  r.im = fe_from_double((a.im + b.im) + c.im);
  return r;
}

static cplx series_d(const series_t* s, int j) {
  if (j == 0 && !s->fe) { cplx r0; r0.re = s->er; r0.im = s->ei; return r0; }
  fec_t d = series_d_fe(s, j);
  cplx r; r.re = fe_to_double(d.re); r.im = fe_to_double(d.im);
  return r;
}

static int series_unstable(const series_t* s, int i) { /* isUnstable, mandelbrot.cpp:138-142 */
  const op_tables* t = s->t;
  if (s->fe) {
    fe_t bmag = fec_sqmag(fec_mul(load_fe(t->b, t->b_exp, i), s->e2));
    fe_t cmag = fec_sqmag(fec_mul(load_fe(t->c, t->c_exp, i), s->e3));
    return fe_lt_nonneg(fe_mul(bmag, fe_from_double(t->tol)), cmag);
  }
  cplx b = cmul(t->b[2 * i], t->b[2 * i + 1], s->e2r, s->e2i);
  cplx c = cmul(t->c[2 * i], t->c[2 * i + 1], s->e3r, s->e3i);
  double bmag = b.re * b.re + b.im * b.im;
  double cmag = c.re * c.re + c.im * c.im;
  return bmag * t->tol < cmag;
}

static int series_scan(const series_t* s, uint64_t* evals) { /* mandelbrot.cpp:165-181 -> d.size() */
  const op_tables* t = s->t;
  for (int i = 1; i < t->M; i++) {
    (*evals)++;
    if (series_unstable(s, i)) {
      int good = i - 3;
      if (good < 1) good = 1;
      return good;
    }
  }
  return t->M;
}

static fec_t eps_of(const op_tables* t, const double* eps_re, const double* eps_im, int r, int c) {
  fec_t e;
  if (t->eps_re_exp && t->eps_im_exp) { /* mantissa + exponent as mpf_get_d_2exp returns them */
    e.re = fe_norm(eps_re[c], t->eps_re_exp[c]);
    e.im = fe_norm(eps_im[r], t->eps_im_exp[r]);
  } else {
    e.re = fe_from_double(eps_re[c]);
    e.im = fe_from_double(eps_im[r]);
  }
  return e;
}

int oraclep_series_L(const op_tables* t, double er, double ei, double* d_re, double* d_im) {
  series_t s; uint64_t ev = 0;
  fec_t eps; eps.re = fe_from_double(er); eps.im = fe_from_double(ei);
  series_init(&s, t, eps, er, ei);
  int L = series_scan(&s, &ev);
  cplx d = series_d(&s, L - 1);
  if (d_re) *d_re = d.re;
  if (d_im) *d_im = d.im;
  return L;
}

static double bailed_mag(const series_t* s, int j, double* yr, double* yi) {
  const op_tables* t = s->t;
  cplx d = series_d(s, j);
  *yr = oraclep_trunc_add3(t->x_hi[2 * j], t->x_lo[2 * j], d.re);
  *yi = oraclep_trunc_add3(t->x_hi[2 * j + 1], t->x_lo[2 * j + 1], d.im);
  return *yr * *yr + *yi * *yi;
}

/* ---- scaled perturbation state (mirrors newman_b200/csrc/k3_scaled.cuh) ---------------------------
 * delta = (dr, di) * 2^e. e == 0: a plain state, iterated exactly as before. e != 0: "scaled" state
 * with max(|dr|, |di|) in [1, 2) at every re-normalisation point; used while |delta| is below what a
 * double product can hold (pixel pitch < ~1e-150: delta*delta underflows; < 1e-308: delta itself).
 * The step is the same expression for both (S = 2^e, S == 1 for plain states):
 *      w  = fma(S, d, 2*Z[j])           (== fma(2, Z[j], d) bit for bit when S == 1)
 *      d' = fma(-+di, wi, fma(dr, wr, eps/2^e))
 *      z  = fma(S, d', Z[j+1])          (== Z[j+1] + d' when S == 1)
 * States are re-normalised when created and before the step from every index j = 0 (mod 64). */
#define E_TO_PLAIN (-300) /* a normalised state with exponent above this becomes plain */
#define E_TO_SCALED (-400) /* a plain state whose larger component is below 2^this becomes scaled */
#define RENORM_MASK 63

typedef struct { double dr, di; int e; } pstate;

static double pow2d(int e) { /* 2^e, flushed to zero below the normal range */
  if (e < -1022) return 0.0;
  if (e > 1023) return INFINITY;
  return ldexp(1.0, e);
}

static pstate state_from_fec(fec_t d) {
  pstate s;
  if (d.re.m == 0.0 && d.im.m == 0.0) { s.dr = d.re.m; s.di = d.im.m; s.e = 0; return s; }
  int E = d.re.m == 0.0 ? d.im.e : (d.im.m == 0.0 ? d.re.e : (d.re.e > d.im.e ? d.re.e : d.im.e));
  if (E > E_TO_PLAIN) { s.dr = fe_to_double(d.re); s.di = fe_to_double(d.im); s.e = 0; return s; }
  s.dr = d.re.m == 0.0 ? d.re.m : fe_scale(d.re.m, d.re.e - E);
  s.di = d.im.m == 0.0 ? d.im.m : fe_scale(d.im.m, d.im.e - E);
  s.e = E;
  return s;
}

static void state_renorm(pstate* s) {
  if (s->e == 0) {
    double m = fmax(fabs(s->dr), fabs(s->di));
    if (!(m < ldexp(1.0, E_TO_SCALED)) || m == 0.0) return;
  }
  fec_t d; d.re = fe_norm(s->dr, s->e); d.im = fe_norm(s->di, s->e);
  *s = state_from_fec(d);
}

static double eps_scaled(fe_t eps, double eps0, int e) { /* eps / 2^e as a double; eps0: its plain value */
  if (e == 0) return eps0;
  if (eps.m == 0.0) return eps.m;
  int k = eps.e - e;
  if (k < -1000) return 0.0 * eps.m;
  if (k > 1000) return eps.m > 0 ? INFINITY : -INFINITY;
  return ldexp(eps.m, k);
}

int64_t oraclep_render_deep(const op_tables* t, const double* eps_re, int nc, const double* eps_im, int nr,
                            int cardioid_mode, const uint8_t* mask, const int32_t* pix_list, int64_t n_list,
                            int mode, op_escape* out, int32_t* rq_pix, int32_t* rq_iter, op_stats* st) {
  const int M = t->M, N = t->N;
  const int Jmax = M + (t->has_escape ? 1 : 0);
  const int64_t W = pix_list ? n_list : (int64_t)nr * nc;
  int64_t n_rq = 0;
  uint64_t executed = 0, evals = 0, skipped = 0, rebased = 0;
  const int scaled = t->eps_re_exp && t->eps_im_exp; /* floatexp eps => scaled delta states */

  /* Z[0] = 0, Z[j] = X[j-1]; gb[j] = glitch_tol * |Z[j]|^2 (0 at j = 0 and at the escaped iterate) */
  double* Z = (double*)calloc((size_t)2 * (Jmax + 1), sizeof(double));
  double* gb = (double*)calloc((size_t)(Jmax + 1), sizeof(double));
  memcpy(Z + 2, t->x_hi, (size_t)2 * Jmax * sizeof(double));
  /* The perturbation loop pairs delta with the orbit rounded TO NEAREST: Z[j] = RN(x_hi + x_lo) (the escaped iterate
   * X[M] has no low part and stays as it is). x_hi is the TRUNCATED double (descend, complex.h:33-35: what phases 1-2
   * need); iterating against it biases every factor 2Z + delta by up to an ulp in the same direction, and the relative
   * error of delta then grows linearly with the iteration count (5e-13 after 7 000 iterations of cfg2, against 5e-15
   * for unbiased roundings) — enough to move the escape count of the ~1 % of samples whose last few hundred iterations
   * are chaotic (tests/golden/k3_truth.json: adjudicated against the reference's own continuation at 2-4x its precision). */
  if (!(t->flags & 1)) /* (flag: the exact mode's probe rendering keeps the truncated orbit) */
    for (int i = 0; i < 2 * M; i++) Z[2 + i] = t->x_hi[i] + t->x_lo[i];
  for (int j = 1; j <= Jmax; j++) {
    if (t->has_escape && j == Jmax) { gb[j] = 0.0; continue; }
    gb[j] = (Z[2 * j] * Z[2 * j] + Z[2 * j + 1] * Z[2 * j + 1]) * t->glitch_tol;
  }

  for (int64_t w = 0; w < W; w++) {
    int pix = pix_list ? pix_list[w] : (int)w;
    op_escape* e = &out[pix];
    if (cardioid_mode == 1 || (cardioid_mode == 2 && mask[pix])) { /* mandelbrot.cpp:149-153 */
      e->iterations = N; e->smoothing = 0.0f; skipped++; continue;
    }
    int r = pix / nc, c = pix - r * nc;
    series_t s;
    const fec_t eps = eps_of(t, eps_re, eps_im, r, c);
    /* eps of a plain (e == 0) state: the host's double, or the floatexp value flushed into double range */
    const double er0 = scaled ? fe_to_double(eps.re) : eps_re[c];
    const double ei0 = scaled ? fe_to_double(eps.im) : eps_im[r];
    series_init(&s, t, eps, er0, ei0);
    int L = series_scan(&s, &evals);

    /* phase 2, mandelbrot.cpp:183-207 */
    int found = L - 1;
    double yr, yi;
    if (bailed_mag(&s, found, &yr, &yi) > BAILOUT2) {
      int low = 0, high = L - 1, mid = L / 2;
      while (low <= high) {
        if (!(bailed_mag(&s, mid, &yr, &yi) > BAILOUT2)) low = mid + 1;
        else { high = mid - 1; found = mid; }
        mid = (low + high) / 2;
      }
      double mag = bailed_mag(&s, found, &yr, &yi);
      e->iterations = found;
      e->smoothing = oraclep_smoothing(mag);
      continue;
    }
    if (L >= N) { e->iterations = N; e->smoothing = 0.0f; continue; } /* :212 loop is empty */

    /* phase 3 as perturbation: state (j, delta) pairs delta with Z[j]; it = j + off */
    pstate ps;
    if (scaled) ps = state_from_fec(series_d_fe(&s, found));
    else { cplx d0 = series_d(&s, found); ps.dr = d0.re; ps.di = d0.im; ps.e = 0; }
    double S = pow2d(ps.e), er = eps_scaled(eps.re, er0, ps.e), ei = eps_scaled(eps.im, ei0, ps.e);
    int j = L, off = -1;
    for (;;) {
      if (scaled && (j & RENORM_MASK) == 0) {
        state_renorm(&ps);
        S = pow2d(ps.e); er = eps_scaled(eps.re, er0, ps.e); ei = eps_scaled(eps.im, ei0, ps.e);
      }
      double dr = ps.dr, di = ps.di;
      double xr = Z[2 * j], xi = Z[2 * j + 1];
      double wr = fma(S, dr, 2.0 * xr);   /* == fma(2.0, xr, dr) when S == 1 */
      double wi = fma(S, di, 2.0 * xi);
      double ndr = fma(-di, wi, fma(dr, wr, er));
      double ndi = fma(di, wr, fma(dr, wi, ei));
      ps.dr = dr = ndr; ps.di = di = ndi;
      ++j;
      executed++;
      double zr = fma(S, dr, Z[2 * j]);   /* == Z + d when S == 1 */
      double zi = fma(S, di, Z[2 * j + 1]);
      double zmag = fma(zi, zi, zr * zr);
      if (zmag > BAILOUT2) { /* bailedOut, :61, :216-219 */
        e->iterations = j + off;
        e->smoothing = oraclep_smoothing(zr * zr + zi * zi);
        break;
      }
      int rebase = 0;
      if (mode == 0) {
        if (j != Jmax && zmag < gb[j]) { /* glitch: flag and re-queue */
          rq_pix[n_rq] = pix; rq_iter[n_rq] = j + off; n_rq++;
          e->iterations = -1; e->smoothing = 0.0f;
          break;
        }
      } else {
        double tr = S * dr, ti = S * di; /* delta itself (exact when S == 1) */
        double dmag = fma(ti, ti, tr * tr);
        rebase = zmag < dmag;
      }
      if (j + off + 1 >= N) { e->iterations = N; e->smoothing = 0.0f; break; } /* :226-228 */
      if (j == Jmax) rebase = 1; /* outlived the reference orbit */
      if (rebase) {
        rebased++; off = j + off; j = 0;
        ps.dr = zr; ps.di = zi; ps.e = 0;
        S = 1.0; er = er0; ei = ei0;
      }
    }
  }
  free(Z); free(gb);
  if (st) {
    st->executed_iters += executed; st->series_evals += evals; st->skipped_pixels += skipped;
    st->glitched += (uint64_t)n_rq; st->rebased += rebased;
  }
  return n_rq;
}

/* ---- double-double continuation ("exact mode", newman_b200/csrc/k3_dd.cuh mirrors it one for one) ----------------
 * FP64 perturbation carries a relative error of ~5e-15 in delta after a few thousand iterations, and samples whose last
 * few hundred iterations are chaotic amplify that past 1. For the samples listed — the ones whose
 * count depends on how the orbit table was rounded — phase 3 is repeated with delta, eps and the orbit in double-double
 * (~106 bits): delta' = delta*(2Z + delta) + eps, z' = Z' + delta', escape test on the high parts, rebasing onto Z[0] = 0
 * when the sample outlives the orbit. Phases 1-2 are the reference's double arithmetic and unchanged.
 * Every operation below is a plain double operation or an explicit fma: no contraction (-ffp-contract=off). */
typedef struct { double hi, lo; } dd_t;
static dd_t dd_fast2sum(double a, double b) { dd_t r; r.hi = a + b; r.lo = b - (r.hi - a); return r; }
static dd_t dd_2sum(double a, double b) {
  dd_t r; r.hi = a + b; double bb = r.hi - a; r.lo = (a - (r.hi - bb)) + (b - bb); return r;
}
static dd_t dd_add(dd_t x, dd_t y) {
  dd_t s = dd_2sum(x.hi, y.hi);
  dd_t t = dd_2sum(x.lo, y.lo);
  s.lo += t.hi;
  s = dd_fast2sum(s.hi, s.lo);
  s.lo += t.lo;
  return dd_fast2sum(s.hi, s.lo);
}
static dd_t dd_mul(dd_t x, dd_t y) {
  dd_t p; p.hi = x.hi * y.hi; p.lo = fma(x.hi, y.hi, -p.hi);
  p.lo += x.hi * y.lo;
  p.lo += x.lo * y.hi;
  return dd_fast2sum(p.hi, p.lo);
}
static dd_t dd_neg(dd_t x) { x.hi = -x.hi; x.lo = -x.lo; return x; }
static dd_t dd_dbl(dd_t x) { x.hi *= 2.0; x.lo *= 2.0; return x; }   /* exact */

int64_t oraclep_refine_dd(const op_tables* t, const double* eps_re, const double* eps_re_lo, int nc, const double* eps_im,
                          const double* eps_im_lo, int nr, const int32_t* pix_list, int64_t n_list, op_escape* out,
                          op_stats* st) {
  const int M = t->M, N = t->N;
  const int Jmax = M + (t->has_escape ? 1 : 0);
  uint64_t executed = 0, evals = 0, rebased = 0;
  (void)nr;
  if (t->eps_re_exp || t->eps_im_exp) return -1; /* scaled frames are not refined */
  for (int64_t w = 0; w < n_list; w++) {
    const int pix = pix_list[w];
    op_escape* e = &out[pix];
    const int r = pix / nc, c = pix - r * nc;
    series_t s;
    const fec_t eps = eps_of(t, eps_re, eps_im, r, c);
    series_init(&s, t, eps, eps_re[c], eps_im[r]);
    int L = series_scan(&s, &evals);
    int found = L - 1;
    double yr, yi;
    if (bailed_mag(&s, found, &yr, &yi) > BAILOUT2) continue; /* phase 2 decided it: nothing to refine */
    if (L >= N) continue;
    cplx d0 = series_d(&s, found);
    dd_t dr = {d0.re, 0.0}, di = {d0.im, 0.0};
    const dd_t er = {eps_re[c], eps_re_lo ? eps_re_lo[c] : 0.0}, ei = {eps_im[r], eps_im_lo ? eps_im_lo[r] : 0.0};
    int j = L, off = -1;
    for (;;) {
      /* Z[j] = X[j-1] as hi + lo (Z[0] = 0; the escaped iterate X[M] has no low part) */
      dd_t xr = {0.0, 0.0}, xi = {0.0, 0.0};
      if (j >= 1) {
        xr.hi = t->x_hi[2 * (j - 1)]; xi.hi = t->x_hi[2 * (j - 1) + 1];
        if (j - 1 < M) { xr.lo = t->x_lo[2 * (j - 1)]; xi.lo = t->x_lo[2 * (j - 1) + 1]; }
      }
      const dd_t wr = dd_add(dd_dbl(xr), dr), wi = dd_add(dd_dbl(xi), di);
      const dd_t ndr = dd_add(dd_add(dd_mul(dr, wr), dd_neg(dd_mul(di, wi))), er);
      const dd_t ndi = dd_add(dd_add(dd_mul(dr, wi), dd_mul(di, wr)), ei);
      dr = ndr; di = ndi;
      ++j;
      executed++;
      dd_t yr2 = {t->x_hi[2 * (j - 1)], 0.0}, yi2 = {t->x_hi[2 * (j - 1) + 1], 0.0};
      if (j - 1 < M) { yr2.lo = t->x_lo[2 * (j - 1)]; yi2.lo = t->x_lo[2 * (j - 1) + 1]; }
      const dd_t zr = dd_add(yr2, dr), zi = dd_add(yi2, di);
      const double zmag = fma(zi.hi, zi.hi, zr.hi * zr.hi);
      if (zmag > BAILOUT2) {
        e->iterations = j + off;
        e->smoothing = oraclep_smoothing(zr.hi * zr.hi + zi.hi * zi.hi);
        break;
      }
      if (j + off + 1 >= N) { e->iterations = N; e->smoothing = 0.0f; break; }
      if (j == Jmax) { rebased++; off = j + off; j = 0; dr = zr; di = zi; }
    }
  }
  if (st) { st->executed_iters += executed; st->series_evals += evals; st->rebased += rebased; }
  return 0;
}

int64_t oraclep_pick_reference(const int32_t* rq_pix, const int32_t* rq_iter, int64_t n) {
  int64_t best = -1;
  for (int64_t i = 0; i < n; i++)
    if (best < 0 || rq_iter[i] < rq_iter[best] || (rq_iter[i] == rq_iter[best] && rq_pix[i] < rq_pix[best])) best = i;
  return best;
}

/* ---- colour resolve (viewer.cpp:84-124; conventions of k4_resolve.cuh) --------------------------- */
static float clampf(float x) { return x < 0.0f ? 0.0f : (x > 255.0f ? 255.0f : x); }
static void get_color(op_escape e, const uint8_t* pal, int n_pal, int N, int smooth, float* rgb) {
  if (e.iterations >= N || e.iterations < 0) { rgb[0] = rgb[1] = rgb[2] = 0.0f; return; }
  int i1 = e.iterations < n_pal ? e.iterations : n_pal - 1;
  const uint8_t* c1 = pal + 3 * i1;
  if (!smooth) { rgb[0] = c1[0]; rgb[1] = c1[1]; rgb[2] = c1[2]; return; }
  int i0 = i1 > 0 ? i1 - 1 : 0;
  const uint8_t* c0 = pal + 3 * i0;
  float t = e.smoothing, u = 1.0f - t;
  for (int k = 0; k < 3; k++) rgb[k] = truncf(clampf(u * c0[k] + t * c1[k]));
}

void oraclep_resolve(const op_escape* grid, int nr, int nc, const uint8_t* pal, int n_pal, int N, int sc,
                     int smooth, uint8_t* rgb) {
  int onr = nr / sc, onc = nc / sc;
  for (int r = 0; r < onr; r++)
    for (int c = 0; c < onc; c++) {
      float acc[3] = {0, 0, 0}, col[3];
      for (int r1 = r * sc; r1 < (r + 1) * sc; r1++)
        for (int c1 = c * sc; c1 < (c + 1) * sc; c1++) {
          get_color(grid[(size_t)r1 * nc + c1], pal, n_pal, N, smooth, col);
          acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2];
        }
      uint8_t* o = rgb + 3 * ((size_t)r * onc + c);
      for (int k = 0; k < 3; k++) {
        float v = acc[k];
        if (sc > 1) v = truncf(clampf(v / (float)(sc * sc)));
        o[k] = (uint8_t)v;
      }
    }
}

op_escape oraclep_at_sc(const op_escape* grid, int nc, int r, int c, int sc) { /* mandelbrot.cpp:320-332 */
  op_escape e;
  float sum = 0.0f;
  for (int r1 = 0; r1 < sc; r1++)
    for (int c1 = 0; c1 < sc; c1++) {
      e = grid[(size_t)(sc * r + r1) * nc + (sc * c + c1)];
      sum += e.iterations + e.smoothing;
    }
  sum /= sc * sc;
  e.iterations = (int)sum;
  e.smoothing = sum - e.iterations;
  return e;
}

/* ---- multiwave palette (multiwave.cpp:5-17, 75-116) --------------------------------------------- */
static uint8_t to_byte(float v) {
  float x = v * 255.0f + 0.5f;
  if (!(x > 0.0f)) return 0;
  if (x >= 255.0f) return 255;
  return (uint8_t)x;
}
static float hue_channel(float p, float q, float t) {
  if (t < 0.0f) t += 1.0f;
  if (t > 1.0f) t -= 1.0f;
  if (t < 1.0f / 6.0f) return p + (q - p) * 6.0f * t;
  if (t < 0.5f) return q;
  if (t < 2.0f / 3.0f) return p + (q - p) * (2.0f / 3.0f - t) * 6.0f;
  return p;
}
static void hsl2rgb(float h_deg, float s, float l, uint8_t* rgb) {
  float h = fmodf(h_deg, 360.0f);
  if (h < 0.0f) h += 360.0f;
  h /= 360.0f;
  if (s <= 0.0f) { rgb[0] = rgb[1] = rgb[2] = to_byte(l); return; }
  float q = l < 0.5f ? l * (1.0f + s) : l + s - l * s;
  float p = 2.0f * l - q;
  rgb[0] = to_byte(hue_channel(p, q, h + 1.0f / 3.0f));
  rgb[1] = to_byte(hue_channel(p, q, h));
  rgb[2] = to_byte(hue_channel(p, q, h - 1.0f / 3.0f));
}
static void interp3(const uint8_t* a, const uint8_t* b, float t, uint8_t* o) {
  for (int k = 0; k < 3; k++) o[k] = (uint8_t)clampf((1.0f - t) * a[k] + t * b[k]);
}
static float cycle_value(const float* v, int n, int period, int step) { /* FloatCycle::value */
  float t = (size_t)n * (step % period) / (float)period;
  int i0 = (int)t, i1 = (i0 + 1) % n;
  t -= i0;
  return (float)((1.0 - t) * v[i0] + t * v[i1]);
}
void oraclep_palette_cache(int n_cycles, const int* hue_counts, const float* hue_values, const int* hue_periods,
                           int hue_period, int n_sat, const float* sat_values, int sat_period, int n_lum,
                           const float* lum_amp, const int* lum_period, int N, uint8_t* rgb) {
  const float tau = (float)(2.0 * 3.14159265358979);
  for (int i = 0; i < N; i++) {
    float sat = cycle_value(sat_values, n_sat, sat_period, i);
    float lum = 0.0f;
    for (int k = 0; k < n_lum; k++) lum += lum_amp[k] * sinf(i * tau / lum_period[k]);
    lum = (float)(1.0 / (1.0 + exp(-lum)));
    float ty = (size_t)n_cycles * (i % hue_period) / (float)hue_period;
    int y0 = (int)ty, y1 = (y0 + 1) % n_cycles;
    ty -= y0;
    uint8_t col[2][3];
    int ys[2] = {y0, y1};
    for (int w = 0; w < 2; w++) {
      const float* hv = hue_values;
      for (int k = 0; k < ys[w]; k++) hv += hue_counts[k];
      int n = hue_counts[ys[w]], per = hue_periods[ys[w]];
      float tx = (size_t)n * (i % per) / (float)per;
      int x0 = (int)tx, x1 = (x0 + 1) % n;
      tx -= x0;
      uint8_t a[3], b[3];
      hsl2rgb(hv[x0], sat, lum, a);
      hsl2rgb(hv[x1], sat, lum, b);
      interp3(a, b, tx, col[w]);
    }
    interp3(col[0], col[1], ty, rgb + 3 * i);
  }
}

/* ---- zoom-video in-betweening (video.cpp:14-34; conventions of newman_b200/csrc/k5_video.cuh) ----- */
static void vid_sample(const uint8_t* src, int H, int W, int sh, int sw, int r, int c, float* rgb) {
  float fy = (r + 0.5f) * ((float)H / (float)sh) - 0.5f;
  float fx = (c + 0.5f) * ((float)W / (float)sw) - 0.5f;
  fy = fy < 0.0f ? 0.0f : (fy > (float)(H - 1) ? (float)(H - 1) : fy);
  fx = fx < 0.0f ? 0.0f : (fx > (float)(W - 1) ? (float)(W - 1) : fx);
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + 1 < H ? y0 + 1 : H - 1, x1 = x0 + 1 < W ? x0 + 1 : W - 1;
  float wy = fy - (float)y0, wx = fx - (float)x0;
  const uint8_t* p00 = src + 3 * ((size_t)y0 * W + x0);
  const uint8_t* p01 = src + 3 * ((size_t)y0 * W + x1);
  const uint8_t* p10 = src + 3 * ((size_t)y1 * W + x0);
  const uint8_t* p11 = src + 3 * ((size_t)y1 * W + x1);
  for (int k = 0; k < 3; k++) {
    float top = (1.0f - wx) * p00[k] + wx * p01[k];
    float bot = (1.0f - wx) * p10[k] + wx * p11[k];
    rgb[k] = truncf((1.0f - wy) * top + wy * bot);
  }
}

void oraclep_video_inbetween(const uint8_t* prev, const uint8_t* next, int H, int W, int nr, int nc, int rate,
                             float v, uint8_t* out) {
  float sm = 2.0f / 3.0f, lg = 1.0f; /* video.cpp:17; v = pow(1.5, 1.0 / rate) is passed in (see header) */
  for (int i = 0; i < rate; i++) {
    float t = (float)i / (float)rate;
    int lh = (int)(H * lg), lw = (int)(W * lg), sh = (int)(H * sm), sw = (int)(W * sm);
    for (int r = 0; r < nr; r++)
      for (int c = 0; c < nc; c++) {
        float o[3] = {0.0f, 0.0f, 0.0f}, s[3];
        int lr = r - (nr - lh) / 2, lc = c - (nc - lw) / 2;
        if (lr >= 0 && lr < lh && lc >= 0 && lc < lw) vid_sample(prev, H, W, lh, lw, lr, lc, o);
        int sr = r - (nr - sh) / 2, sc = c - (nc - sw) / 2;
        if (sr >= 0 && sr < sh && sc >= 0 && sc < sw) {
          vid_sample(next, H, W, sh, sw, sr, sc, s);
          for (int k = 0; k < 3; k++) o[k] = truncf((1.0f - t) * o[k] + t * s[k]);
        }
        uint8_t* q = out + 3 * (((size_t)i * nr + r) * nc + c);
        q[0] = (uint8_t)o[0]; q[1] = (uint8_t)o[1]; q[2] = (uint8_t)o[2];
      }
    sm *= v; lg *= v;
  }
}
