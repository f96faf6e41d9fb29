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

typedef struct {
  double er, ei, e2r, e2i, e3r, e3i;
  const op_tables* t;
} series_t;

static void series_init(series_t* s, const op_tables* t, double er, double ei) {
  s->t = t; s->er = er; s->ei = ei;
  s->e2r = er * er - ei * ei;  /* eps2 = sq(eps), mandelbrot.cpp:160 */
  s->e2i = 2.0 * er * ei;
  cplx e3 = cmul(er, ei, s->e2r, s->e2i); /* eps3 = eps * eps2, :161 */
  s->e3r = e3.re; s->e3i = e3.im;
}

static cplx series_d(const series_t* s, int j) { /* d[j], mandelbrot.cpp:164, 166-172, 180 */
  cplx r;
  if (j == 0) { r.re = s->er; r.im = s->ei; return r; }
  const op_tables* t = s->t;
  cplx a = cmul(t->a[2 * j], t->a[2 * j + 1], s->er, s->ei);
  cplx b = cmul(t->b[2 * j], t->b[2 * j + 1], s->e2r, s->e2i);
  cplx c = cmul(t->c[2 * j], t->c[2 * j + 1], s->e3r, s->e3i);
  r.re = (a.re + b.re) + c.re;
  r.im = (a.im + b.im) + c.im;
  return r;
}

static int series_scan(const series_t* s, uint64_t* evals) { /* mandelbrot.cpp:165-181 -> d.size() */
  const op_tables* t = s->t;
  for (int i = 1; i < t->M; i++) {
    cplx b = cmul(t->b[2 * i], t->b[2 * i + 1], s->e2r, s->e2i);
    cplx c = cmul(t->c[2 * i], t->c[2 * i + 1], s->e3r, s->e3i);
    double bmag = b.re * b.re + b.im * b.im; /* isUnstable, :138-142 */
    double cmag = c.re * c.re + c.im * c.im;
    (*evals)++;
    if (bmag * t->tol < cmag) {
      int good = i - 3;
      if (good < 1) good = 1;
      return good;
    }
  }
  return t->M;
}

int oraclep_series_L(const op_tables* t, double er, double ei, double* d_re, double* d_im) {
  series_t s; uint64_t ev = 0;
  series_init(&s, t, er, ei);
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

int64_t oraclep_render_deep(const op_tables* t, const double* eps_re, int nc, const double* eps_im, int nr,
                            int cardioid_mode, const uint8_t* mask, const int32_t* pix_list, int64_t n_list,
                            int mode, op_escape* out, int32_t* rq_pix, int32_t* rq_iter, op_stats* st) {
  const int M = t->M, N = t->N;
  const int Jmax = M + (t->has_escape ? 1 : 0);
  const int64_t W = pix_list ? n_list : (int64_t)nr * nc;
  int64_t n_rq = 0;
  uint64_t executed = 0, evals = 0, skipped = 0, rebased = 0;

  /* Z[0] = 0, Z[j] = X[j-1]; gb[j] = glitch_tol * |Z[j]|^2 (0 at j = 0 and at the escaped iterate) */
  double* Z = (double*)calloc((size_t)2 * (Jmax + 1), sizeof(double));
  double* gb = (double*)calloc((size_t)(Jmax + 1), sizeof(double));
  memcpy(Z + 2, t->x_hi, (size_t)2 * Jmax * sizeof(double));
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
    series_init(&s, t, eps_re[c], eps_im[r]);
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
    cplx d0 = series_d(&s, found);
    double dr = d0.re, di = d0.im;
    const double er = s.er, ei = s.ei;
    int j = L, off = -1;
    for (;;) {
      double xr = Z[2 * j], xi = Z[2 * j + 1];
      double wr = fma(2.0, xr, dr);
      double wi = fma(2.0, xi, di);
      double ndr = fma(-di, wi, fma(dr, wr, er));
      double ndi = fma(di, wr, fma(dr, wi, ei));
      dr = ndr; di = ndi;
      ++j;
      executed++;
      double zr = Z[2 * j] + dr;
      double zi = Z[2 * j + 1] + di;
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
        double dmag = fma(di, di, dr * dr);
        rebase = zmag < dmag;
      }
      if (j + off + 1 >= N) { e->iterations = N; e->smoothing = 0.0f; break; } /* :226-228 */
      if (j == Jmax) rebase = 1; /* outlived the reference orbit */
      if (rebase) { rebased++; off = j + off; j = 0; dr = zr; di = zi; }
    }
  }
  free(Z); free(gb);
  if (st) {
    st->executed_iters += executed; st->series_evals += evals; st->skipped_pixels += skipped;
    st->glitched += (uint64_t)n_rq; st->rebased += rebased;
  }
  return n_rq;
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
