/* Oracle-P — CPU restatement (plain C) of newman's per-pixel hot path as this repo defines it.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this. Nothing under newman_b200/ links it; the
 * product path fails loudly without its CUDA library and never falls back to this code.
 *
 * Parity status: PINNED. The phases the reference actually has (plain-double escape loop, series
 * scan, phase-2 search, smoothing) are checked bit-for-bit against Oracle-R = the reference's own
 * mandelbrot.cpp compiled unmodified (oracle/_ref, oracle/Makefile) in tests/test_oracle.py and
 * against the committed fixtures in tests/golden/. The perturbation continuation (phase 3) has no
 * counterpart in the reference (SURVEY.md finding 1): its operation order is defined here and in
 * newman_b200/csrc/k3_perturb.cuh, and its agreement with Oracle-R's arbitrary-precision
 * continuation is measured, not assumed (tests/test_oracle.py::test_continuation_vs_reference).
 */
#ifndef NEWMAN_ORACLE_P_H
#define NEWMAN_ORACLE_P_H
#include <stdint.h>

typedef struct { int32_t iterations; float smoothing; } op_escape; /* grid.h:8-16 */

typedef struct {
  int32_t M, N, has_escape, flags; /* flags & 1: iterate against the truncated orbit (NM_TABLES_ORBIT_TRUNCATED) */
  double tol, glitch_tol;
  const double *x_hi, *x_lo, *a, *b, *c; /* same layout as nm_deep_tables */
  const int32_t *a_exp, *b_exp, *c_exp;  /* floatexp series: a/b/c are mantissas, these the exponents */
  const int32_t *eps_re_exp, *eps_im_exp; /* floatexp eps (=> scaled delta states): eps_* are mantissas */
} op_tables;

typedef struct {
  uint64_t executed_iters, series_evals, skipped_pixels, glitched, rebased;
} op_stats;

/* mandelbrot.cpp:231-254 for one raster; in_cardioid may be NULL (then the test of 63-71 is
 * evaluated in long double, 64-bit mantissa like the reference's 64-bit mpf) */
void oraclep_render_hw(const double* c_re, int nc, const double* c_im, int nr, int N,
                       const uint8_t* in_cardioid, op_escape* out, op_stats* st);

/* mandelbrot.cpp:144-207 verbatim in double + the FP64 perturbation continuation.
 * cardioid_mode/mask, pix_list, mode: as nm_frame_deep. rq_pix/rq_iter receive glitched pixels
 * (capacity = number of work items); returns their count. */
int64_t oraclep_render_deep(const op_tables* t, const double* eps_re, int nc, const double* eps_im, int nr,
                            int cardioid_mode, const uint8_t* mask, const int32_t* pix_list, int64_t n_list,
                            int mode, op_escape* out, int32_t* rq_pix, int32_t* rq_iter, op_stats* st);

/* "Exact mode": phase 3 of the listed samples repeated in double-double arithmetic against the orbit as hi + lo, with
 * rebasing at the end of the orbit; eps_*_lo (may be NULL) = low parts of the pixel offsets. Results overwrite out[pix].
 * Mirrors newman_b200/csrc/k3_dd.cuh. Returns 0, or -1 for scaled frames (not refined). */
int64_t oraclep_refine_dd(const op_tables* t, const double* eps_re, const double* eps_re_lo, int nc, const double* eps_im,
                          const double* eps_im_lo, int nr, const int32_t* pix_list, int64_t n_list, op_escape* out,
                          op_stats* st);

/* Per-pixel probe of the series phase: returns L (d.size() after the scan, mandelbrot.cpp:165-181)
 * and d[L-1]. */
int oraclep_series_L(const op_tables* t, double er, double ei, double* d_re, double* d_im);

/* Rule for choosing the next secondary reference among glitched pixels: earliest flagged iteration,
 * then lowest pixel id. Returns an index into the arrays. */
int64_t oraclep_pick_reference(const int32_t* rq_pix, const int32_t* rq_iter, int64_t n);

/* viewer.cpp:84-124 with this repo's colour conventions (k4_resolve.cuh). rgb: interleaved. */
void oraclep_resolve(const op_escape* grid, int nr, int nc, const uint8_t* pal, int n_pal, int N, int sc,
                     int smooth, uint8_t* rgb);

/* mandelbrot.cpp:320-332: iteration-space box average with a float32 accumulator. */
op_escape oraclep_at_sc(const op_escape* grid, int nc, int r, int c, int sc);

/* multiwave.cpp:75-116 with this repo's colour conventions (CSS hsl2rgb, round(255v); interp truncates).
 * hue cycles are passed flattened: cycle k has hue_counts[k] values followed in hue_values. */
void oraclep_palette_cache(int n_cycles, const int* hue_counts, const float* hue_values, const int* hue_periods,
                           int hue_period, int n_sat, const float* sat_values, int sat_period, int n_lum,
                           const float* lum_amp, const int* lum_period, int N, uint8_t* rgb);

/* video.cpp:14-34 with this repo's resampling conventions (k5_video.cuh). v = (float)pow(1.5, 1.0 / rate)
 * (video.cpp:17) is passed in so that the caller's libm decides its last bit, as in the product (host side). */
void oraclep_video_inbetween(const uint8_t* prev, const uint8_t* next, int H, int W, int nr, int nc, int rate,
                             float v, uint8_t* out);

float oraclep_smoothing(double r2);
double oraclep_trunc_add3(double hi, double lo, double d);
#endif
