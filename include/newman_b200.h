/* newman_b200.h — C-ABI of the B200-native Mandelbrot hot path (libnewman_b200.so).
 *
 * The reference (axnjaxn/newman) has no FFI: its de-facto boundary is the C++ surface
 * `class Mandelbrot` (reference mandelbrot.h:22-53) + `RenderGrid` (grid.h:6-25). This header is the
 * thin C shim underneath our drop-in for that class (include/newman_b200/mandelbrot.h): plain
 * pointers and sizes, int status codes, no exceptions, no torch/C++ types. Every entry point names
 * the reference code it replaces.
 *
 * Conventions
 *   - All `nm_*` functions return NM_OK (0) or a negative NM_E* code; nm_last_error() gives text.
 *   - One nm_ctx per host thread and per GPU. Calls on one ctx are stream-ordered on its stream.
 *   - "h/d pointer": source/destination buffers may be HOST or DEVICE memory (cudaMemcpyDefault
 *     decides by pointer attributes), so a multi-GPU caller can hand over NCCL-received buffers.
 *   - Output records are the reference's `RenderGrid::EscapeValue` (grid.h:8-16): 8 bytes,
 *     {int32 iterations; float32 smoothing}, row-major.
 *   - There is NO CPU fallback: without a CUDA device every compute entry returns NM_ENODEV.
 */
#ifndef NEWMAN_B200_H
#define NEWMAN_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define NM_API __attribute__((visibility("default")))
#else
#define NM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define NM_OK 0
#define NM_EINVAL (-1)  /* bad argument */
#define NM_ENODEV (-2)  /* no CUDA device / driver */
#define NM_ECUDA (-3)   /* CUDA runtime error (see nm_last_error) */
#define NM_ENOMEM (-4)
#define NM_ESTATE (-5)  /* call order violated (e.g. launch without a frame) */
#define NM_ERANGE (-6)  /* tables not representable (non-finite coefficient; reference would SIGFPE) */
#define NM_ECANCELLED (-7)

typedef struct nm_ctx nm_ctx;

/* grid.h:8-16 */
typedef struct nm_escape {
  int32_t iterations;
  float smoothing;
} nm_escape;

/* ---- context ------------------------------------------------------------------------------ */
NM_API int nm_create(int device, nm_ctx** out);
NM_API void nm_destroy(nm_ctx* ctx);
NM_API const char* nm_last_error(const nm_ctx* ctx); /* ctx may be NULL: last error of nm_create */
NM_API const char* nm_version(void);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the ctx stream. */
NM_API int nm_set_stream(nm_ctx* ctx, void* cuda_stream);
/* the stream the ctx currently runs on (its own, or the one nm_set_stream installed), as a cudaStream_t */
NM_API int nm_get_stream(nm_ctx* ctx, void** cuda_stream);
NM_API int nm_sync(nm_ctx* ctx);
/* Tuning / verification switches (defaults in parentheses):
 *   NM_OPT_K2_LITERAL (0)  1: K2 scans every series index like the reference (mandelbrot.cpp:165-181)
 *                          instead of testing only the indices the per-index filter cannot rule out.
 *                          Both give the same L by construction; the tests run both. */
#define NM_OPT_K2_LITERAL 1
/*   NM_OPT_K3_GROUP (4)    pixels per lane in the fast perturbation kernel (k3_fast.cuh): 4 or 2;
 *                          0/1 selects the simple one-pixel-per-lane kernel (k3_perturb.cuh). Same results. */
#define NM_OPT_K3_GROUP 2
/*   NM_OPT_K3_FINISH_MAX (65536; environment NM_K3_FINISH_MAX overrides the default at nm_create)
 *                          a frame — or what is left of one after a sweep — with at most this many states is run
 *                          to completion by ONE launch (k3_finish.cuh: one thread per state, no level launches);
 *                          0: always the level kernels. Same results; the tests run both. */
#define NM_OPT_K3_FINISH_MAX 3
/*   NM_OPT_K3_SPLIT (1)    k3_fast: an orbit chunk that follows a chunk in which more than 1/32 of the states
 *                          escaped is run as 4 quarter-chunk launches with a global compaction after each
 *                          (decided on the device from the queue counters) if it holds at least 303 104 states;
 *                          0: always whole chunks; n > 1: that minimum instead (tests). Same results. */
#define NM_OPT_K3_SPLIT 4
NM_API int nm_set_option(nm_ctx* ctx, int key, int value);
/* Abandon the frame in flight (viewer.cpp:177, 221-231 abort mid-frame). Persistent CTAs poll it. */
NM_API int nm_cancel(nm_ctx* ctx);

/* ---- K1: plain-double escape ----------------------------------------------------------------
 * Replaces Mandelbrot::getIterationsHW + inCardioid + getSmoothingMagnitude
 * (mandelbrot.cpp:231-254, 63-71, 133-136) for every pixel of rows x cols.
 * c_re[nc], c_im[nr]: pixel coordinates, already descended (truncated) by the host exactly as
 * computeRow does (mandelbrot.cpp:271, 275, 234). Separable because the pixel map is.
 * The cardioid/bulb test runs in double on the device; pixels whose margin is too small to decide
 * in double are iterated anyway and reported in the "ambiguous" list so the host can repeat the
 * reference's mpf test for just those (nm_frame_ambiguous). */
NM_API int nm_frame_hw(nm_ctx* ctx, const double* c_re, int nc, const double* c_im, int nr, int N);

/* ---- K2+K3: series skip + perturbation ------------------------------------------------------
 * Replaces Mandelbrot::getIterations (mandelbrot.cpp:144-229): phases 1-2 (series scan, binary
 * search) verbatim in double, phase 3 as the FP64 perturbation iteration against the orbit.
 * Tables are produced by the host with the reference's own arbitrary-precision recurrences
 * (findProbe/computeOrbit/computeSeries, mandelbrot.cpp:73-131) and descended once. */
typedef struct nm_deep_tables {
  int32_t M;            /* orbit length = X.size() (mandelbrot.cpp:106) */
  int32_t N;            /* iteration limit */
  int32_t has_escape;   /* 1: x_hi holds M+1 entries, entry M = the escaped iterate the reference drops */
  int32_t flags;        /* NM_TABLES_* */
  double tol;           /* error_tolerance (isUnstable, mandelbrot.cpp:138-142) */
  double glitch_tol;    /* glitch rule |X_n+d_n|^2 < glitch_tol*|X_n|^2 (same squared-magnitude form) */
  const double* x_hi;   /* [2*(M+has_escape)] re,im interleaved: truncated doubles of X[i] */
  const double* x_lo;   /* [2*M] truncated doubles of X[i]-x_hi[i] (phase-2 'X[mid]+d[mid]' in mpf) */
  const double* a;      /* [2*M] descended A[i] (mandelbrot.cpp:166) */
  const double* b;      /* [2*M] descended B[i] */
  const double* c;      /* [2*M] descended C[i] */
  /* floatexp mode (all three non-NULL): a/b/c then hold MANTISSAS with 0.5 <= |m| < 1 and these the
   * binary exponents, value = m * 2^e per component — exactly what mpf_get_d_2exp returns (truncating,
   * like descend). Needed from pixel pitch ~1e-97 on, where |C| leaves double range and the reference
   * itself stops working (SIGFPE); on shallower views both modes give bit-identical rasters. */
  const int32_t* a_exp; /* [2*M] */
  const int32_t* b_exp;
  const int32_t* c_exp;
  /* floatexp eps (both non-NULL; needs the floatexp series tables): eps_re/eps_im handed to
   * nm_frame_deep are then MANTISSAS (0.5 <= |m| < 1) and these their binary exponents. Selects the
   * scaled perturbation states (delta = (dr, di) * 2^e, csrc/floatexp.cuh) that views below a pixel
   * pitch of ~1e-150 need (delta*delta, then delta and eps themselves, leave double range). */
  const int32_t* eps_re_exp; /* [nc] */
  const int32_t* eps_im_exp; /* [nr] */
  /* NM_MODE_DD only (may be NULL = zero): the low parts of the pixel offsets, trunc((pixel - X[0]) - eps) in mpf */
  const double* eps_re_lo;   /* [nc] */
  const double* eps_im_lo;   /* [nr] */
} nm_deep_tables;
#define NM_TABLES_ORBIT_TRUNCATED 1 /* K3 iterates against the orbit TRUNCATED to doubles (x_hi as it is) instead of rounded to
                                       nearest (x_hi + x_lo): the second rendering of the exact mode's sensitivity probe */

#define NM_CARDIOID_NONE 0 /* no pixel of the view is inside cardioid/bulb (mandelbrot.cpp:149) */
#define NM_CARDIOID_ALL 1  /* every pixel is */
#define NM_CARDIOID_MASK 2 /* per-pixel byte mask supplied */

#define NM_MODE_REQUEUE 0 /* glitched pixels are flagged and listed for a secondary reference */
#define NM_MODE_REBASE 1  /* final pass: no glitch flagging; rebase onto orbit start when |z|<|d| */
#define NM_MODE_DD 2      /* listed samples only: phase 3 in double-double arithmetic (csrc/k3_dd.cuh), the refinement pass
                             of the exact mode; not for scaled frames (floatexp eps) */

/* Starts a deep frame over the whole raster (pix_list == NULL) or over the listed pixel ids
 * (secondary-reference rounds, probe search; ids are r*nc+c; unlisted samples keep their values when
 * the previous frame was a deep frame of the same size and are cleared otherwise). eps_re[nc], eps_im[nr] are the truncated doubles of
 * (pixel - X[0]) formed in mpf as mandelbrot.cpp:155-159. */
NM_API int nm_frame_deep(nm_ctx* ctx, const nm_deep_tables* t, const double* eps_re, int nc,
                  const double* eps_im, int nr, int cardioid_mode, const uint8_t* cardioid_mask,
                  const int32_t* pix_list, int64_t n_list, int mode);

/* Enqueue all kernels of the current frame on the ctx stream (asynchronous). */
NM_API int nm_launch(nm_ctx* ctx);

/* After nm_launch: sizes/contents of the host-assist lists (these calls synchronise).
 *  ambiguous: pixel ids whose double cardioid test was undecidable (K1 only).
 *  requeue:   pixel ids flagged as glitched (K3, NM_MODE_REQUEUE), with the iteration of the flag.
 * Buffers may be NULL to query the count. */
NM_API int64_t nm_frame_ambiguous(nm_ctx* ctx, int32_t* pix, int64_t cap);
NM_API int64_t nm_frame_requeue(nm_ctx* ctx, int32_t* pix, int32_t* at_iter, int64_t cap);
/* Overwrite one pixel's record (host verdict for an ambiguous pixel). */
NM_API int nm_poke(nm_ctx* ctx, int64_t pix, nm_escape v);

/* Copy rows [r0, r1) of the finished raster to dst (h/d pointer, (r1-r0)*nc records).
 * Applies the float32-rounding fix-ups first (see DESIGN.md "smoothing"): the handful of pixels
 * whose smoothing value sits within the device libm's error of a float32 rounding boundary are
 * re-evaluated with the host libm the reference uses (mandelbrot.cpp:133-136). */
NM_API int nm_read_rows(nm_ctx* ctx, int r0, int r1, nm_escape* dst);

/* Exact mode (Mandelbrot::exact). The frame renders twice — against the orbit rounded to nearest and
 * against the truncated one (NM_TABLES_ORBIT_TRUNCATED) — and the samples whose escape COUNT depends on that are the ones FP64
 * perturbation cannot resolve: they are repeated with NM_MODE_DD.
 *   nm_raster_keep     remember the current raster on the device
 *   nm_raster_diff     the ids of the samples whose count differs between the current raster and the remembered one; returns
 *                      their number (pix may be NULL to query; at most cap ids are copied) or a negative code
 *   nm_raster_restore  the remembered raster becomes the current one again */
NM_API int nm_raster_keep(nm_ctx* ctx);
NM_API int64_t nm_raster_diff(nm_ctx* ctx, int32_t* pix, int64_t cap);
NM_API int nm_raster_restore(nm_ctx* ctx);

/* Same with a destination row pitch: row r lands at dst + (r - r0) * dst_pitch_bytes. A rank that renders
 * the interleaved rows rank, rank+N, ... of a frame writes its band straight into the shared (pinned)
 * host raster with pitch N * nc * 8 — every GPU over its own PCIe link, no funnel through one GPU. */
NM_API int nm_read_rows_pitched(nm_ctx* ctx, int r0, int r1, nm_escape* dst, size_t dst_pitch_bytes);

/* The same copy without waiting for it: the rows are snapshot on the device (a device-to-device copy on the ctx stream)
 * and leave for dst on a second stream, so the ctx can start the next frame — H2D of its tables, K2, K3 — while this
 * frame's raster is still crossing PCIe. The transfer itself is handed to the copy engine only when the NEXT frame's
 * kernels are enqueued (nm_launch) or by nm_read_wait: copy engines serve the streams of a context in order, and a raster
 * started at the frame boundary holds the next frame's uploads back. dst should be page-locked and must stay valid until
 * nm_read_wait, which returns when every copy started this way has landed. */
NM_API int nm_read_rows_pitched_async(nm_ctx* ctx, int r0, int r1, nm_escape* dst, size_t dst_pitch_bytes);
NM_API int nm_read_wait(nm_ctx* ctx);

/* Page-lock / release a caller-owned host buffer (e.g. a raster in POSIX shared memory that several ranks
 * write with nm_read_rows_pitched) so that copies to it are DMA transfers. NM_ENOMEM when the OS refuses;
 * the buffer then still works as pageable memory. */
NM_API int nm_host_register(nm_ctx* ctx, void* ptr, size_t bytes);
NM_API int nm_host_unregister(nm_ctx* ctx, void* ptr);

/* Gather the records of the listed samples (ids r*nc+c; h/d pointers) — what the probe search reads. */
NM_API int nm_read_pixels(nm_ctx* ctx, const int32_t* pix, int64_t n, nm_escape* dst);

typedef struct nm_stats {
  uint64_t pixels;          /* samples in the frame (or list) */
  uint64_t executed_iters;  /* z<-z^2+c (K1) or delta updates (K3) actually performed */
  uint64_t series_evals;    /* exact isUnstable evaluations performed by K2 */
  uint64_t skipped_pixels;  /* cardioid/bulb pixels */
  uint64_t glitched;        /* pixels flagged for re-queue */
  uint64_t rebased;         /* rebase events (orbit exhausted, or |z|<|d| in NM_MODE_REBASE) */
  uint64_t fixups;          /* smoothing values re-evaluated on the host */
  uint64_t kernel_launches; /* kernels of ours launched for this frame */
  uint64_t sweeps;          /* K3 passes over the orbit */
  uint64_t checked_steps;   /* k3_fast: lane-steps taken one at a time with exact checks (rest run in blocks of 4) */
  float ms_k1, ms_k2, ms_k3, ms_k4; /* device time per stage (CUDA events on the ctx stream) */
} nm_stats;
NM_API int nm_frame_stats(nm_ctx* ctx, nm_stats* out);

/* One-call forms with HOST buffers = what Mandelbrot::precompute()+computeRow() drive:
 * upload, launch, host-assist, download. These are the "e2e" calls bench.py times. */
NM_API int nm_render_hw(nm_ctx* ctx, const double* c_re, int nc, const double* c_im, int nr, int N,
                 nm_escape* out);
NM_API int nm_render_deep(nm_ctx* ctx, const nm_deep_tables* t, const double* eps_re, int nc,
                   const double* eps_im, int nr, int cardioid_mode, const uint8_t* cardioid_mask,
                   const int32_t* pix_list, int64_t n_list, int mode, nm_escape* out);

/* ---- K4: colour resolve ---------------------------------------------------------------------
 * Replaces FractalViewer::getColor/colorLine/recolor (viewer.cpp:84-124): palette lookup, smooth
 * interpolation, sc x sc RGB-space box average, clip. pal_rgb: 3*n_pal bytes from
 * MultiWaveGenerator::cache (multiwave.cpp:75-116). Reads the raster of the current frame; writes
 * (nr/sc) x (nc/sc) x 3 interleaved bytes to rgb_out (h/d pointer). */
NM_API int nm_resolve(nm_ctx* ctx, const uint8_t* pal_rgb, int n_pal, int N, int sc, int smooth,
               uint8_t* rgb_out);
/* Same, on a caller-supplied raster (h/d pointer) instead of the current frame. */
NM_API int nm_resolve_grid(nm_ctx* ctx, const nm_escape* grid, int nr, int nc, const uint8_t* pal_rgb,
                    int n_pal, int N, int sc, int smooth, uint8_t* rgb_out);

/* ---- K5: zoom-video in-betweening -------------------------------------------------------------
 * Replaces the frame loop of VideoZoom::nextFrame (video.cpp:14-34): from the previous and the new key
 * frame (H x W x 3 bytes each, interleaved; the new one rendered 1.5x deeper) all `rate` in-between
 * canvases (nr x nc x 3 each, consecutive in frames_out) in one launch: previous enlarged by 1.5^(i/rate),
 * new one shrunk by (2/3) 1.5^(i/rate) blended over it with weight i/rate, both centred. h/d pointers.
 * Resampling conventions (libbyteimage is not vendored): see csrc/k5_video.cuh. ms_k4 of nm_frame_stats
 * holds the kernel time. Encoding the frames (byteimage::VideoWriter) stays with the caller. */
NM_API int nm_video_inbetween(nm_ctx* ctx, const uint8_t* prev_rgb, const uint8_t* next_rgb, int H, int W, int nr,
                              int nc, int rate, uint8_t* frames_out);

/* ---- K6: palette table on the device ----------------------------------------------------------
 * MultiWaveGenerator::cache (multiwave.cpp:75-116) as a kernel: the N-entry RGB table is generated in
 * HBM from the generator's parameters (hue cycles flattened: cycle k has hue_counts[k] nodes, in degrees,
 * consecutive in hue_values) and kept in the ctx; rgb_out (h/d pointer, may be NULL) receives a copy.
 * nm_resolve_device_palette then recolours the resident raster with it: a palette edit costs two small
 * launches and no table upload (README.md:84-86 "re-index without re-rendering"). Within 1 LSB per channel
 * of the host builder nmp_cache (device libm). */
NM_API int nm_palette_cache(nm_ctx* ctx, int n_cycles, const int* hue_counts, const float* hue_values,
                            const int* hue_periods, int hue_period, int n_sat, const float* sat_values,
                            int sat_period, int n_lum, const float* lum_amp, const int* lum_period, int N,
                            uint8_t* rgb_out);
NM_API int nm_resolve_device_palette(nm_ctx* ctx, int N, int sc, int smooth, uint8_t* rgb_out);

/* ---- measurement helpers --------------------------------------------------------------------
 * FP64-pipe peak probe: runs `iters` dependent-chain DFMA (kind 0), DADD (1), DMUL (2) per thread
 * (kind 3: the former K3 iteration body from registers, counted as 10 instructions per pixel-iteration;
 * kinds 4-7: 8 DFMA chains interleaved with 0/8/16/24 integer-pipe operations, 64 warps per SM; kinds 8-11:
 * the same at 16 warps per SM — only the FP64 instructions are counted; kinds 12-15: DFMA whose instructions
 * read 3 / 2 / 1 distinct 64-bit registers that no neighbour shares, and DADD with 2: the register-operand
 * bandwidth behind the FP64 pipe)
 * over a full-chip grid and returns instructions/s. Used by bench.py for the roofline denominator
 * (MEASURED_PEAKS.json carries no FP64 entry). */
NM_API int nm_fp64_peak(nm_ctx* ctx, int kind, int iters, double* inst_per_s, double* ms);
/* Host-side evaluation of the K3 candidate filter (newman_b200/csrc/k3_filter.cuh) — no GPU involved: the
 * table entry k3_fast would use at an orbit index with Z = (zr, zi) and glitch bound gb
 * (entry[0..3] = lo_r, w_r, lo_i, w_i; entry[4] = escape high word), and the two tests as the kernel makes
 * them (scaled != 0: the state is a scaled one). The filters replace the per-iteration |z|^2 tests of the
 * perturbation loop and must have no false negatives; tests/test_k3_filter.py attacks that claim. */
NM_API void nm_k3_filter_entry(double zr, double zi, double gb, uint32_t entry[5]);
NM_API int nm_k3_filter_fires(const uint32_t entry[5], double dr, double di, int scaled, int* glitch, int* escape);
/* Host-side evaluation of k3_fast's quiet bound for the segment of 16 iterations that starts at orbit index j0
 * (k3_filter.cuh: k3_seg_bound): z = the table Z[0..jmax] as interleaved (re, im) doubles, gb[0..jmax] the glitch
 * bounds, e_max >= |eps| of every sample. Returns the high word T: a state with hi(|dr|) < T and hi(|di|) < T at j0
 * cannot satisfy the glitch test at any of the indices j0+1 .. j0+16 (0 = no state is exempt). */
NM_API int32_t nm_k3_seg_bound(const double* z, const double* gb, int j0, int jmax, double e_max);
NM_API int nm_device_info(nm_ctx* ctx, int* sm_count, int* sm_clock_khz, size_t* hbm_bytes, char* name, int cap);

/* ==== view level: the drop-in class through C ===================================================
 * `nmv_view` wraps one `class Mandelbrot` (include/newman_b200/mandelbrot.h == reference
 * mandelbrot.h:22-53), so bindings without a C++ toolchain (ctypes, cgo, JNI ...) drive the same
 * object the viewer would. Strings are base-10 mpf literals. No C++ exception crosses this ABI:
 * failures return a negative code and nmv_last_error() has the text. */
typedef struct nmv_view nmv_view;

typedef struct nmv_frame_info {
  int32_t hardware /* 1 plain double, 0 series+perturbation, 2 same with floatexp series, 3 floatexp series + scaled deltas */, precision_bits, orbit_len, probe_row, probe_col, references;
  uint64_t executed_iters, series_evals, skipped_pixels, glitched, rebased, fixups, kernel_launches, ambiguous;
  double host_precompute_s, device_ms, frame_s;
  uint64_t probe_iters, probe_exact; /* GPU-assisted findProbe: delta updates on candidates; candidates measured in mpf */
  int32_t probe_consistent; /* 0: the exact check contradicts the ranking (ill-conditioned view): see mandelbrot.h */
  int32_t cancelled;        /* 1: the frame was abandoned by nmv_cancel: the raster is not current */
  uint64_t refined;         /* exact mode: samples repeated in double-double arithmetic */
  double refine_ms;         /* exact mode: device time of the probe rendering + the double-double pass */
} nmv_frame_info;

NM_API nmv_view* nmv_create(int nr, int nc);                      /* Mandelbrot(nr, nc), mandelbrot.cpp:8-17 */
NM_API void nmv_destroy(nmv_view* v);
NM_API const char* nmv_last_error(const nmv_view* v);
/* SURVEY.md 8d construction order: N; sz from strings; zoom(1.0f) (-> setPrecision); centre; tolerance.
 * sz_* / c_* may be NULL to keep the current value. */
NM_API int nmv_set_view(nmv_view* v, int N, const char* sz_re, const char* sz_im, const char* c_re,
                        const char* c_im, double tol);
NM_API int nmv_set_options(nmv_view* v, double glitch_tol, int max_secondary, int device, int host_threads);
/* force: 1 evaluate the series in floatexp, 2 also floatexp eps + scaled delta states, even where doubles
 * suffice (both are automatic once the view needs them); 0 automatic */
NM_API int nmv_set_floatexp(nmv_view* v, int force);
/* Mandelbrot::exact: 1 = after the frame, find the samples FP64 perturbation cannot resolve (the frame is rendered a second
 * time against the truncated orbit; counts that differ) and repeat them in double-double arithmetic. ~2.3x the frame time;
 * on cfg2 every sampled escape count then equals the reference's. Not applied to scaled frames (pitch < 2^-380). */
NM_API int nmv_set_exact(nmv_view* v, int on);
/* Mandelbrot::devices / band_rows: n > 1 splits every frame over these GPUs (one host thread, one context and one NCCL
 * rank per GPU inside the view; see "multi-GPU" below); n = 0 returns to the single `device`. band_rows <= 0 keeps 4. */
NM_API int nmv_set_devices(nmv_view* v, const int* devices, int n, int band_rows);
/* findProbe (mandelbrot.cpp:73-95). mode 0: the reference's exhaustive arbitrary-precision search; 1:
 * GPU-assisted (candidates' orbit lengths by K2/K3 against one reference, exact mpf check of the
 * short-list). nmv_set_probe_search picks what precompute() uses (default 1); nmv_find_probe runs one
 * search and reports the winner, its exact orbit length and how many candidates were measured in mpf. */
NM_API int nmv_set_probe_search(nmv_view* v, int mode);
NM_API int nmv_find_probe(nmv_view* v, int mode, int* row, int* col, int* length, int* n_exact);
NM_API int nmv_get_n(const nmv_view* v);                            /* the public field N (loadLegacy sets it) */
NM_API int nmv_rows(const nmv_view* v);
NM_API int nmv_cols(const nmv_view* v);
NM_API int nmv_use_hardware(nmv_view* v);                          /* mandelbrot.cpp:256-259 */
NM_API int nmv_precision_bits(const nmv_view* v);                  /* what setPrecision chose, 37-51 */
NM_API int nmv_precompute(nmv_view* v);                            /* mandelbrot.cpp:261-267 (+ GPU frame) */
NM_API int nmv_compute_row(nmv_view* v, int r);                    /* mandelbrot.cpp:269-283 */
/* Abandon the frame nmv_precompute / nmv_render is rendering in ANOTHER thread (viewer.cpp:177, 221-231: the viewer stops
 * calling computeRow): that call returns early with nmv_frame_info.cancelled = 1. Thread-safe, never blocks. */
NM_API int nmv_cancel(nmv_view* v);
NM_API int nmv_render(nmv_view* v, nm_escape* out);                /* precompute + every row, raster copied out */
NM_API int nmv_read_grid(const nmv_view* v, nm_escape* out);       /* at(r,c) for all r,c: mandelbrot.cpp:318 */
NM_API int nmv_write_grid(nmv_view* v, const nm_escape* in);
NM_API int nmv_at_sc(nmv_view* v, int r, int c, int sc, nm_escape* out); /* mandelbrot.cpp:320-332 */
NM_API int nmv_scale(nmv_view* v, int sc, int up);                 /* scaleUp / scaleDown, 334-360 */
NM_API int nmv_zoom(nmv_view* v, float scale);                     /* 297-302 */
NM_API int nmv_translate(nmv_view* v, int dr, int dc, int sc);     /* 292-295 */
NM_API int nmv_zoom_at(nmv_view* v, float scale, int r, int c, int sc); /* 304-316 */
NM_API int nmv_load_legacy(nmv_view* v, const char* fn);           /* 19-35 */
NM_API int nmv_save(nmv_view* v, const char* fn);                  /* viewer.cpp:12-23 format */
/* which: 0 centre.re, 1 centre.im, 2 sz.re, 3 sz.im; "<mantissa digits>@<exp>" as mpf_get_str */
NM_API int nmv_view_string(const nmv_view* v, int which, char* buf, int cap);
NM_API int nmv_frame_info_get(const nmv_view* v, nmv_frame_info* out);
NM_API int nmv_resolve(nmv_view* v, const uint8_t* pal_rgb, int n_pal, int sc, int smooth, uint8_t* rgb_out);

/* Host-only pieces of precompute() (no GPU needed): used by the CPU test-suite to pin the
 * arbitrary-precision tables against the compiled reference.
 * nmv_host_tables: row/col < 0 runs findProbe (mandelbrot.cpp:73-95), else uses that pixel as the
 * reference point. Returns M (orbit length) or a negative code; the arrays are then read with
 * nmv_host_table (which: 0 x_hi[2*(M+has_escape)], 1 x_lo[2M], 2 a, 3 b, 4 c, 5 eps_re[nc], 6 eps_im[nr]). */
NM_API int nmv_host_tables(nmv_view* v, int row, int col, int* has_escape, int* probe_row, int* probe_col);
NM_API int nmv_host_table(const nmv_view* v, int which, double* out);
/* floatexp form of the coefficients: nmv_host_table which 7/8/9 = mantissas of a/b/c (0.5 <= |m| < 1),
 * nmv_host_table_exp which 2/3/4 = their binary exponents [2M]; likewise 10/11 = eps_re[nc] / eps_im[nr]
 * mantissas and exponents */
NM_API int nmv_host_table_exp(const nmv_view* v, int which, int32_t* out);
/* 1 if the hand-laid-out arbitrary-precision values of the pipelined table build (csrc/hp_host.cpp: HpStream) behave
 * exactly like mpf_init2 values under the running libgmp (checked once per process); 0: the build falls back to
 * mpf_init2 per value. */
NM_API int nmv_host_selfcheck(void);
NM_API int nmv_host_coords(nmv_view* v, double* c_re, double* c_im);       /* mandelbrot.cpp:271, 275, 234 */
NM_API int nmv_host_cardioid(nmv_view* v, uint8_t* mask_or_null);          /* returns NM_CARDIOID_* */
NM_API int nmv_host_in_cardioid(nmv_view* v, int r, int c);                /* mandelbrot.cpp:63-71 */

/* ==== multi-GPU: one frame over several GPUs of a node ==========================================
 * The path shards by bands of grid rows with no collective on the per-pixel data path (reference viewer.cpp:193-238:
 * the beauty render computes rows one after another; rows are independent once precompute() has run). A render group
 * has one nmm_rank per GPU: threads of one process (what `Mandelbrot::devices` sets up by itself) or one process per
 * GPU (torchrun, MPI ...: rank 0 calls nmm_unique_id and the caller carries the 128 bytes to the others). Rank r renders
 * the blocks of `band_rows` grid rows b = r, r + world, ... (band_rows a multiple of the multisampling factor, so that a
 * colour-resolve block never straddles GPUs). Exchanged per frame: each reference's tables (host GMP on rank 0) by
 * ncclBroadcast; the 8-byte "which glitched sample is the next reference" MIN by ncclAllReduce; the finished bands.
 * NCCL is bound at run time (dlopen libnccl.so.2). world == 1 needs no NCCL. */
typedef struct nmm_rank nmm_rank;
#define NMM_ID_BYTES 128
#define NMM_RETURN_LOCAL 0 /* every rank copies its own bands into the host image IT was given (threads of one process,
                              or a shared page-locked mapping): every GPU over its own PCIe link */
#define NMM_RETURN_ROOT 1  /* bands travel to rank 0 by ncclSend/ncclRecv; rank 0 writes the host image */
/* Host-only (no GPU, no NCCL): the global grid rows rank `rank` of `world` owns when an nr-row raster is dealt in blocks of
 * band_rows rows — what nmm_render uses. rows_out (may be NULL) receives them in the rank's local order; returns their number
 * or NM_EINVAL (band_rows must divide nr). */
NM_API int nmm_band_layout(int nr, int band_rows, int rank, int world, int32_t* rows_out);
NM_API int nmm_unique_id(uint8_t id[NMM_ID_BYTES]);
NM_API int nmm_create(int device, int rank, int world, const uint8_t id[NMM_ID_BYTES], nmm_rank** out);
NM_API void nmm_destroy(nmm_rank* rk);
NM_API const char* nmm_last_error(const nmm_rank* rk);   /* rk may be NULL: last error of nmm_create / nmm_unique_id */
NM_API nm_ctx* nmm_ctx(nmm_rank* rk);
/* Collective: every rank calls it with a view in the SAME state (same size, N, centre, sz, tolerances, options).
 * precompute() (rank 0: probe search, orbit, series; mandelbrot.cpp:261-267) + every row of this rank's bands.
 * out_raster: nr*nc records of the whole frame — see NMM_RETURN_*; NULL (on every rank alike) leaves the bands on the
 * GPUs for nmm_resolve. info (may be NULL): counters summed over the ranks, device_ms the slowest rank's. */
NM_API int nmm_render(nmm_rank* rk, nmv_view* view, int band_rows, nm_escape* out_raster, int return_mode, nmv_frame_info* info);
/* Collective colour resolve of the bands of the last nmm_render (K4 per band: getColor/colorLine, viewer.cpp:84-124),
 * returned as (nr/sc) x (nc/sc) x 3 bytes like nm_resolve. */
NM_API int nmm_resolve(nmm_rank* rk, const uint8_t* pal_rgb, int n_pal, int sc, int smooth, uint8_t* out_rgb, int return_mode);
/* host milliseconds this rank has spent inside exchanges (broadcasts, reductions, band return) since nmm_create */
NM_API double nmm_exchange_ms(const nmm_rank* rk);

/* ==== palette: MultiWaveGenerator through C (include/newman_b200/multiwave.h == reference
 * multiwave.h:8-37). The N-entry RGB table is built on the host, like the reference
 * (multiwave.cpp:75-116, viewer.cpp:66-68); nm_resolve / nmv_resolve consume it. */
typedef struct nmp_palette nmp_palette;
NM_API nmp_palette* nmp_create(void);
NM_API void nmp_destroy(nmp_palette* p);
NM_API int nmp_load_file(nmp_palette* p, const char* fn);            /* .pal text, multiwave.cpp:19-48 */
NM_API int nmp_save_file(const nmp_palette* p, const char* fn);      /* multiwave.cpp:50-73 */
NM_API int nmp_clear(nmp_palette* p);
NM_API int nmp_add_hue_cycle(nmp_palette* p, const float* hues_deg, int n, int period);
NM_API int nmp_set_hue_period(nmp_palette* p, int period);
NM_API int nmp_set_sat_cycle(nmp_palette* p, const float* sats, int n, int period);
NM_API int nmp_add_lum_wave(nmp_palette* p, float amplitude, int period);
NM_API int nmp_cache(const nmp_palette* p, int N, uint8_t* rgb_out); /* cache(N): 3*N bytes r,g,b */
/* the same table generated on the GPU of `ctx` (K6, nm_palette_cache) and kept there for nm_resolve_device_palette */
NM_API int nmp_cache_device(const nmp_palette* p, nm_ctx* ctx, int N, uint8_t* rgb_out);

#ifdef __cplusplus
}
#endif
#endif
