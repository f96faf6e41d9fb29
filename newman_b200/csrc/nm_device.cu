// Device layer of libnewman_b200.so: context, HBM residency of the per-frame tables, kernel
// orchestration and the C-ABI declared in include/newman_b200.h.
//
// HBM layout per frame (all grow-only allocations owned by the ctx):
//   raster   out[nr*nc]            8 B/sample  {int32 iterations, float32 smoothing}  (grid.h:8-16)
//   K1       c_re[nc], c_im[nr]    separable pixel coordinates (doubles)
//   K2/K3    Z[Jmax+1]   16 B      Z[0]=0, Z[j]=X[j-1] (+ the escaped iterate X[M] when known)
//            gb[Jmax+1]   8 B      glitch bound glitch_tol*|Z[j]|^2,  ghi[] = its high word (4 B)
//            Xlo[M]      16 B      low parts of X (phase-2 truncated add)
//            A,B,C[M]    16 B each descended series coefficients
//            eps_re[nc], eps_im[nr]
//            init_d[W] 16 B, init_j[W] 4 B, fresh_ids[~W] 4 B    K2 -> K3 hand-over, sorted by start index
//            q[2][W], rq[2][W]   32 B PixState      level ping-pong queues, rebase queues
//            rq_pix[W], rq_iter[W]                   glitch re-queue list
// No CPU fallback exists: without a device every entry returns NM_ENODEV.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "k1_escape.cuh"
#include "k2_series.cuh"
#include "k3_perturb.cuh"
#include "k3_fast.cuh"
#include "k3_finish.cuh"
#include "k3_dd.cuh"
#include "k4_resolve.cuh"
#include "k5_video.cuh"
#include "k6_palette.cuh"
#include "fp64_probe.h"

using namespace nm;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return (T*)p; }
};

std::string g_create_error;

}  // namespace

struct nm_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own = nullptr, stream = nullptr, side = nullptr;
  std::string err;

  int kind = 0;  // 0 none, 1 hw, 2 deep
  int nr = 0, nc = 0, N = 0;
  long long pixels = 0, W = 0;
  bool have_list = false;
  bool launched = false, finished = false;

  DevBuf out, cre, cim, ctr, ambig, fix, fixapply, palpar, paldev, vprev, vnext, vout;
  DevBuf kept, difflist, cre_lo, cim_lo;   // exact mode: the remembered raster (nm_raster_keep); low parts of the pixel offsets (NM_MODE_DD)
  long long kept_pixels = 0;
  bool have_eps_lo = false;
  int paldev_n = 0;  // entries of the device-generated palette in paldev (0: none)
  unsigned long long ambig_cap = 0, fix_cap = 0;
  // deep
  DevBuf Z, ghi, gb, Z2, k3filt, esc_hi, seg_hi, eps_max, xhi, xlo, a, b, c, mask, list, fa_d[2], fa_i[2], hist, offs, cursor, fresh, q[2], rq[2], qctr,
      rq_pix, rq_iter, pal, rgb, gridtmp, filt, events, snap, aexp, bexp, cexp, cre_e, cim_e;
  int use_fe = 0;   // 0 double series, 1 floatexp series, 2 floatexp series + floatexp eps + scaled K3 states
  int opt_k2_literal = 0;
  int opt_k3_group = 4;  // pixels per lane in k3_fast (0: simple kernel only)
  int opt_k3_split = 1;  // k3_fast: levels that follow an escape-heavy level run as 4 quarter-chunk launches
  long long opt_k3_finish_max = K3_FINISH_MAX_STATES;  // frames / remainders up to this many states: k3_finish
  int occ_k3f[2] = {0, 0}, occ_k3fs[2] = {0, 0};
  int M = 0, Jmax = 0, K = 0, CH = 1024, mode = 0, cardioid_mode = 0, has_escape = 0;
  double tol = 0, gtol = 0;

  nm_stats stats;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_snap = nullptr, ev_copy = nullptr;   // nm_read_rows_pitched_async: snapshot taken / copy-out finished
  bool copy_pending = false;
  // a copy-out whose snapshot is taken but whose device-to-host transfer has not been handed to the copy engine yet
  struct { nm_escape* dst = nullptr; size_t pitch = 0, row_bytes = 0, rows = 0; bool active = false; } deferred;
  unsigned long long* h_ctr = nullptr;  // pinned mirror of the counters
  uint32_t* h_small = nullptr;          // pinned landing zone of kernel-written read-backs (read_back)
  unsigned long long* h_flag = nullptr; // pinned cancel flag source
  double log_bailout = 0;
  int occ_k1 = 0, occ_k3[2] = {0, 0}, occ_k3s[2] = {0, 0};
};

namespace {

int fail(nm_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}

#define NM_CUDA(ctx, call)                                                                     \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? NM_ENOMEM : NM_ECUDA, "%s: %s", #call, \
                  cudaGetErrorString(e__));                                                    \
  } while (0)

// Per-index tables derived from the orbit: the glitch bound (and its high word for k3_level), and for
// k3_fast 2*Z[j] and the candidate-filter entries of k3_filter.cuh.
// The orbit table the perturbation kernels pair delta with: X rounded TO NEAREST, Z[i + 1] = RN(x_hi[i] + x_lo[i]).
// x_hi is the truncated double the reference's descend() yields (complex.h:33-35) and phases 1-2 need; iterating
// against it biases every factor 2Z + delta the same way and the relative error of delta grows linearly with the
// iteration count (5e-13 after cfg2's 7 000 iterations instead of 5e-15), which moves the escape count of samples whose
// last few hundred iterations are chaotic (oracle/oracle_p.c makes the same table).
__global__ void k_round_orbit(double2* Z1, const double2* xhi, const double2* xlo, int M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const double2 h = xhi[i], l = xlo[i];
  Z1[i] = make_double2(h.x + l.x, h.y + l.y);
}

__global__ void k_glitch_bounds(const double2* Z, double* gb, int32_t* ghi, double2* Z2, K3Filt* filt, int32_t* esc_hi,
                                int n, int pad_n, double gtol, int zero_last) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= pad_n) return;
  double v = 0.0;
  const double2 z = Z[j];   // zero beyond the table
  if (j > 0 && j < n && !(zero_last && j == n - 1)) v = (z.x * z.x + z.y * z.y) * gtol;
  gb[j] = v;
  ghi[j] = __double2hiint(v);
  Z2[j] = make_double2(2.0 * z.x, 2.0 * z.y);
  K3Filt f; int32_t e;
  k3_filter_entry(z.x, z.y, v, &f, &e);
  filt[j] = f;
  esc_hi[j] = e;
}

// Largest pixel offset of the frame: |eps| <= sqrt(max|eps_re|^2 + max|eps_im|^2), rounded up (one CTA).
__global__ void k_eps_max(EpsTab t, int nr, double* out) {
  __shared__ double sh[2][32];
  double mr = 0.0, mi = 0.0;
  for (int i = threadIdx.x; i < t.nc; i += blockDim.x) {
    double v = fabs(t.re[i]);
    if (t.re_e) v = scalbn(v, t.re_e[i]);
    mr = v > mr || v != v ? v : mr;
  }
  for (int i = threadIdx.x; i < nr; i += blockDim.x) {
    double v = fabs(t.im[i]);
    if (t.im_e) v = scalbn(v, t.im_e[i]);
    mi = v > mi || v != v ? v : mi;
  }
  for (int o = 16; o; o >>= 1) {
    const double a = __shfl_xor_sync(FULL_MASK, mr, o), b = __shfl_xor_sync(FULL_MASK, mi, o);
    mr = a > mr || a != a ? a : mr;
    mi = b > mi || b != b ? b : mi;
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = mr; sh[1][threadIdx.x >> 5] = mi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const double a = sh[0][w], b = sh[1][w];
      mr = a > mr || a != a ? a : mr;
      mi = b > mi || b != b ? b : mi;
    }
    *out = sqrt(mr * mr + mi * mi) * (1.0 + 9.094947017729282e-13);   // NaN / inf => no segment is quiet
  }
}

// k3_fast's per-segment quiet bounds (k3_filter.cuh: k3_seg_bound), one thread per segment of 16 orbit indices.
__global__ void k_seg_bounds(const double2* Z, const double* gb, int jmax, const double* e_max, int32_t* seg_hi, int n_seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const double* z = (const double*)Z;
  seg_hi[s] = k3_seg_bound(z, z + 1, gb, 2, 16 * s, jmax, *e_max);
}


__global__ void k_apply_fixups(nm_escape* out, const int32_t* pix, const float* val, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[pix[i]].smoothing = val[i];
}

__global__ void k_poke(nm_escape* out, long long pix, nm_escape v) { out[pix] = v; }

__global__ void k_gather(const nm_escape* out, const int32_t* pix, nm_escape* dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = out[pix[i]];
}

int set_device(nm_ctx* ctx) {
  NM_CUDA(ctx, cudaSetDevice(ctx->device));
  return NM_OK;
}

int reset_frame_counters(nm_ctx* ctx) {
  NM_CUDA(ctx, cudaMemsetAsync(ctx->ctr.p, 0, CTR_COUNT * sizeof(unsigned long long), ctx->stream));
  return NM_OK;
}

int size_lists(nm_ctx* ctx) {
  ctx->ambig_cap = (unsigned long long)ctx->W;
  ctx->fix_cap = (unsigned long long)(ctx->W / 16 + 65536);
  NM_CUDA(ctx, ctx->ambig.ensure(ctx->ambig_cap * sizeof(int32_t)));
  NM_CUDA(ctx, ctx->fix.ensure(ctx->fix_cap * sizeof(FixupRec)));
  return NM_OK;
}

// Host re-evaluation of the smoothing values the device could not round safely
// (reference mandelbrot.cpp:133-136, same expression, host libm).
float host_smoothing(double r2) {
  const double bailout = 1024.0;
  return (float)(1.0 - log2(0.5 * log(r2) / log(bailout)));
}

// Small device -> host read-backs on a frame's critical path (queue counters after K2 and after every sweep, the frame's
// counters, the glitch / fix-up / ambiguous lists) are written into page-locked host memory BY A KERNEL, not by a copy
// engine: the device-to-host engine is shared by all streams of the context and works in order, so an 8-byte read-back
// queued behind the previous frame's raster (nm_read_rows_pitched_async: 66 MB, 3-6 ms) waited for all of it — measured on
// 8 x B200: every frame of the overlapped end-to-end loop took 5 ms longer than its kernels (profiles/r02j_*).
constexpr size_t NM_SMALL_D2H = 256 << 10;

__global__ void k_to_host(const uint32_t* __restrict__ src, volatile uint32_t* dst, unsigned n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
__global__ void k_three_to_host(const unsigned long long* a, const unsigned long long* b, const unsigned long long* c,
                                volatile unsigned long long* dst) {
  if (threadIdx.x == 0) { dst[0] = *a; dst[1] = *b; dst[2] = *c; }
  __threadfence_system();
}

// dst <- bytes at device address src, stream-ordered, returns after they have arrived
int read_back(nm_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return NM_OK;
  if (bytes <= NM_SMALL_D2H && (bytes & 3) == 0 && ((uintptr_t)src & 3) == 0) {
    const unsigned n = (unsigned)(bytes / 4);
    unsigned blocks = (n + 255) / 256;
    if (blocks > 64) blocks = 64;
    k_to_host<<<blocks, 256, 0, ctx->stream>>>((const uint32_t*)src, ctx->h_small, n);
    NM_CUDA(ctx, cudaGetLastError());
    NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(dst, ctx->h_small, bytes);
    return NM_OK;
  }
  NM_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NM_OK;
}

// The device-to-host transfer of a deferred copy-out (nm_read_rows_pitched_async). Copies, memsets and uploads of the NEXT
// frame's set-up share copy engines with it and are served in order: started at the frame boundary, a 66 MB raster held
// the next frame's first uploads back for milliseconds (8 x B200: 36.3 ms per end-to-end step against 31.1 ms of kernels,
// profiles/r02j_e2e_debug_n8.txt). So the transfer is handed to the second stream only once the next frame's sweep is
// enqueued (launch_deep / launch_hw call this), when nothing but long kernels is in flight — or by nm_read_wait.
int issue_deferred_copy(nm_ctx* ctx) {
  if (!ctx->deferred.active) return NM_OK;
  ctx->deferred.active = false;
  NM_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_snap, 0));
  NM_CUDA(ctx, cudaMemcpy2DAsync(ctx->deferred.dst, ctx->deferred.pitch, ctx->snap.p, ctx->deferred.row_bytes, ctx->deferred.row_bytes,
                                 ctx->deferred.rows, cudaMemcpyDefault, ctx->side));
  NM_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->side));
  ctx->copy_pending = true;
  return NM_OK;
}

int finish_frame(nm_ctx* ctx) {
  if (!ctx->launched) return fail(ctx, NM_ESTATE, "no frame launched");
  if (ctx->finished) return NM_OK;
  if (int rc = read_back(ctx, ctx->h_ctr, ctx->ctr.p, CTR_COUNT * sizeof(unsigned long long))) return rc;
  unsigned long long nfix = ctx->h_ctr[CTR_FIXUP];
  if (nfix > ctx->fix_cap) return fail(ctx, NM_ESTATE, "smoothing fix-up list overflow (%llu > %llu)", nfix, ctx->fix_cap);
  if (ctx->h_ctr[CTR_AMBIG] > ctx->ambig_cap) return fail(ctx, NM_ESTATE, "ambiguous list overflow");
  if (nfix) {
    std::vector<FixupRec> recs(nfix);
    if (int rc = read_back(ctx, recs.data(), ctx->fix.p, nfix * sizeof(FixupRec))) return rc;
    std::vector<int32_t> pix(nfix);
    std::vector<float> val(nfix);
    for (size_t i = 0; i < nfix; i++) { pix[i] = recs[i].pix; val[i] = host_smoothing(recs[i].r2); }
    NM_CUDA(ctx, ctx->fixapply.ensure(nfix * 8));
    int32_t* dpix = ctx->fixapply.as<int32_t>();
    float* dval = (float*)(dpix + nfix);
    NM_CUDA(ctx, cudaMemcpyAsync(dpix, pix.data(), nfix * 4, cudaMemcpyHostToDevice, ctx->stream));
    NM_CUDA(ctx, cudaMemcpyAsync(dval, val.data(), nfix * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_apply_fixups<<<(unsigned)((nfix + 255) / 256), 256, 0, ctx->stream>>>(ctx->out.as<nm_escape>(), dpix, dval, (int)nfix);
    ctx->stats.kernel_launches++;
    NM_CUDA(ctx, cudaGetLastError());
    NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->stats.pixels = (uint64_t)ctx->W;
  ctx->stats.executed_iters = ctx->h_ctr[CTR_EXECUTED];
  ctx->stats.series_evals = ctx->h_ctr[CTR_SERIES];
  ctx->stats.skipped_pixels = ctx->h_ctr[CTR_SKIPPED];
  ctx->stats.glitched = ctx->h_ctr[CTR_REQUEUE];
  ctx->stats.rebased = ctx->h_ctr[CTR_REBASED];
  ctx->stats.checked_steps = ctx->h_ctr[CTR_CHECKED];
  ctx->stats.fixups = nfix;
  float ms = 0;
  if (ctx->kind == 1) {
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2]);
    ctx->stats.ms_k1 = ms;
  } else {
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_k2 = ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_k3 = ms;
  }
  ctx->finished = true;
  if (ctx->h_ctr[CTR_CANCEL]) return fail(ctx, NM_ECANCELLED, "frame cancelled");
  return NM_OK;
}

int launch_hw(nm_ctx* ctx) {
  K1Params p;
  p.c_re = ctx->cre.as<double>();
  p.c_im = ctx->cim.as<double>();
  p.nr = ctx->nr; p.nc = ctx->nc; p.N = ctx->N;
  p.out = ctx->out.as<nm_escape>();
  p.ctr = ctx->ctr.as<unsigned long long>();
  p.ambig = ctx->ambig.as<int32_t>();
  p.ambig_cap = ctx->ambig_cap;
  p.fix = ctx->fix.as<FixupRec>();
  p.fix_cap = ctx->fix_cap;
  p.log_bailout = ctx->log_bailout;
  long long warps_needed = (ctx->pixels + 31) / 32;
  long long blocks = (warps_needed + (K1_THREADS / 32) - 1) / (K1_THREADS / 32);
  long long maxb = (long long)ctx->sm_count * ctx->occ_k1;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  NM_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  k1_escape<<<(unsigned)blocks, K1_THREADS, 0, ctx->stream>>>(p);
  ctx->stats.kernel_launches++;
  NM_CUDA(ctx, cudaGetLastError());
  NM_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  return issue_deferred_copy(ctx);   // the previous frame's raster leaves while this one's kernel runs
}

template <int MODE, bool SCALED>
cudaError_t launch_level(nm_ctx* ctx, const K3Params& p, unsigned blocks, size_t smem) {
  k3_level<MODE, SCALED><<<blocks, K3_THREADS, smem, ctx->stream>>>(p);
  return cudaGetLastError();
}

template <bool LITERAL>
void launch_k2(nm_ctx* ctx, const K2Params& k2, unsigned blocks) {
  if (ctx->use_fe == 2) k2_series<LITERAL, 2><<<blocks, K2_THREADS, 0, ctx->stream>>>(k2);
  else if (ctx->use_fe == 1) k2_series<LITERAL, 1><<<blocks, K2_THREADS, 0, ctx->stream>>>(k2);
  else k2_series<LITERAL, 0><<<blocks, K2_THREADS, 0, ctx->stream>>>(k2);
}

FreshArrays fresh_set(nm_ctx* ctx, int which) {
  FreshArrays f;
  const size_t Wn = (size_t)(ctx->W > 0 ? ctx->W : 1);
  f.d = ctx->fa_d[which].as<double2>();
  f.j = ctx->fa_i[which].as<int32_t>();
  f.off = f.j + Wn;
  f.pix = f.j + 2 * Wn;
  f.e = f.j + 3 * Wn;
  return f;
}

int launch_deep(nm_ctx* ctx) {
  const int CH = ctx->CH;
  const int K = ctx->K;
  cudaStream_t st = ctx->stream;
  unsigned long long* qctr = ctx->qctr.as<unsigned long long>();
  // qctr layout: qcount[K+2] | head[8(K+2)] | subcount[4(K+2)] | rcount[2] | carry_count[2]
  // (head: per level, per quarter-chunk launch, one deal cursor for the 4-states-per-lane waves and one for the
  //  one-state-per-lane remainder: k3_fast.cuh)
  unsigned long long* qcount = qctr;
  unsigned long long* head = qctr + (K + 2);
  unsigned long long* subcount = qctr + 9 * (K + 2);
  unsigned long long* rcount = qctr + 13 * (K + 2);
  unsigned long long* ccount = rcount + 2;
  unsigned long long* ctr = ctx->ctr.as<unsigned long long>();

  const int G = (ctx->mode == NM_MODE_REQUEUE && ctx->opt_k3_group > 1) ? ctx->opt_k3_group : 1;
  const bool fast = G > 1;
  const int nbins = ctx->Jmax + 1;  // start indices in [0, Jmax]
  const size_t fresh_cap = (size_t)ctx->W + (size_t)4 * (nbins + 2);

  NM_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  NM_CUDA(ctx, cudaMemsetAsync(ctx->hist.p, 0, (size_t)(nbins + 2) * sizeof(unsigned), st));
  NM_CUDA(ctx, cudaMemsetAsync(&ctr[CTR_MINJ], 0xFF, sizeof(unsigned long long), st));

  CheckedParams ck;
  ck.Z = ctx->Z.as<double2>(); ck.gb = ctx->gb.as<double>(); ck.Jmax = ctx->Jmax; ck.N = ctx->N;
  ck.out = ctx->out.as<nm_escape>(); ck.ctr = ctr;
  ck.fix = ctx->fix.as<FixupRec>(); ck.fix_cap = ctx->fix_cap;
  ck.rq_pix = ctx->rq_pix.as<int32_t>(); ck.rq_iter = ctx->rq_iter.as<int32_t>();
  ck.log_bailout = ctx->log_bailout;

  // ---- K2 ---------------------------------------------------------------------------------------
  K2Params k2;
  k2.A = ctx->a.as<double2>(); k2.B = ctx->b.as<double2>(); k2.C = ctx->c.as<double2>();
  k2.Ae = ctx->aexp.as<int2>(); k2.Be = ctx->bexp.as<int2>(); k2.Ce = ctx->cexp.as<int2>();
  k2.Xhi = ctx->xhi.as<double2>(); k2.Xlo = ctx->xlo.as<double2>();
  k2.M = ctx->M; k2.N = ctx->N; k2.tol = ctx->tol;
  EpsTab eps;
  eps.re = ctx->cre.as<double>(); eps.im = ctx->cim.as<double>(); eps.nc = ctx->nc;
  eps.re_e = ctx->use_fe == 2 ? ctx->cre_e.as<int32_t>() : nullptr;
  eps.im_e = ctx->use_fe == 2 ? ctx->cim_e.as<int32_t>() : nullptr;
  const bool scaled = ctx->use_fe == 2;
  k2.eps = eps; k2.nc = ctx->nc;
  k2.pix_list = ctx->have_list ? ctx->list.as<int32_t>() : nullptr;
  k2.W = ctx->W;
  k2.cardioid_mode = ctx->cardioid_mode;
  k2.mask = ctx->mask.as<uint8_t>();
  k2.fresh = fresh_set(ctx, 0);
  k2.align4 = fast ? 1 : 0;
  k2.ck = ck;
  k2.hist = ctx->hist.as<unsigned>();
  k2.out = ctx->out.as<nm_escape>();
  k2.ctr = ctr;
  k2.fix = ctx->fix.as<FixupRec>(); k2.fix_cap = ctx->fix_cap;
  k2.log_bailout = ctx->log_bailout;
  long long b2 = (ctx->W + K2_THREADS - 1) / K2_THREADS;
  long long maxb2 = (long long)ctx->sm_count * 8;
  if (b2 > maxb2) b2 = maxb2;
  if (b2 < 1) b2 = 1;
  const bool literal = ctx->opt_k2_literal || !(ctx->tol > 0.0) || !std::isfinite(ctx->tol) || ctx->M < 2;
  if (!literal) {
    double* fb = ctx->filt.as<double>();
    const size_t Mn = (size_t)ctx->M;
    k2.f.rlog = fb; k2.f.a = fb + Mn; k2.f.b = fb + 2 * Mn; k2.f.ov = fb + 3 * Mn;
    k2.f.pmin_rlog = fb + 4 * Mn; k2.f.pmin_ov = fb + 5 * Mn; k2.f.pmin_a = fb + 6 * Mn; k2.f.pmax_b = fb + 7 * Mn;
    if (ctx->use_fe) k2_prepare<true><<<(ctx->M + 255) / 256, 256, 0, st>>>(k2.B, k2.C, k2.Be, k2.Ce, ctx->M, ctx->tol, k2.f);
    else k2_prepare<false><<<(ctx->M + 255) / 256, 256, 0, st>>>(k2.B, k2.C, k2.Be, k2.Ce, ctx->M, ctx->tol, k2.f);
    NM_CUDA(ctx, cudaGetLastError());
    {
      const int n_tiles = (ctx->M + 1023) / 1024;
      double* tile_tot = fb + 8 * Mn;  // [4][n_tiles]
      k2_prefix_tiles<<<n_tiles, 1024, 0, st>>>(ctx->M, k2.f, tile_tot, n_tiles);
      NM_CUDA(ctx, cudaGetLastError());
      if (n_tiles > 1) {
        k2_prefix_carry<<<1, 1024, 0, st>>>(tile_tot, n_tiles);
        NM_CUDA(ctx, cudaGetLastError());
        k2_prefix_apply<<<n_tiles, 1024, 0, st>>>(ctx->M, k2.f, tile_tot, n_tiles);
        NM_CUDA(ctx, cudaGetLastError());
        ctx->stats.kernel_launches += 2;
      }
    }
    ctx->stats.kernel_launches += 2;
    launch_k2<false>(ctx, k2, (unsigned)b2);
  } else {
    launch_k2<true>(ctx, k2, (unsigned)b2);
  }
  NM_CUDA(ctx, cudaGetLastError());
  ctx->stats.kernel_launches++;
  NM_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));

  // ---- K3 ---------------------------------------------------------------------------------------
  K3Params p;
  p.Z = ctx->Z.as<double2>(); p.ghi = ctx->ghi.as<int32_t>(); p.gb = ctx->gb.as<double>();
  p.Z2 = ctx->Z2.as<double2>(); p.filt = ctx->k3filt.as<int4>(); p.esc_hi = ctx->esc_hi.as<int32_t>();
  p.seg_hi = ctx->seg_hi.as<int32_t>();
  p.Jmax = ctx->Jmax; p.N = ctx->N; p.CH = CH;
  p.eps = eps; p.nc = ctx->nc;
  p.fresh_ids = ctx->fresh.as<int32_t>();
  p.out = ctx->out.as<nm_escape>();
  p.ctr = ctr;
  p.fix = ctx->fix.as<FixupRec>(); p.fix_cap = ctx->fix_cap;
  p.rq_pix = ctx->rq_pix.as<int32_t>(); p.rq_iter = ctx->rq_iter.as<int32_t>();
  p.log_bailout = ctx->log_bailout;

  // k3_level: Z + glitch-bound words; k3_fast: 2Z + filter entries + escape words + the slot records (k3_fast.cuh)
  const size_t smem = fast ? k3f_table_bytes(CH) + (G == 4 ? K3Slots<4>::bytes() : K3Slots<2>::bytes())
                           : (size_t)(CH + 4) * (sizeof(double2) + sizeof(double));
  int occ = (scaled ? ctx->occ_k3s : ctx->occ_k3)[ctx->mode == NM_MODE_REBASE ? 1 : 0];
  if (G == 2) occ = (scaled ? ctx->occ_k3fs : ctx->occ_k3f)[0];
  if (G == 4) occ = (scaled ? ctx->occ_k3fs : ctx->occ_k3f)[1];
  const unsigned blocks = (unsigned)(ctx->sm_count * occ);
  NM_CUDA(ctx, cudaMemsetAsync(ccount, 0, 2 * sizeof(unsigned long long), st));

  if (ctx->mode == NM_MODE_DD) {   // exact mode: the listed samples once more, phase 3 in double-double (k3_dd.cuh)
    DDParams q;
    q.Xhi = ctx->xhi.as<double2>(); q.Xlo = ctx->xlo.as<double2>();
    q.M = ctx->M; q.Jmax = ctx->Jmax; q.N = ctx->N; q.nc = ctx->nc;
    q.eps_re = ctx->cre.as<double>(); q.eps_im = ctx->cim.as<double>();
    q.eps_re_lo = ctx->have_eps_lo ? ctx->cre_lo.as<double>() : nullptr;
    q.eps_im_lo = ctx->have_eps_lo ? ctx->cim_lo.as<double>() : nullptr;
    q.out = ctx->out.as<nm_escape>(); q.ctr = ctr;
    q.fix = ctx->fix.as<FixupRec>(); q.fix_cap = ctx->fix_cap; q.log_bailout = ctx->log_bailout;
    long long bd = (ctx->W + K3_DD_THREADS - 1) / K3_DD_THREADS;
    if (bd > (long long)ctx->sm_count * 16) bd = (long long)ctx->sm_count * 16;
    if (bd < 1) bd = 1;
    k3_dd<<<(unsigned)bd, K3_DD_THREADS, 0, st>>>(q, fresh_set(ctx, 0), ctx->W);
    NM_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
    ctx->stats.sweeps++;
    NM_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    return NM_OK;
  }

  // Lowest table index any state of the coming sweep starts at: the levels below it have no work and
  // are not launched (an empty launch still costs ~4 us; M/1024 of them per sweep add up on small
  // frames). One 8-byte read-back + sync after K2; later sweeps piggy-back on the sweep-end read.
  if (int rc = read_back(ctx, ctx->h_ctr + 2, &ctr[CTR_MINJ], sizeof(unsigned long long))) return rc;
  unsigned long long min_j = ctx->h_ctr[2];

  static const bool debug_levels = getenv("NM_DEBUG_LEVELS") != nullptr;
  cudaEvent_t dbg_ev = nullptr;
  unsigned long long dbg_prev = 0;
  if (debug_levels) {
    cudaEventCreate(&dbg_ev);
    cudaMemcpyAsync(&dbg_prev, &ctr[CTR_EXECUTED], 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(dbg_ev, st);
  }

  for (int sweep = 0;; ++sweep) {
    const int par = sweep & 1;
    // Few states (left): run them to completion in one launch instead of sweeping every level (k3_finish.cuh).
    // Sweep 0: K2's hand-over (W entries, finished ones marked); later sweeps of the fast path: the carried states.
    const unsigned long long n_states = sweep == 0 ? (unsigned long long)ctx->W : ctx->h_ctr[0];
    if ((sweep == 0 || fast) && n_states <= (unsigned long long)ctx->opt_k3_finish_max) {
      unsigned fb = (unsigned)((n_states + K3_FINISH_THREADS - 1) / K3_FINISH_THREADS);
      if (fb < 1) fb = 1;
      const unsigned long long* cnt = sweep == 0 ? nullptr : &ccount[par];
      const FreshArrays fa = fresh_set(ctx, par);
      const long long nmax = (long long)ctx->W;
      if (ctx->mode == NM_MODE_REBASE) {
        if (scaled) k3_finish<NM_MODE_REBASE, true><<<fb, K3_FINISH_THREADS, 0, st>>>(ck, eps, fa, cnt, nmax);
        else k3_finish<NM_MODE_REBASE, false><<<fb, K3_FINISH_THREADS, 0, st>>>(ck, eps, fa, cnt, nmax);
      } else {
        if (scaled) k3_finish<NM_MODE_REQUEUE, true><<<fb, K3_FINISH_THREADS, 0, st>>>(ck, eps, fa, cnt, nmax);
        else k3_finish<NM_MODE_REQUEUE, false><<<fb, K3_FINISH_THREADS, 0, st>>>(ck, eps, fa, cnt, nmax);
      }
      NM_CUDA(ctx, cudaGetLastError());
      ctx->stats.kernel_launches++;
      ctx->stats.sweeps++;
      break;
    }
    int kstart = 0;
    if (fast || sweep == 0) kstart = min_j >= (unsigned long long)(K * CH) ? K : (int)(min_j / (unsigned long long)CH);
    // chunk-sorted "fresh" list of this sweep: K2's hand-over (sweep 0) or the carried states
    const bool have_fresh = fast || sweep == 0;
    if (have_fresh) {
      k2_scan<<<1, 1024, 0, st>>>(ctx->hist.as<unsigned>(), ctx->offs.as<unsigned>(), ctx->cursor.as<unsigned>(), nbins, (unsigned)G);
      NM_CUDA(ctx, cudaGetLastError());
      NM_CUDA(ctx, cudaMemsetAsync(ctx->fresh.p, 0xFF, fresh_cap * sizeof(int32_t), st));
      k2_scatter<<<(unsigned)b2, 256, 0, st>>>(fresh_set(ctx, par).j, ctx->W, sweep == 0 ? nullptr : &ccount[par],
                                                ctx->cursor.as<unsigned>(), ctx->fresh.as<int32_t>());
      NM_CUDA(ctx, cudaGetLastError());
      ctx->stats.kernel_launches += 2;
      NM_CUDA(ctx, cudaMemsetAsync(ctx->hist.p, 0, (size_t)(nbins + 2) * sizeof(unsigned), st));
    }
    NM_CUDA(ctx, cudaMemsetAsync(qcount, 0, 13 * (K + 2) * sizeof(unsigned long long), st));
    NM_CUDA(ctx, cudaMemsetAsync(&rcount[par ^ 1], 0, sizeof(unsigned long long), st));
    NM_CUDA(ctx, cudaMemsetAsync(&ccount[par ^ 1], 0, sizeof(unsigned long long), st));
    NM_CUDA(ctx, cudaMemsetAsync(&ctr[CTR_EVENTS], 0, sizeof(unsigned long long), st));
    NM_CUDA(ctx, cudaMemsetAsync(&ctr[CTR_MINJ], 0xFF, sizeof(unsigned long long), st));
    p.fresh = fresh_set(ctx, par);
    for (int k = kstart; k < K; ++k) {
      p.k = k;
      if (k == 0) {
        p.cur = ctx->rq[par].as<PixState>();
        p.cur_count = (fast || sweep == 0) ? nullptr : &rcount[par];
      } else {
        p.cur = ctx->q[k & 1].as<PixState>();
        p.cur_count = &qcount[k];
      }
      p.next = ctx->q[(k + 1) & 1].as<PixState>();
      p.next_count = &qcount[k + 1];
      p.restart = ctx->rq[par ^ 1].as<PixState>();
      p.restart_count = &rcount[par ^ 1];
      p.head = &head[2 * K3F_SUBS * k];
      p.fresh_off = have_fresh ? ctx->offs.as<unsigned>() : nullptr;
      p.qcount = qcount;
      p.tmp[0] = ctx->rq[0].as<PixState>(); p.tmp[1] = ctx->rq[1].as<PixState>();   // idle in the fast path
      p.split_min = ctx->opt_k3_split > 1 ? (unsigned long long)ctx->opt_k3_split : K3F_SPLIT_MIN;
      p.sub_count = (ctx->opt_k3_split && (unsigned long long)ctx->W >= p.split_min) ? &subcount[K3F_SUBS * k] : nullptr;
      cudaError_t e = cudaSuccess;
      PixState* evq = ctx->events.as<PixState>();
      if (fast) {   // K3F_SUBS launches per level; all but the first return at once unless the level is split (K3Work)
        // (a frame that cannot reach the split minimum is spared the three empty launches per level)
        const int subs = (ctx->opt_k3_split && (unsigned long long)ctx->W >= p.split_min) ? K3F_SUBS : 1;
        for (int sub = 0; sub < subs && e == cudaSuccess; ++sub) {
          p.sub = sub;
          if (G == 4 && scaled) k3_fast<4, true><<<blocks, K3F_THREADS, smem, st>>>(p, evq);
          else if (G == 4) k3_fast<4, false><<<blocks, K3F_THREADS, smem, st>>>(p, evq);
          else if (scaled) k3_fast<2, true><<<blocks, K3F_THREADS, smem, st>>>(p, evq);
          else k3_fast<2, false><<<blocks, K3F_THREADS, smem, st>>>(p, evq);
          e = cudaGetLastError();
          if (sub) ctx->stats.kernel_launches++;
        }
      }
      else if (scaled) e = ctx->mode == NM_MODE_REBASE ? launch_level<NM_MODE_REBASE, true>(ctx, p, blocks, smem)
                                                       : launch_level<NM_MODE_REQUEUE, true>(ctx, p, blocks, smem);
      else e = ctx->mode == NM_MODE_REBASE ? launch_level<NM_MODE_REBASE, false>(ctx, p, blocks, smem)
                                           : launch_level<NM_MODE_REQUEUE, false>(ctx, p, blocks, smem);
      if (e != cudaSuccess) return fail(ctx, NM_ECUDA, "k3 level launch: %s", cudaGetErrorString(e));
      ctx->stats.kernel_launches++;
      if (debug_levels) {  // NM_DEBUG_LEVELS=1: per-level device time and executed iterations (serialises the frame)
        cudaEventRecord(ctx->ev[3], st);
        unsigned long long ex = 0, nx = 0;
        cudaMemcpyAsync(&ex, &ctr[CTR_EXECUTED], 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&nx, &qcount[k + 1], 8, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        float ms = 0;
        cudaEventElapsedTime(&ms, dbg_ev, ctx->ev[3]);
        fprintf(stderr, "nm level sweep %d k %3d: %8.3f ms  executed +%llu  (%.1f Giter/s)  -> next %llu\n", sweep, k, ms,
                ex - dbg_prev, ms > 0 ? (double)(ex - dbg_prev) / ms * 1e-6 : 0.0, nx);
        dbg_prev = ex;
        cudaEventRecord(dbg_ev, st);
      }
    }
    if (int rc = issue_deferred_copy(ctx)) return rc;   // the previous frame's raster leaves while this sweep runs
    if (fast) {
      // finish what the branch-free kernel exported — escapes, glitches, limits, rebases, false alarms — one thread per
      // state, to the end (k3_finish.cuh); nothing is carried into another sweep
      const unsigned fbk = (unsigned)(ctx->sm_count * 16);
      const FreshArrays none = fresh_set(ctx, par ^ 1);
      if (scaled)
        k3_finish<NM_MODE_REQUEUE, true, true><<<fbk, K3_FINISH_THREADS, 0, st>>>(ck, eps, none, &ctr[CTR_EVENTS], 0, ctx->events.as<PixState>());
      else
        k3_finish<NM_MODE_REQUEUE, false, true><<<fbk, K3_FINISH_THREADS, 0, st>>>(ck, eps, none, &ctr[CTR_EVENTS], 0, ctx->events.as<PixState>());
      NM_CUDA(ctx, cudaGetLastError());
      ctx->stats.kernel_launches++;
    }
    ctx->stats.sweeps++;
    // (carried / restart count, cancel flag, lowest start index of the next sweep: one kernel-written read-back)
    k_three_to_host<<<1, 32, 0, st>>>(fast ? &ccount[par ^ 1] : &rcount[par ^ 1], &ctr[CTR_CANCEL], &ctr[CTR_MINJ],
                                      (volatile unsigned long long*)ctx->h_small);
    NM_CUDA(ctx, cudaGetLastError());
    NM_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(ctx->h_ctr, ctx->h_small, 3 * sizeof(unsigned long long));
    min_j = ctx->h_ctr[2];
    if (ctx->h_ctr[0] == 0 || ctx->h_ctr[1] != 0) break;
  }
  if (dbg_ev) cudaEventDestroy(dbg_ev);
  NM_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
  return NM_OK;
}

int run_resolve(nm_ctx* ctx, const nm_escape* dgrid, int nr, int nc, const uint8_t* pal_rgb, int n_pal, int N, int sc,
                int smooth, uint8_t* rgb_out) {
  if (sc < 1 || nr % sc || nc % sc || n_pal < 1 || !pal_rgb || !rgb_out) return fail(ctx, NM_EINVAL, "nm_resolve: bad arguments");
  NM_CUDA(ctx, ctx->pal.ensure((size_t)3 * n_pal));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->pal.p, pal_rgb, (size_t)3 * n_pal, cudaMemcpyDefault, ctx->stream));
  size_t obytes = (size_t)(nr / sc) * (nc / sc) * 3;
  NM_CUDA(ctx, ctx->rgb.ensure(obytes));
  K4Params p;
  p.grid = dgrid; p.nr = nr; p.nc = nc; p.pal = ctx->pal.as<uint8_t>();
  p.n_pal = n_pal; p.N = N; p.sc = sc; p.smooth = smooth; p.rgb = ctx->rgb.as<uint8_t>();
  long long total = (long long)(nr / sc) * (nc / sc);
  long long blocks = (total + 255) / 256;
  long long maxb = (long long)ctx->sm_count * 16;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  NM_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  k4_resolve<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p);
  NM_CUDA(ctx, cudaGetLastError());
  cudaEvent_t evend = ctx->ev[1];
  if (ctx->kind != 0) {  // ev[1] belongs to the frame; use a scratch event
    NM_CUDA(ctx, cudaEventCreate(&evend));
  }
  NM_CUDA(ctx, cudaEventRecord(evend, ctx->stream));
  ctx->stats.kernel_launches++;
  NM_CUDA(ctx, cudaMemcpyAsync(rgb_out, ctx->rgb.p, obytes, cudaMemcpyDefault, ctx->stream));
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[3], evend);
  ctx->stats.ms_k4 = ms;
  if (evend != ctx->ev[1]) cudaEventDestroy(evend);
  return NM_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* nm_version(void) { return "newman_b200 0.1 (sm_100a)"; }

const char* nm_last_error(const nm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int nm_create(int device, nm_ctx** out) {
  if (!out) return NM_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, NM_ENODEV, "no CUDA device (%s); newman_b200 has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, NM_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
  nm_ctx* ctx = new nm_ctx();
  ctx->device = device;
#define NM_CREATE_CUDA(call)                                                       \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      int rc = fail(nullptr, NM_ECUDA, "%s: %s", #call, cudaGetErrorString(e__));  \
      delete ctx;                                                                  \
      return rc;                                                                   \
    }                                                                              \
  } while (0)
  NM_CREATE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  NM_CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    int rc = fail(nullptr, NM_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    delete ctx;
    return rc;
  }
  ctx->sm_count = prop.multiProcessorCount;
  NM_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking));
  NM_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
  ctx->stream = ctx->own;
  for (int i = 0; i < 4; i++) NM_CREATE_CUDA(cudaEventCreate(&ctx->ev[i]));
  NM_CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_snap, cudaEventDisableTiming));
  NM_CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
  NM_CREATE_CUDA(ctx->ctr.ensure(CTR_COUNT * sizeof(unsigned long long)));
  NM_CREATE_CUDA(cudaMemset(ctx->ctr.p, 0, CTR_COUNT * sizeof(unsigned long long)));
  NM_CREATE_CUDA(cudaMallocHost((void**)&ctx->h_ctr, CTR_COUNT * sizeof(unsigned long long)));
  NM_CREATE_CUDA(cudaMallocHost((void**)&ctx->h_flag, sizeof(unsigned long long)));
  NM_CREATE_CUDA(cudaMallocHost((void**)&ctx->h_small, NM_SMALL_D2H));
  *ctx->h_flag = 1ULL;
  ctx->log_bailout = log(1024.0);
  const size_t smem = (size_t)(ctx->CH + 4) * (sizeof(double2) + sizeof(double));
#define NM_K3_SETUP(fn, threads, occ_out)                                                                     \
  NM_CREATE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
  NM_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&(occ_out), fn, threads, smem));                \
  if ((occ_out) < 1) (occ_out) = 1;
  NM_K3_SETUP((k3_level<NM_MODE_REQUEUE, false>), K3_THREADS, ctx->occ_k3[0]);
  NM_K3_SETUP((k3_level<NM_MODE_REBASE, false>), K3_THREADS, ctx->occ_k3[1]);
  NM_K3_SETUP((k3_level<NM_MODE_REQUEUE, true>), K3_THREADS, ctx->occ_k3s[0]);
  NM_K3_SETUP((k3_level<NM_MODE_REBASE, true>), K3_THREADS, ctx->occ_k3s[1]);
#undef NM_K3_SETUP
#define NM_K3_SETUP(fn, P, occ_out)                                                                             \
  {                                                                                                             \
    const size_t smem_f = k3f_table_bytes(ctx->CH) + K3Slots<P>::bytes();   /* k3_fast.cuh */                   \
    NM_CREATE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));         \
    NM_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&(occ_out), fn, K3F_THREADS, smem_f));         \
    if ((occ_out) < 1) (occ_out) = 1;                                                                           \
  }
  NM_K3_SETUP((k3_fast<2, false>), 2, ctx->occ_k3f[0]);
  NM_K3_SETUP((k3_fast<4, false>), 4, ctx->occ_k3f[1]);
  NM_K3_SETUP((k3_fast<2, true>), 2, ctx->occ_k3fs[0]);
  NM_K3_SETUP((k3_fast<4, true>), 4, ctx->occ_k3fs[1]);
  if (const char* e = getenv("NM_K3F_OCC")) {   // experiment: fewer resident CTAs per SM for k3_fast
    const int cap = atoi(e);
    if (cap >= 1)
      for (int q = 0; q < 2; ++q) {
        if (ctx->occ_k3f[q] > cap) ctx->occ_k3f[q] = cap;
        if (ctx->occ_k3fs[q] > cap) ctx->occ_k3fs[q] = cap;
      }
  }
#undef NM_K3_SETUP
  NM_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_k1, k1_escape, K1_THREADS, 0));
#undef NM_CREATE_CUDA
  if (ctx->occ_k1 < 1) ctx->occ_k1 = 1;
  if (const char* fm = getenv("NM_K3_FINISH_MAX")) { const long long v = atoll(fm); if (v >= 0) ctx->opt_k3_finish_max = v; }
  memset(&ctx->stats, 0, sizeof ctx->stats);
  *out = ctx;
  return NM_OK;
}

void nm_destroy(nm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->own);
  DevBuf* bufs[] = {&ctx->out, &ctx->cre, &ctx->cim, &ctx->ctr, &ctx->ambig, &ctx->fix, &ctx->fixapply, &ctx->Z, &ctx->ghi, &ctx->Z2, &ctx->k3filt, &ctx->esc_hi, &ctx->seg_hi, &ctx->eps_max,
                    &ctx->gb, &ctx->xhi, &ctx->xlo, &ctx->a, &ctx->b, &ctx->c, &ctx->mask, &ctx->list, &ctx->fa_d[0], &ctx->fa_d[1], &ctx->fa_i[0], &ctx->fa_i[1],
                    &ctx->hist, &ctx->offs, &ctx->cursor, &ctx->fresh, &ctx->q[0], &ctx->q[1], &ctx->rq[0], &ctx->rq[1],
                    &ctx->qctr, &ctx->rq_pix, &ctx->rq_iter, &ctx->pal, &ctx->rgb, &ctx->gridtmp, &ctx->filt, &ctx->events, &ctx->snap, &ctx->aexp, &ctx->bexp, &ctx->cexp, &ctx->cre_e, &ctx->cim_e,
                    &ctx->palpar, &ctx->paldev, &ctx->vprev, &ctx->vnext, &ctx->vout, &ctx->kept, &ctx->difflist, &ctx->cre_lo, &ctx->cim_lo};
  for (DevBuf* b : bufs) b->release();
  if (ctx->side) cudaStreamSynchronize(ctx->side);
  for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->ev_snap) cudaEventDestroy(ctx->ev_snap);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
  if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  if (ctx->own) cudaStreamDestroy(ctx->own);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  delete ctx;
}

int nm_set_stream(nm_ctx* ctx, void* cuda_stream) {
  if (!ctx) return NM_EINVAL;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own;
  return NM_OK;
}

int nm_get_stream(nm_ctx* ctx, void** cuda_stream) {
  if (!ctx || !cuda_stream) return NM_EINVAL;
  *cuda_stream = (void*)ctx->stream;
  return NM_OK;
}

int nm_set_option(nm_ctx* ctx, int key, int value) {
  if (!ctx) return NM_EINVAL;
  switch (key) {
    case NM_OPT_K2_LITERAL: ctx->opt_k2_literal = value ? 1 : 0; return NM_OK;
    case NM_OPT_K3_GROUP:
      if (value != 0 && value != 1 && value != 2 && value != 4) return fail(ctx, NM_EINVAL, "NM_OPT_K3_GROUP must be 0, 1, 2 or 4");
      ctx->opt_k3_group = value;
      return NM_OK;
    case NM_OPT_K3_SPLIT:
      if (value < 0) return fail(ctx, NM_EINVAL, "NM_OPT_K3_SPLIT must be >= 0");
      ctx->opt_k3_split = value;
      return NM_OK;
    case NM_OPT_K3_FINISH_MAX:
      if (value < 0) return fail(ctx, NM_EINVAL, "NM_OPT_K3_FINISH_MAX must be >= 0");
      ctx->opt_k3_finish_max = value;
      return NM_OK;
    default: return fail(ctx, NM_EINVAL, "nm_set_option: unknown key %d", key);
  }
}

int nm_sync(nm_ctx* ctx) {
  if (!ctx) return NM_EINVAL;
  if (set_device(ctx)) return NM_ECUDA;
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NM_OK;
}

int nm_cancel(nm_ctx* ctx) {
  if (!ctx) return NM_EINVAL;
  // side stream: must not queue behind the frame it is cancelling
  NM_CUDA(ctx, cudaMemcpyAsync(&ctx->ctr.as<unsigned long long>()[CTR_CANCEL], ctx->h_flag, sizeof(unsigned long long),
                               cudaMemcpyHostToDevice, ctx->side));
  return NM_OK;
}

int nm_frame_hw(nm_ctx* ctx, const double* c_re, int nc, const double* c_im, int nr, int N) {
  if (!ctx) return NM_EINVAL;
  if (!c_re || !c_im || nr < 1 || nc < 1 || N < 0 || (long long)nr * nc > 0x7fffffffLL)
    return fail(ctx, NM_EINVAL, "nm_frame_hw: bad arguments");
  if (int rc = set_device(ctx)) return rc;
  ctx->kind = 1; ctx->nr = nr; ctx->nc = nc; ctx->N = N;
  ctx->pixels = (long long)nr * nc; ctx->W = ctx->pixels; ctx->have_list = false;
  ctx->launched = ctx->finished = false;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  NM_CUDA(ctx, ctx->out.ensure((size_t)ctx->pixels * sizeof(nm_escape)));
  NM_CUDA(ctx, ctx->cre.ensure((size_t)nc * sizeof(double)));
  NM_CUDA(ctx, ctx->cim.ensure((size_t)nr * sizeof(double)));
  if (int rc = size_lists(ctx)) return rc;
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->cre.p, c_re, (size_t)nc * sizeof(double), cudaMemcpyDefault, ctx->stream));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->cim.p, c_im, (size_t)nr * sizeof(double), cudaMemcpyDefault, ctx->stream));
  return NM_OK;
}

int nm_frame_deep(nm_ctx* ctx, const nm_deep_tables* t, const double* eps_re, int nc, const double* eps_im, int nr,
                  int cardioid_mode, const uint8_t* cardioid_mask, const int32_t* pix_list, int64_t n_list, int mode) {
  if (!ctx) return NM_EINVAL;
  if (!t || !eps_re || !eps_im || nr < 1 || nc < 1 || (long long)nr * nc > 0x7fffffffLL)
    return fail(ctx, NM_EINVAL, "nm_frame_deep: bad arguments");
  if (t->M < 1 || t->N < t->M || !t->x_hi || !t->x_lo || !t->a || !t->b || !t->c)
    return fail(ctx, NM_EINVAL, "nm_frame_deep: bad tables (M=%d N=%d)", t->M, t->N);
  if (!t->has_escape && t->M != t->N && cardioid_mode != NM_CARDIOID_ALL)
    return fail(ctx, NM_EINVAL, "nm_frame_deep: orbit shorter than N needs its escaped iterate (has_escape)");
  if (cardioid_mode == NM_CARDIOID_MASK && !cardioid_mask) return fail(ctx, NM_EINVAL, "cardioid mask missing");
  if (mode != NM_MODE_REQUEUE && mode != NM_MODE_REBASE && mode != NM_MODE_DD) return fail(ctx, NM_EINVAL, "bad mode");
  if (mode == NM_MODE_DD && (!pix_list || t->eps_re_exp || t->eps_im_exp))
    return fail(ctx, NM_EINVAL, "NM_MODE_DD refines listed samples of frames with plain (not scaled) delta states");
  if (int rc = set_device(ctx)) return rc;
  // A pixel-list frame writes only the listed samples. On top of a deep frame of the same size the
  // others keep their values (secondary-reference rounds); otherwise (probe search: the candidate
  // samples of a view that has no raster yet) the rest of the raster is cleared.
  const bool fresh_raster = pix_list && (ctx->kind != 2 || ctx->nr != nr || ctx->nc != nc);

  ctx->kind = 2; ctx->nr = nr; ctx->nc = nc; ctx->N = t->N;
  ctx->pixels = (long long)nr * nc;
  ctx->have_list = pix_list != nullptr;
  ctx->W = pix_list ? (long long)n_list : ctx->pixels;
  ctx->launched = ctx->finished = false;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  ctx->M = t->M; ctx->has_escape = t->has_escape ? 1 : 0;
  ctx->use_fe = (t->a_exp && t->b_exp && t->c_exp) ? 1 : 0;
  if (!ctx->use_fe && (t->a_exp || t->b_exp || t->c_exp)) return fail(ctx, NM_EINVAL, "floatexp tables need all of a_exp, b_exp, c_exp");
  if (t->eps_re_exp || t->eps_im_exp) {
    if (!ctx->use_fe || !t->eps_re_exp || !t->eps_im_exp)
      return fail(ctx, NM_EINVAL, "floatexp eps needs both eps_re_exp and eps_im_exp and floatexp series tables");
    ctx->use_fe = 2;
  }
  ctx->Jmax = t->M + ctx->has_escape;
  ctx->K = (ctx->Jmax + ctx->CH - 1) / ctx->CH;
  ctx->tol = t->tol; ctx->gtol = t->glitch_tol;
  ctx->mode = mode; ctx->cardioid_mode = cardioid_mode;
  const int M = t->M, J1 = ctx->Jmax + 1, K = ctx->K;
  const size_t Wn = (size_t)(ctx->W > 0 ? ctx->W : 1);

  NM_CUDA(ctx, ctx->out.ensure((size_t)ctx->pixels * sizeof(nm_escape)));
  if (fresh_raster) NM_CUDA(ctx, cudaMemsetAsync(ctx->out.p, 0, (size_t)ctx->pixels * sizeof(nm_escape), ctx->stream));
  NM_CUDA(ctx, ctx->cre.ensure((size_t)nc * sizeof(double)));
  NM_CUDA(ctx, ctx->cim.ensure((size_t)nr * sizeof(double)));
  if (int rc = size_lists(ctx)) return rc;
  NM_CUDA(ctx, ctx->Z.ensure((size_t)(J1 + 8) * sizeof(double2)));
  NM_CUDA(ctx, ctx->gb.ensure((size_t)(J1 + 8) * sizeof(double)));
  NM_CUDA(ctx, ctx->ghi.ensure((size_t)(J1 + 8) * sizeof(int32_t)));
  NM_CUDA(ctx, ctx->Z2.ensure((size_t)(J1 + 8) * sizeof(double2)));
  NM_CUDA(ctx, ctx->k3filt.ensure((size_t)(J1 + 8) * sizeof(K3Filt)));
  NM_CUDA(ctx, ctx->esc_hi.ensure((size_t)(J1 + 8) * sizeof(int32_t)));
  NM_CUDA(ctx, ctx->seg_hi.ensure((size_t)(J1 / 16 + 8) * sizeof(int32_t)));
  NM_CUDA(ctx, ctx->eps_max.ensure(sizeof(double)));
  NM_CUDA(ctx, ctx->xhi.ensure((size_t)(M + 1) * sizeof(double2)));
  NM_CUDA(ctx, ctx->xlo.ensure((size_t)M * sizeof(double2)));
  NM_CUDA(ctx, ctx->a.ensure((size_t)M * sizeof(double2)));
  NM_CUDA(ctx, ctx->b.ensure((size_t)M * sizeof(double2)));
  NM_CUDA(ctx, ctx->c.ensure((size_t)M * sizeof(double2)));
  NM_CUDA(ctx, ctx->filt.ensure(((size_t)M * 8 + 4 * ((size_t)M / 1024 + 1)) * sizeof(double)));
  for (int i = 0; i < 2; i++) {
    NM_CUDA(ctx, ctx->fa_d[i].ensure(Wn * sizeof(double2)));
    NM_CUDA(ctx, ctx->fa_i[i].ensure(Wn * 4 * sizeof(int32_t)));
  }
  NM_CUDA(ctx, ctx->fresh.ensure((Wn + (size_t)4 * (J1 + 2)) * sizeof(int32_t)));
  NM_CUDA(ctx, ctx->hist.ensure((size_t)(J1 + 4) * sizeof(unsigned)));
  NM_CUDA(ctx, ctx->offs.ensure((size_t)(J1 + 4) * sizeof(unsigned)));
  NM_CUDA(ctx, ctx->cursor.ensure((size_t)(J1 + 4) * sizeof(unsigned)));
  for (int i = 0; i < 2; i++) {
    NM_CUDA(ctx, ctx->q[i].ensure(Wn * sizeof(PixState)));
    NM_CUDA(ctx, ctx->rq[i].ensure(Wn * sizeof(PixState)));
  }
  NM_CUDA(ctx, ctx->qctr.ensure((size_t)(13 * (K + 2) + 4) * sizeof(unsigned long long)));
  NM_CUDA(ctx, ctx->rq_pix.ensure(Wn * sizeof(int32_t)));
  NM_CUDA(ctx, ctx->rq_iter.ensure(Wn * sizeof(int32_t)));
  NM_CUDA(ctx, ctx->events.ensure(Wn * sizeof(PixState)));

  cudaStream_t s = ctx->stream;
  NM_CUDA(ctx, cudaMemsetAsync(ctx->Z.p, 0, (size_t)(J1 + 8) * sizeof(double2), s));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->Z.as<double2>() + 1, t->x_hi, (size_t)(M + ctx->has_escape) * sizeof(double2), cudaMemcpyDefault, s));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->xhi.p, t->x_hi, (size_t)(M + ctx->has_escape) * sizeof(double2), cudaMemcpyDefault, s));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->xlo.p, t->x_lo, (size_t)M * sizeof(double2), cudaMemcpyDefault, s));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->a.p, t->a, (size_t)M * sizeof(double2), cudaMemcpyDefault, s));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->b.p, t->b, (size_t)M * sizeof(double2), cudaMemcpyDefault, s));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->c.p, t->c, (size_t)M * sizeof(double2), cudaMemcpyDefault, s));
  if (ctx->use_fe) {
    NM_CUDA(ctx, ctx->aexp.ensure((size_t)M * sizeof(int2)));
    NM_CUDA(ctx, ctx->bexp.ensure((size_t)M * sizeof(int2)));
    NM_CUDA(ctx, ctx->cexp.ensure((size_t)M * sizeof(int2)));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->aexp.p, t->a_exp, (size_t)M * sizeof(int2), cudaMemcpyDefault, s));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->bexp.p, t->b_exp, (size_t)M * sizeof(int2), cudaMemcpyDefault, s));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->cexp.p, t->c_exp, (size_t)M * sizeof(int2), cudaMemcpyDefault, s));
  }
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->cre.p, eps_re, (size_t)nc * sizeof(double), cudaMemcpyDefault, s));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->cim.p, eps_im, (size_t)nr * sizeof(double), cudaMemcpyDefault, s));
  if (ctx->use_fe == 2) {
    NM_CUDA(ctx, ctx->cre_e.ensure((size_t)nc * sizeof(int32_t)));
    NM_CUDA(ctx, ctx->cim_e.ensure((size_t)nr * sizeof(int32_t)));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->cre_e.p, t->eps_re_exp, (size_t)nc * sizeof(int32_t), cudaMemcpyDefault, s));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->cim_e.p, t->eps_im_exp, (size_t)nr * sizeof(int32_t), cudaMemcpyDefault, s));
  }
  ctx->have_eps_lo = mode == NM_MODE_DD && t->eps_re_lo && t->eps_im_lo;
  if (ctx->have_eps_lo) {
    NM_CUDA(ctx, ctx->cre_lo.ensure((size_t)nc * sizeof(double)));
    NM_CUDA(ctx, ctx->cim_lo.ensure((size_t)nr * sizeof(double)));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->cre_lo.p, t->eps_re_lo, (size_t)nc * sizeof(double), cudaMemcpyDefault, s));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->cim_lo.p, t->eps_im_lo, (size_t)nr * sizeof(double), cudaMemcpyDefault, s));
  }
  if (cardioid_mode == NM_CARDIOID_MASK) {
    NM_CUDA(ctx, ctx->mask.ensure((size_t)ctx->pixels));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->mask.p, cardioid_mask, (size_t)ctx->pixels, cudaMemcpyDefault, s));
  }
  if (pix_list) {
    NM_CUDA(ctx, ctx->list.ensure(Wn * sizeof(int32_t)));
    NM_CUDA(ctx, cudaMemcpyAsync(ctx->list.p, pix_list, (size_t)ctx->W * sizeof(int32_t), cudaMemcpyDefault, s));
  }
  {
    // K3 iterates against the orbit rounded to nearest (k_round_orbit); K2's phase 2 keeps the truncated hi/lo pair
    if (!(t->flags & NM_TABLES_ORBIT_TRUNCATED)) {   // (the exact mode's probe rendering keeps the truncated orbit)
      k_round_orbit<<<(M + 255) / 256, 256, 0, s>>>(ctx->Z.as<double2>() + 1, ctx->xhi.as<double2>(), ctx->xlo.as<double2>(), M);
      NM_CUDA(ctx, cudaGetLastError());
      ctx->stats.kernel_launches++;
    }
    int pad_n = J1 + 8;
    k_glitch_bounds<<<(pad_n + 255) / 256, 256, 0, s>>>(ctx->Z.as<double2>(), ctx->gb.as<double>(), ctx->ghi.as<int32_t>(),
                                                        ctx->Z2.as<double2>(), ctx->k3filt.as<K3Filt>(), ctx->esc_hi.as<int32_t>(), J1,
                                                        pad_n, ctx->gtol, ctx->has_escape);
    NM_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
    // quiet bounds of k3_fast's segments: they depend on the orbit AND on the frame's largest pixel offset
    EpsTab et;
    et.re = ctx->cre.as<double>(); et.im = ctx->cim.as<double>(); et.nc = nc;
    et.re_e = ctx->use_fe == 2 ? ctx->cre_e.as<int32_t>() : nullptr;
    et.im_e = ctx->use_fe == 2 ? ctx->cim_e.as<int32_t>() : nullptr;
    const int n_seg = J1 / 16 + 2;
    k_eps_max<<<1, 1024, 0, s>>>(et, nr, ctx->eps_max.as<double>());
    NM_CUDA(ctx, cudaGetLastError());
    k_seg_bounds<<<(n_seg + 127) / 128, 128, 0, s>>>(ctx->Z.as<double2>(), ctx->gb.as<double>(), ctx->Jmax, ctx->eps_max.as<double>(),
                                                     ctx->seg_hi.as<int32_t>(), n_seg);
    NM_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 2;
  }
  return NM_OK;
}

int nm_launch(nm_ctx* ctx) {
  if (!ctx) return NM_EINVAL;
  if (ctx->kind == 0) return fail(ctx, NM_ESTATE, "nm_launch: no frame set up");
  if (int rc = set_device(ctx)) return rc;
  if (int rc = reset_frame_counters(ctx)) return rc;
  ctx->finished = false;
  int rc = ctx->kind == 1 ? launch_hw(ctx) : (ctx->W > 0 ? launch_deep(ctx) : NM_OK);
  if (rc == NM_OK && ctx->kind == 2 && ctx->W == 0) {
    cudaEventRecord(ctx->ev[0], ctx->stream); cudaEventRecord(ctx->ev[1], ctx->stream); cudaEventRecord(ctx->ev[2], ctx->stream);
  }
  if (rc == NM_OK) ctx->launched = true;
  return rc;
}

int64_t nm_frame_ambiguous(nm_ctx* ctx, int32_t* pix, int64_t cap) {
  if (!ctx) return NM_EINVAL;
  if (int rc = finish_frame(ctx)) return rc;
  int64_t n = (int64_t)ctx->h_ctr[CTR_AMBIG];
  if (pix && n) {
    int64_t m = n < cap ? n : cap;
    if (int rc = read_back(ctx, pix, ctx->ambig.p, (size_t)m * sizeof(int32_t))) return rc;
  }
  return n;
}

int64_t nm_frame_requeue(nm_ctx* ctx, int32_t* pix, int32_t* at_iter, int64_t cap) {
  if (!ctx) return NM_EINVAL;
  if (int rc = finish_frame(ctx)) return rc;
  int64_t n = (int64_t)ctx->h_ctr[CTR_REQUEUE];
  int64_t m = n < cap ? n : cap;
  if (pix && m) if (int rc = read_back(ctx, pix, ctx->rq_pix.p, (size_t)m * sizeof(int32_t))) return rc;
  if (at_iter && m) if (int rc = read_back(ctx, at_iter, ctx->rq_iter.p, (size_t)m * sizeof(int32_t))) return rc;
  return n;
}

int nm_poke(nm_ctx* ctx, int64_t pix, nm_escape v) {
  if (!ctx) return NM_EINVAL;
  if (pix < 0 || pix >= ctx->pixels) return fail(ctx, NM_EINVAL, "nm_poke: pixel out of range");
  k_poke<<<1, 1, 0, ctx->stream>>>(ctx->out.as<nm_escape>(), pix, v);
  NM_CUDA(ctx, cudaGetLastError());
  return NM_OK;
}

int nm_read_rows(nm_ctx* ctx, int r0, int r1, nm_escape* dst) {
  if (!ctx) return NM_EINVAL;
  if (r0 < 0 || r1 > ctx->nr || r0 > r1 || !dst) return fail(ctx, NM_EINVAL, "nm_read_rows: bad range");
  if (int rc = finish_frame(ctx)) return rc;
  NM_CUDA(ctx, cudaMemcpyAsync(dst, ctx->out.as<nm_escape>() + (size_t)r0 * ctx->nc,
                               (size_t)(r1 - r0) * ctx->nc * sizeof(nm_escape), cudaMemcpyDefault, ctx->stream));
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NM_OK;
}

int nm_raster_keep(nm_ctx* ctx) {
  if (!ctx) return NM_EINVAL;
  if (int rc = finish_frame(ctx)) return rc;
  NM_CUDA(ctx, ctx->kept.ensure((size_t)ctx->pixels * sizeof(nm_escape)));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->kept.p, ctx->out.p, (size_t)ctx->pixels * sizeof(nm_escape), cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->kept_pixels = ctx->pixels;
  return NM_OK;
}

int64_t nm_raster_diff(nm_ctx* ctx, int32_t* pix, int64_t cap) {
  if (!ctx) return NM_EINVAL;
  if (int rc = finish_frame(ctx)) return rc;
  if (ctx->kept_pixels != ctx->pixels || ctx->pixels == 0) return fail(ctx, NM_ESTATE, "nm_raster_diff: no raster of this size was kept");
  unsigned long long* ctr = ctx->ctr.as<unsigned long long>();
  NM_CUDA(ctx, ctx->difflist.ensure((size_t)ctx->pixels * sizeof(int32_t)));
  NM_CUDA(ctx, cudaMemsetAsync(&ctr[CTR_Q_CUR], 0, sizeof(unsigned long long), ctx->stream));
  long long blocks = (ctx->pixels + 255) / 256;
  if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
  k_raster_diff<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ctx->kept.as<nm_escape>(), ctx->out.as<nm_escape>(), ctx->pixels,
                                                           ctx->difflist.as<int32_t>(), &ctr[CTR_Q_CUR]);
  NM_CUDA(ctx, cudaGetLastError());
  unsigned long long n = 0;
  if (int rc = read_back(ctx, &n, &ctr[CTR_Q_CUR], sizeof n)) return rc;
  if (pix && n) {
    const unsigned long long m = n < (unsigned long long)cap ? n : (unsigned long long)cap;
    if (int rc = read_back(ctx, pix, ctx->difflist.p, (size_t)m * sizeof(int32_t))) return rc;
  }
  return (int64_t)n;
}

int nm_raster_restore(nm_ctx* ctx) {
  if (!ctx) return NM_EINVAL;
  if (ctx->kept_pixels != ctx->pixels || ctx->pixels == 0) return fail(ctx, NM_ESTATE, "nm_raster_restore: no raster of this size was kept");
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->out.p, ctx->kept.p, (size_t)ctx->pixels * sizeof(nm_escape), cudaMemcpyDeviceToDevice, ctx->stream));
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NM_OK;
}

int nm_read_rows_pitched(nm_ctx* ctx, int r0, int r1, nm_escape* dst, size_t dst_pitch_bytes) {
  if (!ctx) return NM_EINVAL;
  const size_t row_bytes = (size_t)ctx->nc * sizeof(nm_escape);
  if (r0 < 0 || r1 > ctx->nr || r0 > r1 || !dst || dst_pitch_bytes < row_bytes)
    return fail(ctx, NM_EINVAL, "nm_read_rows_pitched: bad range or pitch");
  if (int rc = finish_frame(ctx)) return rc;
  if (r1 > r0)
    NM_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch_bytes, ctx->out.as<nm_escape>() + (size_t)r0 * ctx->nc, row_bytes, row_bytes,
                                   (size_t)(r1 - r0), cudaMemcpyDefault, ctx->stream));
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NM_OK;
}

int nm_read_rows_pitched_async(nm_ctx* ctx, int r0, int r1, nm_escape* dst, size_t dst_pitch_bytes) {
  if (!ctx) return NM_EINVAL;
  const size_t row_bytes = (size_t)ctx->nc * sizeof(nm_escape);
  if (r0 < 0 || r1 > ctx->nr || r0 > r1 || !dst || dst_pitch_bytes < row_bytes)
    return fail(ctx, NM_EINVAL, "nm_read_rows_pitched_async: bad range or pitch");
  if (int rc = finish_frame(ctx)) return rc;
  if (r1 == r0) return NM_OK;
  if (int rc = issue_deferred_copy(ctx)) return rc;   // an earlier request that no launch has picked up
  const size_t bytes = (size_t)(r1 - r0) * row_bytes;
  NM_CUDA(ctx, ctx->snap.ensure(bytes));
  // the snapshot may only be overwritten once the previous copy-out has drained
  if (ctx->copy_pending) NM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->snap.p, ctx->out.as<nm_escape>() + (size_t)r0 * ctx->nc, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  NM_CUDA(ctx, cudaEventRecord(ctx->ev_snap, ctx->stream));
  ctx->deferred.dst = dst; ctx->deferred.pitch = dst_pitch_bytes; ctx->deferred.row_bytes = row_bytes;
  ctx->deferred.rows = (size_t)(r1 - r0); ctx->deferred.active = true;
  return NM_OK;
}

int nm_read_wait(nm_ctx* ctx) {
  if (!ctx) return NM_EINVAL;
  if (int rc = set_device(ctx)) return rc;
  if (int rc = issue_deferred_copy(ctx)) return rc;
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->side));
  ctx->copy_pending = false;
  return NM_OK;
}

int nm_host_register(nm_ctx* ctx, void* ptr, size_t bytes) {
  if (!ctx || !ptr || !bytes) return NM_EINVAL;
  if (int rc = set_device(ctx)) return rc;
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();  // not sticky: clear it so that the caller can fall back to pageable / NCCL paths
    return fail(ctx, NM_ENOMEM, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e));
  }
  return NM_OK;
}

int nm_host_unregister(nm_ctx* ctx, void* ptr) {
  if (!ctx || !ptr) return NM_EINVAL;
  if (int rc = set_device(ctx)) return rc;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, NM_ECUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
  return NM_OK;
}

int nm_read_pixels(nm_ctx* ctx, const int32_t* pix, int64_t n, nm_escape* dst) {
  if (!ctx) return NM_EINVAL;
  if (n < 0 || (n && (!pix || !dst))) return fail(ctx, NM_EINVAL, "nm_read_pixels: bad arguments");
  if (int rc = finish_frame(ctx)) return rc;
  if (n == 0) return NM_OK;
  NM_CUDA(ctx, ctx->list.ensure((size_t)n * sizeof(int32_t)));
  NM_CUDA(ctx, ctx->gridtmp.ensure((size_t)n * sizeof(nm_escape)));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->list.p, pix, (size_t)n * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
  k_gather<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->out.as<nm_escape>(), ctx->list.as<int32_t>(),
                                                                 ctx->gridtmp.as<nm_escape>(), (long long)n);
  NM_CUDA(ctx, cudaGetLastError());
  NM_CUDA(ctx, cudaMemcpyAsync(dst, ctx->gridtmp.p, (size_t)n * sizeof(nm_escape), cudaMemcpyDefault, ctx->stream));
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NM_OK;
}

int nm_frame_stats(nm_ctx* ctx, nm_stats* out) {
  if (!ctx || !out) return NM_EINVAL;
  if (ctx->launched) { if (int rc = finish_frame(ctx)) return rc; }
  *out = ctx->stats;
  return NM_OK;
}

int nm_render_hw(nm_ctx* ctx, const double* c_re, int nc, const double* c_im, int nr, int N, nm_escape* out) {
  if (int rc = nm_frame_hw(ctx, c_re, nc, c_im, nr, N)) return rc;
  if (int rc = nm_launch(ctx)) return rc;
  return nm_read_rows(ctx, 0, nr, out);
}

int nm_render_deep(nm_ctx* ctx, const nm_deep_tables* t, const double* eps_re, int nc, const double* eps_im, int nr,
                   int cardioid_mode, const uint8_t* cardioid_mask, const int32_t* pix_list, int64_t n_list, int mode,
                   nm_escape* out) {
  if (int rc = nm_frame_deep(ctx, t, eps_re, nc, eps_im, nr, cardioid_mode, cardioid_mask, pix_list, n_list, mode)) return rc;
  if (int rc = nm_launch(ctx)) return rc;
  return nm_read_rows(ctx, 0, nr, out);
}

int nm_resolve(nm_ctx* ctx, const uint8_t* pal_rgb, int n_pal, int N, int sc, int smooth, uint8_t* rgb_out) {
  if (!ctx) return NM_EINVAL;
  if (int rc = finish_frame(ctx)) return rc;
  return run_resolve(ctx, ctx->out.as<nm_escape>(), ctx->nr, ctx->nc, pal_rgb, n_pal, N, sc, smooth, rgb_out);
}

int nm_resolve_grid(nm_ctx* ctx, const nm_escape* grid, int nr, int nc, const uint8_t* pal_rgb, int n_pal, int N, int sc,
                    int smooth, uint8_t* rgb_out) {
  if (!ctx || !grid) return NM_EINVAL;
  if (int rc = set_device(ctx)) return rc;
  size_t bytes = (size_t)nr * nc * sizeof(nm_escape);
  NM_CUDA(ctx, ctx->gridtmp.ensure(bytes));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->gridtmp.p, grid, bytes, cudaMemcpyDefault, ctx->stream));
  int saved = ctx->kind;
  ctx->kind = -1;
  int rc = run_resolve(ctx, ctx->gridtmp.as<nm_escape>(), nr, nc, pal_rgb, n_pal, N, sc, smooth, rgb_out);
  ctx->kind = saved;
  return rc;
}

int nm_video_inbetween(nm_ctx* ctx, const uint8_t* prev_rgb, const uint8_t* next_rgb, int H, int W, int nr, int nc, int rate,
                       uint8_t* frames_out) {
  if (!ctx) return NM_EINVAL;
  if (!prev_rgb || !next_rgb || !frames_out || H < 1 || W < 1 || nr < 1 || nc < 1 || rate < 1)
    return fail(ctx, NM_EINVAL, "nm_video_inbetween: bad arguments");
  if (int rc = set_device(ctx)) return rc;
  const size_t kb = (size_t)H * W * 3, ob = (size_t)rate * nr * nc * 3;
  NM_CUDA(ctx, ctx->vprev.ensure(kb));
  NM_CUDA(ctx, ctx->vnext.ensure(kb));
  NM_CUDA(ctx, ctx->vout.ensure(ob));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->vprev.p, prev_rgb, kb, cudaMemcpyDefault, ctx->stream));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->vnext.p, next_rgb, kb, cudaMemcpyDefault, ctx->stream));
  K5Params p;
  p.prev = ctx->vprev.as<uint8_t>(); p.next = ctx->vnext.as<uint8_t>();
  p.H = H; p.W = W; p.nr = nr; p.nc = nc; p.rate = rate; p.out = ctx->vout.as<uint8_t>();
  p.v = (float)pow(1.5, 1.0 / rate);
  long long total = (long long)rate * nr * nc;
  long long blocks = (total + 255) / 256;
  const long long maxb = (long long)ctx->sm_count * 16;
  if (blocks > maxb) blocks = maxb;
  NM_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  k5_inbetween<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p);
  NM_CUDA(ctx, cudaGetLastError());
  cudaEvent_t evend;
  NM_CUDA(ctx, cudaEventCreate(&evend));
  NM_CUDA(ctx, cudaEventRecord(evend, ctx->stream));
  ctx->stats.kernel_launches++;
  NM_CUDA(ctx, cudaMemcpyAsync(frames_out, ctx->vout.p, ob, cudaMemcpyDefault, ctx->stream));
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[3], evend);
  ctx->stats.ms_k4 = ms;
  cudaEventDestroy(evend);
  return NM_OK;
}

int nm_palette_cache(nm_ctx* ctx, int n_cycles, const int* hue_counts, const float* hue_values, const int* hue_periods,
                     int hue_period, int n_sat, const float* sat_values, int sat_period, int n_lum, const float* lum_amp,
                     const int* lum_period, int N, uint8_t* rgb_out) {
  if (!ctx) return NM_EINVAL;
  if (n_cycles < 1 || !hue_counts || !hue_values || !hue_periods || hue_period < 1 || n_sat < 1 || !sat_values ||
      sat_period < 1 || n_lum < 0 || (n_lum && (!lum_amp || !lum_period)) || N < 0)
    return fail(ctx, NM_EINVAL, "nm_palette_cache: bad arguments");
  if (int rc = set_device(ctx)) return rc;
  int n_hue = 0;
  std::vector<int> offs((size_t)n_cycles);
  for (int k = 0; k < n_cycles; k++) {
    if (hue_counts[k] < 1 || hue_periods[k] < 1) return fail(ctx, NM_EINVAL, "nm_palette_cache: empty hue cycle");
    offs[k] = n_hue;
    n_hue += hue_counts[k];
  }
  for (int k = 0; k < n_lum; k++) if (lum_period[k] < 1) return fail(ctx, NM_EINVAL, "nm_palette_cache: bad wave period");
  // one packed parameter block: [counts | offsets | periods | lum_period] ints, then [hue | sat | lum_amp] floats
  const size_t n_int = (size_t)3 * n_cycles + n_lum, n_flt = (size_t)n_hue + n_sat + n_lum;
  std::vector<int32_t> pack(n_int + n_flt);
  memcpy(&pack[0], hue_counts, sizeof(int) * n_cycles);
  memcpy(&pack[n_cycles], offs.data(), sizeof(int) * n_cycles);
  memcpy(&pack[2 * n_cycles], hue_periods, sizeof(int) * n_cycles);
  if (n_lum) memcpy(&pack[3 * n_cycles], lum_period, sizeof(int) * n_lum);
  float* pf = reinterpret_cast<float*>(&pack[n_int]);
  memcpy(pf, hue_values, sizeof(float) * n_hue);
  memcpy(pf + n_hue, sat_values, sizeof(float) * n_sat);
  if (n_lum) memcpy(pf + n_hue + n_sat, lum_amp, sizeof(float) * n_lum);
  NM_CUDA(ctx, ctx->palpar.ensure(pack.size() * 4));
  NM_CUDA(ctx, ctx->paldev.ensure((size_t)3 * (N > 0 ? N : 1)));
  NM_CUDA(ctx, cudaMemcpyAsync(ctx->palpar.p, pack.data(), pack.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `pack` is pageable and dies with this scope
  K6Params p;
  const int* di = ctx->palpar.as<int>();
  const float* df = reinterpret_cast<const float*>(di + n_int);
  p.n_cycles = n_cycles; p.hue_counts = di; p.hue_offsets = di + n_cycles; p.hue_periods = di + 2 * n_cycles;
  p.lum_period = di + 3 * n_cycles; p.hue_values = df; p.sat_values = df + n_hue; p.lum_amp = df + n_hue + n_sat;
  p.hue_period = hue_period; p.n_sat = n_sat; p.sat_period = sat_period; p.n_lum = n_lum; p.N = N;
  p.rgb = ctx->paldev.as<uint8_t>();
  if (N > 0) {
    int blocks = (N + 255) / 256;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k6_palette<<<blocks, 256, 0, ctx->stream>>>(p);
    NM_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
  }
  ctx->paldev_n = N;
  if (rgb_out && N > 0) {
    NM_CUDA(ctx, cudaMemcpyAsync(rgb_out, ctx->paldev.p, (size_t)3 * N, cudaMemcpyDefault, ctx->stream));
    NM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return NM_OK;
}

int nm_resolve_device_palette(nm_ctx* ctx, int N, int sc, int smooth, uint8_t* rgb_out) {
  if (!ctx) return NM_EINVAL;
  if (ctx->paldev_n < 1) return fail(ctx, NM_ESTATE, "nm_resolve_device_palette: no palette generated (nm_palette_cache)");
  if (int rc = finish_frame(ctx)) return rc;
  return run_resolve(ctx, ctx->out.as<nm_escape>(), ctx->nr, ctx->nc, ctx->paldev.as<uint8_t>(), ctx->paldev_n, N, sc, smooth,
                     rgb_out);
}

void nm_k3_filter_entry(double zr, double zi, double gb, uint32_t entry[5]) {
  K3Filt f; int32_t e;
  k3_filter_entry(zr, zi, gb, &f, &e);
  entry[0] = f.lo_r; entry[1] = f.w_r; entry[2] = f.lo_i; entry[3] = f.w_i; entry[4] = (uint32_t)e;
}

int32_t nm_k3_seg_bound(const double* z, const double* gb, int j0, int jmax, double e_max) {
  if (!z || !gb || j0 < 0) return 0;
  return k3_seg_bound(z, z + 1, gb, 2, j0, jmax, e_max);
}


int nm_k3_filter_fires(const uint32_t entry[5], double dr, double di, int scaled, int* glitch, int* escape) {
  K3Filt f; f.lo_r = entry[0]; f.w_r = entry[1]; f.lo_i = entry[2]; f.w_i = entry[3];
  const uint32_t m = scaled ? 0xffffffffu : 0u;
  const bool g = k3_filter_glitch(f, dr, di, m), e = k3_filter_escape((int32_t)entry[4], dr, di, m);
  if (glitch) *glitch = g;
  if (escape) *escape = e;
  return (g || e) ? 1 : 0;
}

int nm_fp64_peak(nm_ctx* ctx, int kind, int iters, double* inst_per_s, double* ms_out) {
  if (!ctx || kind < 0 || kind > 15 || iters < 1) return NM_EINVAL;
  if (int rc = set_device(ctx)) return rc;
  NM_CUDA(ctx, ctx->fixapply.ensure(64));
  const int e = nm_probe_run((void*)ctx->stream, ctx->sm_count, kind, iters, ctx->fixapply.as<double>(), inst_per_s, ms_out);   // fp64_probe.cu
  if (e != 0) return fail(ctx, NM_ECUDA, "nm_fp64_peak: %s", cudaGetErrorString((cudaError_t)e));
  return NM_OK;
}

int nm_device_info(nm_ctx* ctx, int* sm_count, int* sm_clock_khz, size_t* hbm_bytes, char* name, int cap) {
  if (!ctx) return NM_EINVAL;
  cudaDeviceProp prop;
  NM_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (sm_clock_khz) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device); *sm_clock_khz = khz; }
  if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
  if (name && cap > 0) { strncpy(name, prop.name, cap - 1); name[cap - 1] = 0; }
  return NM_OK;
}

}  // extern "C"
