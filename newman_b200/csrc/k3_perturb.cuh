// K3 — FP64 perturbation iteration against the high-precision reference orbit.
// Replaces phase 3 of Mandelbrot::getIterations (reference mandelbrot.cpp:209-224), which continues
// every pixel in full mpf arithmetic, by
//        delta' = delta*(2*Z[j] + delta) + eps,      z' = Z[j+1] + delta'
// in double. The reference has no such loop, so the operation order below is *defined here* and
// restated verbatim by the CPU oracle (oracle/oracle_p.c: oraclep_step); explicit FMAs are used
// (the build is -fmad=false, so only the __fma_rn written here exist):
//        wr = fma(2, xr, dr)            wi = fma(2, xi, di)              (2*x is exact)
//        dr' = fma(-di, wi, fma(dr, wr, er))
//        di' = fma( di, wr, fma(dr, wi, ei))
//        zr = Xr' + dr'   zi = Xi' + di'   |z|^2 = fma(zi, zi, zr*zr)
// = 10 FP64-pipe instructions per executed iteration (7 DFMA, 2 DADD, 1 DMUL; 17 flops). The escape
// and glitch comparisons are done on the integer pipe from the high word of |z|^2 (non-negative
// doubles order like their bit patterns); a hit is a *candidate* that the slow path re-checks with
// the full doubles, so the decisions are exactly `|z|^2 > 2^20` (bailedOut, mandelbrot.cpp:61) and
// `|z|^2 < glitch_tol*|Z[j]|^2`.
//
// Table index j: Z[0] = 0 (virtual iterate before the orbit, used for rebasing), Z[j] = X[j-1].
// A pixel state (j, delta) pairs delta with Z[j]; the reference's iteration index is it = j + off
// (off = -1 until the pixel is rebased).
//
// Scheduling ("levels"): the orbit is consumed in chunks of CH entries. One launch = one chunk:
// the chunk (CH+1 entries of Z plus the glitch-bound high words) is pulled into shared memory by a
// single bulk-async copy (TMA, cp.async.bulk + mbarrier) per CTA, and every warp then runs
// persistently against it: idle lanes are re-dealt pixel states from this level's input queue
// (ballot + one atomicAdd per warp), lanes iterate independently (each lane has its own j, so a
// warp is a bag of pixels, not a tile), and lanes that reach the chunk end append their state to
// the next level's queue. Escaped pixels write their EscapeValue; glitched pixels go to the
// re-queue list for a secondary reference; pixels that outlive the orbit are rebased onto Z[0].
#pragma once
#include "k3_checked.cuh"

namespace nm {



struct K3Params {
  const double2* Z;      // [Jmax+1 (+pad)]
  const int32_t* ghi;    // high words of gb[j] = glitch_tol*|Z[j]|^2
  const double* gb;      // full doubles (slow-path exact check)
  const double2* Z2;     // k3_fast: 2*Z[j] (exact)
  const int4* filt;      // k3_fast: glitch-filter entries (k3_filter.cuh: K3Filt)
  const int32_t* esc_hi; // k3_fast: escape-filter high words
  const int32_t* seg_hi; // k3_fast: per 16-iteration segment, the quiet bound on delta's high words (k3_seg_bound)
  int Jmax;              // last valid table index
  int N, CH, k;
  EpsTab eps;
  int nc;
  const PixState* cur;
  const unsigned long long* cur_count;
  const int32_t* fresh_ids;
  const unsigned* fresh_off;   // [Jmax+2] offsets by exact start index; nullptr after sweep 0
  FreshArrays fresh;           // hand-over records the fresh_ids index
  PixState* next;
  unsigned long long* next_count;
  // k3_fast only (k3_fast.cuh: K3Work): which of the level's K3F_SUBS launches this is, the queue counters of
  // all levels (qcount[k] = survivors handed to level k), two scratch queues and their counters for a split level
  int sub;
  const unsigned long long* qcount;
  PixState* tmp[2];
  unsigned long long* sub_count;
  unsigned long long split_min;   // levels with fewer states are never split
  PixState* restart;
  unsigned long long* restart_count;
  unsigned long long* head;
  nm_escape* out;
  unsigned long long* ctr;
  FixupRec* fix;
  unsigned long long fix_cap;
  int32_t* rq_pix;
  int32_t* rq_iter;
  double log_bailout;
};

constexpr int K3_THREADS = 256;
constexpr int K3_BURST = 256;
constexpr int ESC_HI = 0x41300000;  // high word of 2^20

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// MODE: NM_MODE_REQUEUE (flag glitches) or NM_MODE_REBASE (rebase when |z|^2 < |delta|^2).
// SCALED: states carry a scale exponent (floatexp.cuh); bursts then stop at every index = 0 (mod 64),
// where the state is re-normalised before the next step.
template <int MODE, bool SCALED>
__global__ void __launch_bounds__(K3_THREADS) k3_level(K3Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CH = p.CH;
  const int jbase = p.k * CH;
  int nload = p.Jmax + 1 - jbase;
  if (nload > CH + 1) nload = CH + 1;
  const int nload4 = (nload + 3) & ~3;
  double2* sZ = (double2*)smem_raw;
  int32_t* sG = (int32_t*)(smem_raw + (size_t)(CH + 4) * sizeof(double2));
  __shared__ __align__(8) uint64_t bar;

  const unsigned long long n_cur = p.cur_count ? *p.cur_count : 0ULL;
  unsigned long long n_fresh = 0;
  unsigned fresh_begin = 0;
  if (p.fresh_off) {  // fresh samples are sorted by exact start index; runs are padded with -1
    int l1 = jbase + CH;
    if (l1 > p.Jmax + 1) l1 = p.Jmax + 1;
    fresh_begin = p.fresh_off[jbase];
    n_fresh = p.fresh_off[l1] - fresh_begin;
  }
  const unsigned long long total = n_cur + n_fresh;
  if (total == 0) return;  // uniform across the grid: nothing lives in this chunk

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t bz = (uint32_t)nload * (uint32_t)sizeof(double2);
    uint32_t bg = (uint32_t)nload4 * (uint32_t)sizeof(int32_t);
    mbar_expect_tx(&bar, bz + bg);
    bulk_g2s(sZ, p.Z + jbase, bz, &bar);
    bulk_g2s(sG, p.ghi + jbase, bg, &bar);
  }

  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int jend = jbase + CH;  // states reaching jend move to the next level

  bool active = false, drained = false;
  double dr = 0, di = 0, er = 0, ei = 0, S = 1.0;
  int j = 0, off = -1, pix = 0, e = 0;
  EpsVal<SCALED> eps;
  eps.r0 = eps.i0 = 0.0;
  unsigned long long executed = 0, rebased = 0;

  mbar_wait(&bar, 0);

  for (;;) {
    // ---- re-deal -------------------------------------------------------------------------------
    while (!drained) {
      unsigned idle = __ballot_sync(FULL_MASK, !active);
      if (!idle) break;
      unsigned long long base = 0;
      if (lane == 0) {
        if (((volatile unsigned long long*)p.ctr)[CTR_CANCEL]) base = total;
        else base = atomicAdd(p.head, (unsigned long long)__popc(idle));
      }
      base = __shfl_sync(FULL_MASK, base, 0);
      if (base >= total) { drained = true; break; }
      if (!active) {
        unsigned long long idx = base + __popc(idle & lt_mask);
        if (idx < total) {
          bool got = true;
          if (idx < n_cur) {
            PixState s = p.cur[idx];
            dr = s.dr; di = s.di; pix = s.pix; j = s.j; off = s.off; e = SCALED ? s.e : 0;
          } else {
            int w = p.fresh_ids[fresh_begin + (unsigned)(idx - n_cur)];
            got = w >= 0;
            if (got) {
              double2 d0 = p.fresh.d[w];
              dr = d0.x; di = d0.y;
              j = p.fresh.j[w];
              off = p.fresh.off[w];
              pix = p.fresh.pix[w];
              e = SCALED ? p.fresh.e[w] : 0;
            }
          }
          if (got) {
            eps.load(p.eps, pix);
            er = eps.re_at(e); ei = eps.im_at(e);
            if (SCALED) S = pow2d(e);
            active = true;
          }
        }
      }
      break;  // one reservation per round; leftovers are picked up after the next burst
    }
    if (!__any_sync(FULL_MASK, active)) break;

    // ---- burst ---------------------------------------------------------------------------------
    double zr = 0, zi = 0, zmag = 0, dmag = 0;
    bool stepped = false;
    if (active) {
      // per-lane stop: chunk end, iteration limit (it = j+off = N-1), end of the orbit table
      int jstop = jend;
      int jN = p.N - 1 - off;
      if (jN < jstop) jstop = jN;
      if (p.Jmax < jstop) jstop = p.Jmax;
      int jlim = j + K3_BURST;
      if (jstop < jlim) jlim = jstop;
      if (SCALED) {
        const int jr = (j | RENORM_MASK) + 1;  // next re-normalisation index above j
        if (jr < jlim) jlim = jr;
      }
      if (j < jlim) {
        stepped = true;
        const int j0 = j;
        if (SCALED && (j & RENORM_MASK) == 0) {
          pstate ps; ps.dr = dr; ps.di = di; ps.e = e;
          state_renorm(ps);
          dr = ps.dr; di = ps.di; e = ps.e;
          S = pow2d(e); er = eps.re_at(e); ei = eps.im_at(e);
        }
        double2 x = sZ[j - jbase];
        double xr = x.x, xi = x.y;
        for (;;) {
          const int jl = j + 1 - jbase;
          double2 y = sZ[jl];
          int g = sG[jl];
          double wr, wi;
          if (SCALED) { wr = __fma_rn(S, dr, 2.0 * xr); wi = __fma_rn(S, di, 2.0 * xi); }
          else { wr = __fma_rn(2.0, xr, dr); wi = __fma_rn(2.0, xi, di); }
          double ndr = __fma_rn(-di, wi, __fma_rn(dr, wr, er));
          double ndi = __fma_rn(di, wr, __fma_rn(dr, wi, ei));
          dr = ndr; di = ndi;
          xr = y.x; xi = y.y;
          ++j;
          if (SCALED) { zr = __fma_rn(S, dr, xr); zi = __fma_rn(S, di, xi); }
          else { zr = xr + dr; zi = xi + di; }
          zmag = __fma_rn(zi, zi, zr * zr);
          int hi = __double2hiint(zmag);
          bool cand = (hi >= ESC_HI);
          if (MODE == NM_MODE_REQUEUE) {
            cand = cand || (hi <= g);
          } else {
            if (SCALED) { const double tr = S * dr, ti = S * di; dmag = __fma_rn(ti, ti, tr * tr); }
            else dmag = __fma_rn(di, di, dr * dr);
            cand = cand || (hi <= __double2hiint(dmag));
          }
          if (cand || j == jlim) break;
        }
        executed += (unsigned long long)(j - j0);
      }
    }

    // ---- classify (converged: queue appends are warp-aggregated) ------------------------------
    bool esc = false, glitch = false, rebase = false, atN = false, toNext = false;
    if (active) {
      if (stepped) {
        esc = zmag > BAILOUT2;
        if (!esc) {
          if (MODE == NM_MODE_REQUEUE) glitch = (j != p.Jmax) && (zmag < p.gb[j]);
          else rebase = zmag < dmag;
        }
      }
      if (!esc && !glitch) {
        if (j + off + 1 >= p.N) { atN = true; rebase = false; }
        else if (j == p.Jmax) rebase = true;
        else if (!rebase && j == jend) toNext = true;
      }
    }
    if (esc) {
      double r2 = zr * zr + zi * zi;  // sqMag as the reference forms it (complex.h:23)
      bool unc;
      float s = smoothing_f32(r2, p.log_bailout, &unc);
      p.out[pix].iterations = j + off;
      p.out[pix].smoothing = s;
      if (unc) push_fixup(p.ctr, p.fix, p.fix_cap, pix, r2);
      active = false;
    } else if (atN) {
      p.out[pix].iterations = p.N;
      p.out[pix].smoothing = 0.0f;
      active = false;
    }
    {
      unsigned long long slot = warp_reserve(&p.ctr[CTR_REQUEUE], glitch);
      if (glitch) {
        p.rq_pix[slot] = pix;
        p.rq_iter[slot] = j + off;
        p.out[pix].iterations = -1;
        p.out[pix].smoothing = 0.0f;
        active = false;
      }
    }
    if (rebase) {
      // continue from the virtual iterate Z[0] = 0 with delta = z (exact algebra: z' = z^2 + c)
      rebased++;
      off = j + off;
      j = 0;
      dr = zr; di = zi;
      if (SCALED) { e = 0; S = 1.0; er = eps.r0; ei = eps.i0; }
      if (p.k != 0) {
        // chunk 0 is not resident: park the state for the next sweep
      } else {
        rebase = false;  // chunk 0 is this chunk: keep iterating in place
      }
    }
    {
      unsigned long long slot = warp_reserve(p.restart_count, rebase);
      if (rebase) {
        PixState s; s.dr = dr; s.di = di; s.pix = pix; s.j = 0; s.off = off; s.e = 0;
        p.restart[slot] = s;
        active = false;
      }
    }
    {
      unsigned long long slot = warp_reserve(p.next_count, toNext);
      if (toNext) {
        PixState s; s.dr = dr; s.di = di; s.pix = pix; s.j = j; s.off = off; s.e = e;
        p.next[slot] = s;
        active = false;
      }
    }
  }

  for (int o = 16; o; o >>= 1) {
    executed += __shfl_xor_sync(FULL_MASK, executed, o);
    rebased += __shfl_xor_sync(FULL_MASK, rebased, o);
  }
  if (lane == 0) {
    if (executed) atomicAdd(&p.ctr[CTR_EXECUTED], executed);
    if (rebased) atomicAdd(&p.ctr[CTR_REBASED], rebased);
  }
}

}  // namespace nm
