// K3 (fast path, NM_MODE_REQUEUE) — same arithmetic as k3_perturb.cuh (see there for the operation
// order, the table layout and the level/chunk scheduling), re-organised around what ncu showed for
// the first version (profiles/r01a_k3_level_v1_metrics.txt): the FP64 pipe was only 44 % busy
// because the shared-memory pipe was 93 % busy — a warp-wide 16-byte load returns 512 B through a
// 128 B/clk path (4 wavefronts even when every lane reads the same address), i.e. as many SM
// cycles per iteration as the FP64 work itself, doubled again by bank conflicts once lanes had
// drifted to different orbit indices.
//
// Here every lane carries P pixels that sit at the SAME orbit index j, so one Z[j+1] / glitch-bound
// load feeds P delta updates (shared-memory wavefronts per pixel-iteration drop ~8x), and the P
// independent dependency chains give the FP64 pipe ILP. Within a lane the P pixels advance in lock
// step from where they were picked up to the end of the chunk. Survivors of the previous chunk all
// start at the chunk boundary; fresh pixels from K2 are grouped by their exact start index L (runs
// padded to a multiple of P by the scatter), so a group always shares j.
//
// Inner loop: blocks of 4 iterations without any branch. Per iteration and pixel: 10 FP64
// instructions + one integer compare that ORs "high word of |z|^2 <= high word of the glitch bound"
// into a flag; the escape test is made once per block on the last |z|^2 (once |z| > 1024 it grows
// monotonically and cannot overflow within 3 more steps). A flagged block is replayed from its
// saved start state one step at a time with the exact double comparisons, so every decision is
// identical to the simple kernel and to the oracle.
//
// The pass is warp-synchronous: the block loop's trip condition is a warp vote, lanes whose block
// was flagged (or that have < 4 steps left) take the checked steps together in one converged
// section, and escapes are only *recorded* there (pixel, iteration, |z|^2) with one warp-aggregated
// reservation per slot; the smoothing logarithms run afterwards in k3_smooth over the dense list.
// (An earlier version let each lane loop on its own: after the first escape the lanes of a warp
// drifted apart and the level where most pixels escape ran 3x slower than a full level.)
#pragma once
#include "k3_perturb.cuh"

namespace nm {

constexpr int K3F_THREADS = 256;

struct __align__(16) EscRec {  // an escaped pixel waiting for its smoothing value
  int32_t pix;
  int32_t it;
  double r2;
};

template <int P>
__global__ void __launch_bounds__(K3F_THREADS) k3_fast(K3Params p, EscRec* esc_list) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CH = p.CH;
  const int jbase = p.k * CH;
  int nload = p.Jmax + 1 - jbase;
  if (nload > CH + 1) nload = CH + 1;
  const int nload4 = (nload + 1) & ~1;  // bulk copies move multiples of 16 bytes
  double2* sZ = (double2*)smem_raw;
  double* sGB = (double*)(smem_raw + (size_t)(CH + 4) * sizeof(double2));  // full glitch bounds gb[j]
  const int32_t* sG = (const int32_t*)sGB;                                 // their high words: sG[2*i + 1]
  __shared__ __align__(8) uint64_t bar;

  const unsigned long long n_cur = p.cur_count ? *p.cur_count : 0ULL;
  unsigned long long n_fresh = 0;
  unsigned fresh_begin = 0;
  if (p.fresh_off) {
    int l1 = jbase + CH;
    if (l1 > p.Jmax + 1) l1 = p.Jmax + 1;
    fresh_begin = p.fresh_off[jbase];
    n_fresh = p.fresh_off[l1] - fresh_begin;  // multiple of the group size by construction
  }
  const unsigned long long g_cur = (n_cur + P - 1) / P;
  const unsigned long long g_total = g_cur + n_fresh / P;
  if (g_total == 0) return;

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t bz = (uint32_t)nload * (uint32_t)sizeof(double2);
    uint32_t bg = (uint32_t)nload4 * (uint32_t)sizeof(double);
    mbar_expect_tx(&bar, bz + bg);
    bulk_g2s(sZ, p.Z + jbase, bz, &bar);
    bulk_g2s(sGB, p.gb + jbase, bg, &bar);
  }

  const int lane = threadIdx.x & 31;
  const int jend = jbase + CH;
  const int jcap = jend < p.Jmax ? jend : p.Jmax;  // a pass can never step beyond this index

  double dr[P], di[P], er[P], ei[P];
  int pix[P], off[P];
  bool drained = false;
  int j = 0;
  unsigned long long executed = 0, rebased = 0, checked = 0;
#pragma unroll
  for (int s = 0; s < P; ++s) { dr[s] = di[s] = er[s] = ei[s] = 0.0; pix[s] = -1; off[s] = -1; }

  mbar_wait(&bar, 0);

  for (;;) {
    // ---- re-deal: one group of P same-index pixels per lane (every lane is idle here) ------------
    bool lane_active = false;
    if (!drained) {
      unsigned long long base = 0;
      if (lane == 0) {
        if (((volatile unsigned long long*)p.ctr)[CTR_CANCEL]) base = g_total;
        else base = atomicAdd(p.head, 32ULL);
      }
      base = __shfl_sync(FULL_MASK, base, 0);
      if (base >= g_total) drained = true;
      else {
        unsigned long long g = base + lane;
        if (g < g_total) {
#pragma unroll
          for (int s = 0; s < P; ++s) {
            dr[s] = di[s] = er[s] = ei[s] = 0.0; pix[s] = -1; off[s] = -1;
            if (g < g_cur) {
              unsigned long long idx = g * P + s;
              if (idx < n_cur) {
                PixState st = p.cur[idx];
                dr[s] = st.dr; di[s] = st.di; pix[s] = st.pix; off[s] = st.off; j = st.j;
              }
            } else {
              int w = p.fresh_ids[fresh_begin + (unsigned)((g - g_cur) * P + s)];
              if (w >= 0) {
                double2 d0 = p.init_d[w];
                dr[s] = d0.x; di[s] = d0.y; off[s] = -1; j = p.init_j[w];
                pix[s] = p.pix_list ? p.pix_list[w] : w;
              }
            }
            if (pix[s] >= 0) {
              int r = pix[s] / p.nc, c = pix[s] - r * p.nc;
              er[s] = p.eps_re[c];
              ei[s] = p.eps_im[r];
              lane_active = true;
            }
          }
        }
      }
    }
    if (!__any_sync(FULL_MASK, lane_active)) break;

    // ---- one warp-synchronous pass through the chunk -----------------------------------------------
    const int j_in = j;  // every slot of this lane's group entered at this index
    for (;;) {
      // this lane's stop index: chunk end / end of the orbit table / nearest iteration limit
      int jstop = jcap;
      bool any_live = false;
#pragma unroll
      for (int s = 0; s < P; ++s)
        if (pix[s] >= 0) {
          any_live = true;
          int jN = p.N - 1 - off[s];
          if (jN < jstop) jstop = jN;
        }
      const bool running = lane_active && any_live && j < jstop;
      if (!__any_sync(FULL_MASK, running)) break;

      // fast blocks: the whole warp advances block by block; as soon as any lane's block is flagged
      // the warp leaves the loop so that lane can take its checked steps while the others wait 4
      // steps ahead of nobody (they are at the same index again afterwards)
      bool flagged = false;
      bool cand[P];  // slots whose checks must be repeated exactly (all of them for a short tail)
#pragma unroll
      for (int s = 0; s < P; ++s) cand[s] = true;
      for (;;) {
        const bool can = running && (jstop - j >= 4);
        if (!__any_sync(FULL_MASK, can)) break;
        if (can) {
          double dr0[P], di0[P];
#pragma unroll
          for (int s = 0; s < P; ++s) { dr0[s] = dr[s]; di0[s] = di[s]; }
          double2 x = sZ[j - jbase];
          bool bad[P];
          int hi_last[P];
#pragma unroll
          for (int s = 0; s < P; ++s) bad[s] = false;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int jl = j + t + 1 - jbase;
            const double2 y = sZ[jl];
            const int g = sG[2 * jl + 1];
#pragma unroll
            for (int s = 0; s < P; ++s) {
              double wr = __fma_rn(2.0, x.x, dr[s]);
              double wi = __fma_rn(2.0, x.y, di[s]);
              double ndr = __fma_rn(-di[s], wi, __fma_rn(dr[s], wr, er[s]));
              double ndi = __fma_rn(di[s], wr, __fma_rn(dr[s], wi, ei[s]));
              dr[s] = ndr; di[s] = ndi;
              double zr = y.x + ndr, zi = y.y + ndi;
              double zmag = __fma_rn(zi, zi, zr * zr);
              int hi = __double2hiint(zmag);
              bad[s] = bad[s] || (hi <= g);
              hi_last[s] = hi;
            }
            x = y;
          }
          bool any_bad = false;
#pragma unroll
          for (int s = 0; s < P; ++s) { bad[s] = bad[s] || (hi_last[s] >= ESC_HI); any_bad = any_bad || bad[s]; }
          if (any_bad) {  // replay this block with exact checks (below, together with the other lanes)
#pragma unroll
            for (int s = 0; s < P; ++s) { dr[s] = dr0[s]; di[s] = di0[s]; cand[s] = bad[s]; }
            flagged = true;
          } else {
            j += 4;
          }
        }
        if (__any_sync(FULL_MASK, flagged)) break;
      }

      // checked steps (converged): a flagged block, or the < 4 iterations left before jstop
      int nslow = (running && (flagged || jstop - j < 4)) ? jstop - j : 0;
      if (nslow > 4) nslow = 4;
      checked += (unsigned long long)nslow;
      bool esc[P], glt[P];
      int ev_it[P];
      double ev_r2[P];
#pragma unroll
      for (int s = 0; s < P; ++s) { esc[s] = glt[s] = false; ev_it[s] = 0; ev_r2[s] = 0.0; }
      if (nslow > 0) {
        double2 x = sZ[j - jbase];
        for (int t = 0; t < nslow; ++t) {
          const double2 y = sZ[j + 1 - jbase];
          ++j;
#pragma unroll
          for (int s = 0; s < P; ++s) {
            double wr = __fma_rn(2.0, x.x, dr[s]);
            double wi = __fma_rn(2.0, x.y, di[s]);
            double ndr = __fma_rn(-di[s], wi, __fma_rn(dr[s], wr, er[s]));
            double ndi = __fma_rn(di[s], wr, __fma_rn(dr[s], wi, ei[s]));
            dr[s] = ndr; di[s] = ndi;
            if (cand[s] && pix[s] >= 0 && !esc[s] && !glt[s]) {
              double zr = y.x + ndr, zi = y.y + ndi;
              double zmag = __fma_rn(zi, zi, zr * zr);
              if (zmag > BAILOUT2) {
                esc[s] = true;
                ev_it[s] = j + off[s];
                ev_r2[s] = zr * zr + zi * zi;  // sqMag as the reference forms it (complex.h:23)
              } else if (j != p.Jmax && zmag < sGB[j - jbase]) {
                glt[s] = true;
                ev_it[s] = j + off[s];
              }
              if (esc[s] || glt[s]) {  // finished: park the slot on the reference orbit (delta = eps = 0)
                dr[s] = di[s] = er[s] = ei[s] = 0.0;
                executed += (unsigned long long)(j - j_in);
              }
            }
          }
          x = y;
        }
      }
      // record the events of this section: one reservation per slot for the whole warp
      bool any_ev = false;
#pragma unroll
      for (int s = 0; s < P; ++s) any_ev = any_ev || esc[s] || glt[s];
      if (__any_sync(FULL_MASK, any_ev))
#pragma unroll
      for (int s = 0; s < P; ++s) {
        unsigned long long slot = warp_reserve(&p.ctr[CTR_ESCAPED], esc[s]);
        if (esc[s]) {
          EscRec e; e.pix = pix[s]; e.it = ev_it[s]; e.r2 = ev_r2[s];
          esc_list[slot] = e;
          pix[s] = -1;
        }
        slot = warp_reserve(&p.ctr[CTR_REQUEUE], glt[s]);
        if (glt[s]) {
          p.rq_pix[slot] = pix[s];
          p.rq_iter[slot] = ev_it[s];
          p.out[pix[s]].iterations = -1;
          p.out[pix[s]].smoothing = 0.0f;
          pix[s] = -1;
        }
      }
      // iteration limit reached by some pixels of this lane: (N, 0)   (mandelbrot.cpp:226-228)
      if (running && j == jstop) {
#pragma unroll
        for (int s = 0; s < P; ++s)
          if (pix[s] >= 0 && j + off[s] + 1 >= p.N) {
            p.out[pix[s]].iterations = p.N;
            p.out[pix[s]].smoothing = 0.0f;
            pix[s] = -1;
            dr[s] = di[s] = er[s] = ei[s] = 0.0;
            executed += (unsigned long long)(j - j_in);
          }
      }
    }

    // ---- hand the survivors on (converged; warp-aggregated appends) -----------------------------
#pragma unroll
    for (int s = 0; s < P; ++s) {
      bool live = lane_active && pix[s] >= 0;
      if (live) executed += (unsigned long long)(j - j_in);
      bool rebase = live && j == p.Jmax;
      bool toNext = live && !rebase && j == jend;
      if (rebase) {
        // continue from the virtual iterate Z[0] = 0 with delta = z (exact algebra: z' = z^2 + c)
        double2 xj = sZ[j - jbase];
        dr[s] = xj.x + dr[s];
        di[s] = xj.y + di[s];
        rebased++;
      }
      unsigned long long slot = warp_reserve(p.restart_count, rebase);
      if (rebase) {
        PixState st; st.dr = dr[s]; st.di = di[s]; st.pix = pix[s]; st.j = 0; st.off = j + off[s]; st.pad = 0;
        p.restart[slot] = st;
      }
      slot = warp_reserve(p.next_count, toNext);
      if (toNext) {
        PixState st; st.dr = dr[s]; st.di = di[s]; st.pix = pix[s]; st.j = j; st.off = off[s]; st.pad = 0;
        p.next[slot] = st;
      }
      pix[s] = -1;
    }
  }

  for (int o = 16; o; o >>= 1) {
    executed += __shfl_xor_sync(FULL_MASK, executed, o);
    rebased += __shfl_xor_sync(FULL_MASK, rebased, o);
    checked += __shfl_xor_sync(FULL_MASK, checked, o);
  }
  if (lane == 0) {
    if (executed) atomicAdd(&p.ctr[CTR_EXECUTED], executed);
    if (rebased) atomicAdd(&p.ctr[CTR_REBASED], rebased);
    if (checked) atomicAdd(&p.ctr[CTR_CHECKED], checked);
  }
}

// Smoothing for the escapes recorded by k3_fast (mandelbrot.cpp:133-136, 218): dense, converged.
__global__ void __launch_bounds__(256) k3_smooth(const EscRec* list, const unsigned long long* count, nm_escape* out,
                                                 unsigned long long* ctr, FixupRec* fix, unsigned long long fix_cap,
                                                 double log_bailout) {
  const unsigned long long n = *count;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    EscRec e = list[i];
    bool unc;
    float s = smoothing_f32(e.r2, log_bailout, &unc);
    nm_escape v; v.iterations = e.it; v.smoothing = s;
    out[e.pix] = v;
    if (unc) push_fixup(ctr, fix, fix_cap, e.pix, e.r2);
  }
}

}  // namespace nm
