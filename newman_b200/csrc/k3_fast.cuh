// K3 (fast path, NM_MODE_REQUEUE) — same arithmetic as k3_perturb.cuh (see there for the operation
// order, the table layout and the level/chunk scheduling), re-organised around what ncu showed
// (profiles/): the simple kernel kept the FP64 pipe only 44 % busy because the shared-memory pipe
// was 93 % busy — a warp-wide 16-byte load returns 512 B through a 128 B/clk path (4 wavefronts
// even when every lane reads the same address), as many SM cycles per iteration as the FP64 work.
//
// k3_fast<P>: every lane carries P pixels that sit at the SAME orbit index j, so one Z[j+1] /
// glitch-bound load feeds P delta updates (shared-memory wavefronts per pixel-iteration drop ~4P x)
// and the P independent dependency chains give the FP64 pipe ILP. All states handled here have
// j == 0 (mod 4) and advance in branch-free blocks of 4 iterations: per iteration and pixel 10 FP64
// instructions + one integer compare that ORs "high word of |z|^2 <= high word of the glitch bound"
// into a per-pixel flag; the escape test is made once per block on the last |z|^2 (once |z| > 1024
// it grows monotonically and cannot overflow within 3 more steps).
//
// Nothing is decided inside this kernel. A pixel whose block was flagged is rolled back to the
// block's start state and EXPORTED (state, index) while its lane-mates simply keep their end-of-
// block state; so is a pixel that cannot take another whole block (iteration limit or end of the
// orbit table less than 4 steps away). Exported pixels are parked on the reference orbit itself
// (delta = eps = 0, which stays 0 and never flags) until the lane's pass through the chunk ends, and
// are appended to the event queue with the same warp-aggregated reservation as the survivors that
// move on to the next chunk. k3_events then replays at most 4 checked steps per exported pixel with
// the exact double comparisons — escape (+ smoothing), glitch (-> re-queue list), iteration limit,
// rebase at the end of the orbit, or "false alarm" (-> carried into the next sweep, index again a
// multiple of 4). Every decision is therefore the one k3_perturb.cuh and the oracle take.
//
// (Two earlier versions replayed flagged blocks inside the warp; on the level where half of the
// pixels escape that cost 3x, later 1.6x, the time of a full level, and latency-bound tail levels
// ran 5x slower than the simple kernel — profiles/r01b_*.)
#pragma once
#include "k3_checked.cuh"
#include "k3_perturb.cuh"

namespace nm {

constexpr int K3F_THREADS = 256;
constexpr int K3_EVENT_BUDGET = 252;  // extra checked steps k3_events grants a freshly rebased state (multiple of 4)

template <bool SCALED>
__global__ void __launch_bounds__(256) k3_events(CheckedParams p, EpsTab eps_tab,
                                                 const PixState* events, const unsigned long long* count,
                                                 FreshArrays carry, unsigned long long* carry_count, unsigned* hist) {
  const unsigned long long n = *count;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long executed = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    PixState e = events[i];
    EpsVal<SCALED> eps;
    eps.load(eps_tab, e.pix);
    int steps = 0;
    bool cont = advance_checked<SCALED>(p, e.pix, eps, e.off, e.dr, e.di, e.e, e.j, 4, &steps);
    // A state that was just rebased onto the start of the orbit (it outlived the reference: |z| is
    // large) nearly always escapes within a few steps: finish it here rather than carrying it through
    // another sweep of (mostly empty) level launches. Same steps, same decisions, same order.
    if (cont && e.j == 0) {
      int more = 0;
      cont = advance_checked<SCALED>(p, e.pix, eps, e.off, e.dr, e.di, e.e, e.j, K3_EVENT_BUDGET, &more);
      steps += more;
    }
    if (cont) {
      // carried into the next sweep (index is a multiple of 4 again: 0 after a rebase, block start + 4 else)
      unsigned long long slot = atomicAdd(carry_count, 1ULL);
      carry.d[slot] = make_double2(e.dr, e.di);
      carry.j[slot] = e.j;
      carry.off[slot] = e.off;
      carry.pix[slot] = e.pix;
      if (SCALED) carry.e[slot] = e.e;
      atomicAdd(&hist[e.j], 1u);
      atomicMin(&p.ctr[CTR_MINJ], (unsigned long long)e.j);
    }
    executed += (unsigned long long)steps;
  }
  for (int o = 16; o; o >>= 1) executed += __shfl_xor_sync(FULL_MASK, executed, o);
  if ((threadIdx.x & 31) == 0 && executed) {
    atomicAdd(&p.ctr[CTR_EXECUTED], executed);
    atomicAdd(&p.ctr[CTR_CHECKED], executed);
  }
}

// SCALED: states carry a scale exponent (floatexp.cuh). A lane's slots share j, so "re-normalise
// before the step from j = 0 (mod 64)" is one test per lane and block; the per-slot scale S = 2^e and
// eps / 2^e live in registers between re-normalisations.
template <int P, bool SCALED>
__device__ __forceinline__ void k3_fast_body(const K3Params& p, PixState* events) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CH = p.CH;
  const int jbase = p.k * CH;
  int nload = p.Jmax + 1 - jbase;
  if (nload > CH + 1) nload = CH + 1;
  const int nload2 = (nload + 1) & ~1;  // bulk copies move multiples of 16 bytes
  double2* sZ = (double2*)smem_raw;
  const int32_t* sG = (const int32_t*)(smem_raw + (size_t)(CH + 4) * sizeof(double2));  // gb[] doubles; high word at [2i+1]
  // 2*Z[j], formed once per CTA: w = 2Z + delta is then a DADD of two table values instead of fma(2, Z, delta)
  // (bit-identical: 2Z is exact) — DADD issues ~9 % faster than DFMA on this part (profiles/r01_fp64_peak.json)
  double2* sZ2 = (double2*)(smem_raw + (size_t)(CH + 4) * (sizeof(double2) + sizeof(double)));
  __shared__ __align__(8) uint64_t bar;

  const unsigned long long n_cur = p.cur_count ? *p.cur_count : 0ULL;
  unsigned long long n_fresh = 0;
  unsigned fresh_begin = 0;
  {
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
    uint32_t bg = (uint32_t)nload2 * (uint32_t)sizeof(double);
    mbar_expect_tx(&bar, bz + bg);
    bulk_g2s(sZ, p.Z + jbase, bz, &bar);
    bulk_g2s((void*)sG, p.gb + jbase, bg, &bar);
  }

  const int lane = threadIdx.x & 31;
  const int jend = jbase + CH;
  const int jcap = jend < p.Jmax ? jend : p.Jmax;  // a pass can never step beyond this index

  // slot state: 0 empty/parked, 1 live, 2 exported (ev_* hold the state to hand to k3_events)
  double dr[P], di[P], er[P], ei[P], ev_dr[P], ev_di[P], S[P];
  int pix[P], off[P], st[P], ev_j[P], sc[P];
  bool drained = false;
  int j = 0;
  unsigned long long executed = 0;
#pragma unroll
  for (int s = 0; s < P; ++s) {
    dr[s] = di[s] = er[s] = ei[s] = ev_dr[s] = ev_di[s] = 0.0; S[s] = 1.0;
    pix[s] = -1; off[s] = -1; st[s] = 0; ev_j[s] = 0; sc[s] = 0;
  }

  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < nload; i += blockDim.x) { const double2 z = sZ[i]; sZ2[i] = make_double2(2.0 * z.x, 2.0 * z.y); }
  __syncthreads();

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
            dr[s] = di[s] = er[s] = ei[s] = 0.0; pix[s] = -1; off[s] = -1; st[s] = 0; sc[s] = 0; S[s] = 1.0;
            if (g < g_cur) {
              unsigned long long idx = g * P + s;
              if (idx < n_cur) {
                PixState q = p.cur[idx];
                dr[s] = q.dr; di[s] = q.di; pix[s] = q.pix; off[s] = q.off; j = q.j; sc[s] = SCALED ? q.e : 0;
              }
            } else {
              int w = p.fresh_ids[fresh_begin + (unsigned)((g - g_cur) * P + s)];
              if (w >= 0) {
                double2 d0 = p.fresh.d[w];
                dr[s] = d0.x; di[s] = d0.y; off[s] = p.fresh.off[w]; j = p.fresh.j[w]; pix[s] = p.fresh.pix[w];
                if (SCALED) sc[s] = p.fresh.e[w];
              }
            }
            if (pix[s] >= 0) {
              EpsVal<SCALED> eps;
              eps.load(p.eps, pix[s]);
              er[s] = eps.re_at(sc[s]);
              ei[s] = eps.im_at(sc[s]);
              if (SCALED) S[s] = pow2d(sc[s]);
              st[s] = 1;
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
      // Export the live slots that cannot take another whole block before the chunk end / table end /
      // their iteration limit — unless they simply reached the chunk end with room beyond it: those
      // move on to the next level after the loop. nb = whole blocks every remaining live slot can take.
      int nb = 0x7fffffff;
      bool any_live = false;
#pragma unroll
      for (int s = 0; s < P; ++s)
        if (st[s] == 1) {
          const int jN = p.N - 1 - off[s];
          const int lim = jcap < jN ? jcap : jN;
          const int room = (lim - j) >> 2;
          if (room > 0) { any_live = true; if (room < nb) nb = room; }
          else if (!(j == jend && j < p.Jmax && jN > j)) {
            st[s] = 2; ev_dr[s] = dr[s]; ev_di[s] = di[s]; ev_j[s] = j;
            executed += (unsigned long long)(j - j_in);
            dr[s] = di[s] = er[s] = ei[s] = 0.0;
          }
        }
      if (!(lane_active && any_live)) nb = 0;
      if (!__any_sync(FULL_MASK, nb > 0)) break;

      // nb branch-free blocks; a flagged slot is rolled back, exported and parked on the spot, its
      // lane-mates keep going (their limits can only be farther away, so nb stays valid)
      for (int b = 0; __any_sync(FULL_MASK, b < nb); ++b) {
        if (b < nb) {
          if (SCALED && (j & RENORM_MASK) == 0) {
#pragma unroll
            for (int s = 0; s < P; ++s)
              if (st[s] == 1) {
                pstate ps; ps.dr = dr[s]; ps.di = di[s]; ps.e = sc[s];
                state_renorm(ps);
                if (ps.e != sc[s]) {
                  EpsVal<SCALED> eps;
                  eps.load(p.eps, pix[s]);
                  er[s] = eps.re_at(ps.e); ei[s] = eps.im_at(ps.e);
                  S[s] = pow2d(ps.e);
                }
                dr[s] = ps.dr; di[s] = ps.di; sc[s] = ps.e;
              }
          }
          double dr0[P], di0[P];
#pragma unroll
          for (int s = 0; s < P; ++s) { dr0[s] = dr[s]; di0[s] = di[s]; }
          double2 x2 = sZ2[j - jbase];
          bool bad[P];
          int hi_last[P];
#pragma unroll
          for (int s = 0; s < P; ++s) bad[s] = false;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int jl = j + t + 1 - jbase;
            const double2 y = sZ[jl];
            const double2 y2 = sZ2[jl];
            const int g = sG[2 * jl + 1];
#pragma unroll
            for (int s = 0; s < P; ++s) {
              double wr, wi;
              if (SCALED) { wr = __fma_rn(S[s], dr[s], x2.x); wi = __fma_rn(S[s], di[s], x2.y); }
              else { wr = x2.x + dr[s]; wi = x2.y + di[s]; }   // == fma(2, Z, delta): 2Z is exact
              double ndr = __fma_rn(-di[s], wi, __fma_rn(dr[s], wr, er[s]));
              double ndi = __fma_rn(di[s], wr, __fma_rn(dr[s], wi, ei[s]));
              dr[s] = ndr; di[s] = ndi;
              double zr, zi;
              if (SCALED) { zr = __fma_rn(S[s], ndr, y.x); zi = __fma_rn(S[s], ndi, y.y); }
              else { zr = y.x + ndr; zi = y.y + ndi; }
              double zmag = __fma_rn(zi, zi, zr * zr);
              int hi = __double2hiint(zmag);
              bad[s] = bad[s] || (hi <= g);
              hi_last[s] = hi;
            }
            x2 = y2;
          }
          bool any_bad = false;
#pragma unroll
          for (int s = 0; s < P; ++s) { bad[s] = bad[s] || (hi_last[s] >= ESC_HI); any_bad = any_bad || bad[s]; }
          if (any_bad) {
#pragma unroll
            for (int s = 0; s < P; ++s)
              if (st[s] == 1 && bad[s]) {
                st[s] = 2; ev_dr[s] = dr0[s]; ev_di[s] = di0[s]; ev_j[s] = j;
                executed += (unsigned long long)(j - j_in);
                dr[s] = di[s] = er[s] = ei[s] = 0.0;
              }
          }
          j += 4;
        }
      }
    }

    // ---- hand over (converged; warp-aggregated appends) --------------------------------------------
#pragma unroll
    for (int s = 0; s < P; ++s) {
      const bool toNext = lane_active && st[s] == 1;   // reached the chunk end alive
      const bool toEvents = lane_active && st[s] == 2;
      if (toNext) executed += (unsigned long long)(j - j_in);
      unsigned long long slot = warp_reserve(p.next_count, toNext);
      if (toNext) {
        PixState q; q.dr = dr[s]; q.di = di[s]; q.pix = pix[s]; q.j = j; q.off = off[s]; q.e = SCALED ? sc[s] : 0;
        p.next[slot] = q;
      }
      slot = warp_reserve(&p.ctr[CTR_EVENTS], toEvents);
      if (toEvents) {
        PixState q; q.dr = ev_dr[s]; q.di = ev_di[s]; q.pix = pix[s]; q.j = ev_j[s]; q.off = off[s]; q.e = SCALED ? sc[s] : 0;  // an exported slot is parked: its exponent is never re-normalised
        events[slot] = q;
      }
      st[s] = 0; pix[s] = -1;
    }
  }

  for (int o = 16; o; o >>= 1) executed += __shfl_xor_sync(FULL_MASK, executed, o);
  if (lane == 0 && executed) atomicAdd(&p.ctr[CTR_EXECUTED], executed);
}

// A level with so few states that every one can have a lane of its own is latency bound: a lane's
// pass takes 1024 x (P x 10 FP64 instructions issued from ONE warp, ~7 cycles apart) — 146 ns per
// iteration with P = 4 on the tail levels of cfg2 against 31 ns with one pixel per lane. Such levels
// (uniformly for the whole grid: the counts are launch-wide) run the same body with P = 1.
constexpr unsigned long long K3F_SPARSE_MAX = 148ULL * 2 * K3F_THREADS;

template <int P, bool SCALED>
__global__ void __launch_bounds__(K3F_THREADS, (P == 4 ? 2 : (SCALED ? 2 : 3)))
k3_fast(K3Params p, PixState* events) {
  const int jbase = p.k * p.CH;
  int l1 = jbase + p.CH;
  if (l1 > p.Jmax + 1) l1 = p.Jmax + 1;
  const unsigned long long n_states = (p.cur_count ? *p.cur_count : 0ULL) + (p.fresh_off[l1] - p.fresh_off[jbase]);
  if (P > 1 && n_states <= K3F_SPARSE_MAX) k3_fast_body<1, SCALED>(p, events);
  else k3_fast_body<P, SCALED>(p, events);
}

}  // namespace nm
