// K3 (fast path, NM_MODE_REQUEUE) — same arithmetic as k3_perturb.cuh (see there for the operation
// order, the table layout and the level/chunk scheduling), re-organised around what ncu showed
// (profiles/): the simple kernel kept the FP64 pipe only 44 % busy because the shared-memory pipe
// was 93 % busy — a warp-wide 16-byte load returns 512 B through a 128 B/clk path (4 wavefronts
// even when every lane reads the same address), as many SM cycles per iteration as the FP64 work.
//
// k3_fast<P>: every lane carries P pixels that sit at the SAME orbit index j, so one Z[j+1] /
// glitch-bound load feeds P delta updates (shared-memory wavefronts per pixel-iteration drop ~4P x)
// and the P independent dependency chains give the FP64 pipe ILP. All states handled here have
// j == 0 (mod 4) and advance in branch-free blocks of 4 iterations. The kernel runs the delta recurrence
// ALONE — per iteration and pixel 6 FP64 instructions (2 DADD + 4 DFMA) — and never forms z = Z + delta:
// whether the exact glitch / escape comparisons could fire is decided on the integer pipe from the high
// words of delta against a per-index table (k3_filter.cuh: conservative, no false negatives; the glitch
// filter every iteration, the escape filter once per block on the last delta — once |z| > 1024 it grows
// monotonically and cannot overflow within 3 more steps). The earlier form computed z and |z|^2 every
// iteration (10 FP64 instructions) on a kernel that ncu showed bound by the FP64 pipe.
//
// Nothing is decided inside this kernel. A pixel for which a filter fired is rolled back to the last
// checkpoint (at most 16 iterations back, see k3_fast_body) and EXPORTED (state, index) while its lane-mates
// simply keep going; so is a pixel that cannot take another whole block (iteration limit or end of the
// orbit table less than 4 steps away). Exported pixels are parked on the reference orbit itself
// (delta = eps = 0, which stays 0 and never flags) until the lane's pass through the chunk ends, and
// are appended to the event queue with the same warp-aggregated reservation as the survivors that
// move on to the next chunk. After the sweep k3_finish<REQUEUE, SCALED, EVENTS> (k3_finish.cuh) runs every
// exported state to its end with the exact double comparisons — escape (+ smoothing), glitch (-> re-queue
// list), iteration limit, rebase at the end of the orbit and on from Z[0]; a false alarm simply keeps
// iterating there. Every decision is therefore the one k3_perturb.cuh and the oracle take.
//
// Quiet segments: most of a sample's life its delta is far too small to come near -Z within the next 16 steps. A
// per-segment bound on delta's high words (k3_filter.cuh: k3_seg_bound, rebuilt per frame from the orbit and the
// frame's largest pixel offset) proves that; a warp whose lanes all pass it runs the segment as k3_block_quiet — the
// recurrence alone, 6.5 SASS instructions per sample-iteration instead of 11.2 — and only the escape filter is
// looked at when the segment ends (profiles/r01p_*: full levels are entirely quiet, escape levels stay loud).
//
// (Measured and dropped, round 2 — profiles/r02_loud_queue_experiments.txt: exporting the few non-quiet slots of an
// otherwise quiet warp at the segment start so that the warp stays quiet. The level kernels got faster — cfg2's escape
// levels 5.96 -> 5.0 ms, full levels +3 % — but a sample is first non-quiet ~160 iterations before it escapes, not ~30:
// 1.35 G of cfg2's 54 G iterations moved to whichever kernel finished the exported states, and none of the three
// finishers tried — k3_level per level, k3_finish once per sweep, a one-state-per-lane pass of this kernel per level —
// ran them at more than 1/6 of this kernel's rate, where break-even needs 1/3. All three were bit-identical in results.)
//
// (Two earlier versions replayed flagged blocks inside the warp; on the level where half of the
// pixels escape that cost 3x, later 1.6x, the time of a full level, and latency-bound tail levels
// ran 5x slower than the simple kernel — profiles/r01b_*.)
#pragma once
#include "k3_checked.cuh"
#include "k3_filter.cuh"
#include "k3_perturb.cuh"

namespace nm {

constexpr int K3F_THREADS = 256;
#ifndef K3F_CTAS_PLAIN
#define K3F_CTAS_PLAIN 2    // CTAs per SM the plain / the scaled kernel is compiled for: 2 = up to 128 registers per
#endif                      // thread, 3 = 80 (measured both ways for both kernels: 3 is 1.7 % slower)
#ifndef K3F_CTAS_SCALED
#define K3F_CTAS_SCALED 2
#endif
#ifndef K3F_QUIET
#define K3F_QUIET 1         // quiet segments (k3_filter.cuh: k3_seg_bound) run without the per-iteration glitch filter
#endif
#define K3F_MIN_CTAS(P, SCALED) ((SCALED) ? K3F_CTAS_SCALED : K3F_CTAS_PLAIN)
// shared-memory bytes of the per-chunk tables (2Z, filter entries, escape words) / of everything k3_fast<P> needs
__host__ __device__ constexpr size_t k3f_table_bytes(int CH) {
  return (((size_t)(CH + 4) * (sizeof(double2) + sizeof(int4) + sizeof(int32_t))) + 15) & ~(size_t)15;
}
constexpr unsigned long long K3F_SPLIT_MIN = 4ULL * 148 * 2 * 256;   // levels with fewer states are not split (4 x K3F_SPARSE_MAX)

// One branch-free block of 4 iterations for the P slots of a lane, in place. jrel = the block's start index
// relative to the chunk. bad[s] |= the glitch filter (k3_filter.cuh) fired for slot s in this block.
template <int P, bool SCALED>
__device__ __forceinline__ void k3_block(double (&dr)[P], double (&di)[P], const double (&er)[P], const double (&ei)[P],
                                         const double (&S)[P], const uint32_t (&sm)[P],
                                         const double2* __restrict__ sZ2, const int4* __restrict__ sF, int jrel,
                                         bool (&bad)[P]) {
  double2 x2 = sZ2[jrel];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int jl = jrel + t + 1;   // the new delta pairs with Z[j + t + 1]
    const int4 f = sF[jl];
    double2 x2n = x2;
    if (t < 3) x2n = sZ2[jl];
#pragma unroll
    for (int s = 0; s < P; ++s) {
      double wr, wi;
      if (SCALED) { wr = __fma_rn(S[s], dr[s], x2.x); wi = __fma_rn(S[s], di[s], x2.y); }
      else { wr = x2.x + dr[s]; wi = x2.y + di[s]; }   // == fma(2, Z, delta): 2Z is exact
      const double ndr = __fma_rn(-di[s], wi, __fma_rn(dr[s], wr, er[s]));
      const double ndi = __fma_rn(di[s], wr, __fma_rn(dr[s], wi, ei[s]));
      dr[s] = ndr; di[s] = ndi;
      uint32_t ca = (uint32_t)__double2hiint(ndr) - (uint32_t)f.x;   // k3_filter_glitch
      uint32_t cb = (uint32_t)__double2hiint(ndi) - (uint32_t)f.z;
      if (SCALED) { ca |= sm[s]; cb |= sm[s]; }
      bad[s] = bad[s] | ((ca <= (uint32_t)f.y) & (cb <= (uint32_t)f.w));
    }
    x2 = x2n;
  }
}

// The same block for a QUIET segment (k3_filter.cuh: no slot of the warp can glitch before the segment ends): the
// recurrence alone.
template <int P, bool SCALED>
__device__ __forceinline__ void k3_block_quiet(double (&dr)[P], double (&di)[P], const double (&er)[P], const double (&ei)[P],
                                               const double (&S)[P], const double2* __restrict__ sZ2, int jrel) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const double2 x2 = sZ2[jrel + t];
#pragma unroll
    for (int s = 0; s < P; ++s) {
      double wr, wi;
      if (SCALED) { wr = __fma_rn(S[s], dr[s], x2.x); wi = __fma_rn(S[s], di[s], x2.y); }
      else { wr = x2.x + dr[s]; wi = x2.y + di[s]; }
      const double ndr = __fma_rn(-di[s], wi, __fma_rn(dr[s], wr, er[s]));
      const double ndi = __fma_rn(di[s], wr, __fma_rn(dr[s], wi, ei[s]));
      dr[s] = ndr; di[s] = ndi;
    }
  }
}

// What one launch of k3_fast works on. A level (orbit chunk k) is launched K3F_SUBS times (sub = 0..3). If few
// states escaped in the level before, launch 0 takes the whole chunk and the others return at once. Otherwise
// ("split": the level before lost more than 1/32 of its states) launch s takes the quarter-chunk
// [jbase + s*CH/4, jbase + (s+1)*CH/4) and hands its survivors to launch s+1 through two scratch queues, so the
// states are compacted globally every CH/4 instead of every CH iterations: where samples escape, a lane's
// slots die at different times and a dead slot costs the same issue cycles as a live one until the lane's
// pass ends (profiles/r01h_cfg2_levels.txt: such levels ran at 45-65 % of the rate of full ones). Decided on
// the device from the queue counters — every CTA of every launch of the level reads the same values — so the
// host enqueues blindly and nothing is read back.
constexpr int K3F_SUBS = 4;
struct K3Work {
  int jlo, jhi;                      // orbit indices this launch covers; passes end at jhi
  const PixState* cur; unsigned long long n_cur;
  PixState* next; unsigned long long* next_count;
  unsigned long long* head;
  unsigned fresh_begin; unsigned long long n_fresh;
  bool run;
};

__device__ __forceinline__ K3Work k3f_work(const K3Params& p) {
  K3Work w;
  const int CH = p.CH, k = p.k, jbase = k * CH;
  int l1 = jbase + CH;
  if (l1 > p.Jmax + 1) l1 = p.Jmax + 1;
  const unsigned long long n_in = (p.cur_count ? *p.cur_count : 0ULL) + (p.fresh_off[l1] - p.fresh_off[jbase]);
  bool split = false;
  if (p.sub_count && k > 0 && n_in >= p.split_min) {
    const unsigned long long prev_in = p.qcount[k - 1] + (p.fresh_off[jbase] - p.fresh_off[jbase - CH]);
    const unsigned long long prev_out = p.qcount[k];
    split = prev_in > prev_out && (prev_in - prev_out) * 32ULL > prev_in;
  }
  w.run = split || p.sub == 0;
  if (!split) {
    w.jlo = jbase; w.jhi = jbase + CH;
    w.cur = p.cur; w.n_cur = p.cur_count ? *p.cur_count : 0ULL;
    w.next = p.next; w.next_count = p.next_count;
  } else {
    const int q = CH / K3F_SUBS;
    w.jlo = jbase + p.sub * q; w.jhi = w.jlo + q;
    w.cur = p.sub == 0 ? p.cur : p.tmp[(p.sub - 1) & 1];
    w.n_cur = p.sub == 0 ? (p.cur_count ? *p.cur_count : 0ULL) : p.sub_count[p.sub - 1];
    w.next = p.sub == K3F_SUBS - 1 ? p.next : p.tmp[p.sub & 1];
    w.next_count = p.sub == K3F_SUBS - 1 ? p.next_count : &p.sub_count[p.sub];
  }
  w.head = p.head + p.sub;
  int f1 = w.jhi;
  if (f1 > p.Jmax + 1) f1 = p.Jmax + 1;
  int f0 = w.jlo;
  if (f0 > f1) f0 = f1;
  w.fresh_begin = p.fresh_off[f0];
  w.n_fresh = p.fresh_off[f1] - w.fresh_begin;  // multiple of the group size by construction
  return w;
}

// Per-thread slot records in shared memory (k3_fast keeps only delta and eps in registers).
// Layout [field][slot][thread]: conflict-free for the warp-wide accesses.
template <int P>
struct K3Slots {
  double2* ck;     // delta at the last checkpoint = the state an exported slot hands to k3_finish
  int32_t* pix;
  int32_t* off;
  int32_t* evj;    // orbit index of the checkpoint an exported slot was rolled back to
  __device__ __forceinline__ K3Slots(unsigned char* base) {
    ck = (double2*)base;
    pix = (int32_t*)(base + (size_t)P * K3F_THREADS * sizeof(double2));
    off = pix + P * K3F_THREADS;
    evj = off + P * K3F_THREADS;
  }
  static constexpr size_t bytes() { return (size_t)P * K3F_THREADS * (sizeof(double2) + 3 * sizeof(int32_t)); }
};

// SCALED: states carry a scale exponent (floatexp.cuh). A lane's slots share j, so "re-normalise
// before the step from j = 0 (mod 64)" is one test per lane and segment; the per-slot scale S = 2^e and
// eps / 2^e live in registers between re-normalisations.
//
// A lane's pass through the chunk runs in SEGMENTS that end at the orbit indices = 0 (mod 16) (at most 4 blocks).
// At the start of a segment the lane's deltas are checkpointed in shared memory; the filters' verdicts are
// looked at once per segment (the glitch flags are sticky; the escape filter looks at the segment's last delta:
// a slot that escaped earlier in the segment has |delta| growing without bound since — inf or NaN at worst,
// whose high words pass the filter too, and which disturb nobody: slots do not interact). A flagged slot is
// exported with its checkpoint state, so k3_finish replays at most 16 steps to reach the event.
// Against the one-block-at-a-time form (profiles/r01k_*: 250 instructions per 96 FP64, of which 34 register
// moves for the roll-back copy and 18 for the per-block escape test) this leaves ~150.
//
// Units: the launch's states in dealing order — the input queue padded to a multiple of G (the launch's group
// size: G consecutive units sit at the same orbit index), then the chunk-sorted fresh list (runs padded to G).
// A body<P> call deals the units [u0, u1) (u0 a multiple of P, P | G), P consecutive ones per lane.
// GL = the launch's group size: the slot records keep ITS layout in both calls (warps of one CTA may be in
// different calls at the same time).
template <int P, int GL, bool SCALED>
__device__ __forceinline__ void k3_fast_body(const K3Params& p, const K3Work& wk, PixState* events,
                                             unsigned long long u0, unsigned long long u1, unsigned long long u_cur,
                                             unsigned long long* head) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CH = p.CH;
  const int jbase = p.k * CH;
  // per index of the chunk: 2*Z[j] (w = 2Z + delta is a DADD of a table value: bit-identical to fma(2, Z, delta)
  // since 2Z is exact, and DADD issues ~9 % faster than DFMA on this part, profiles/r01_fp64_peak.json), the
  // glitch-filter entry and the escape-filter high word (k3_filter.cuh); built once per table upload
  const double2* sZ2 = (const double2*)smem_raw;
  const int4* sF = (const int4*)(smem_raw + (size_t)(CH + 4) * sizeof(double2));
  const int32_t* sE = (const int32_t*)(smem_raw + (size_t)(CH + 4) * (sizeof(double2) + sizeof(int4)));
  K3Slots<GL> slots(smem_raw + k3f_table_bytes(CH));

  const unsigned long long n_cur = wk.n_cur;
  const unsigned fresh_begin = wk.fresh_begin;
  const unsigned long long g_total = (u1 - u0) / P;   // deals of this call
  if (g_total == 0) return;

  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  const int jend = wk.jhi;   // passes end here (the chunk end, or the end of this launch's quarter)
  const int jcap = jend < p.Jmax ? jend : p.Jmax;  // a pass can never step beyond this index

  // registers: delta, eps (and the scale of a scaled state); bit s of `live` / `expo`: slot s is iterating /
  // was exported (its record sits in `slots`); neither: empty. Exported and empty slots are parked on the
  // reference orbit itself (delta = eps = 0 stays 0 and never flags).
  double dr[P], di[P], er[P], ei[P], S[P];
  int sc[P];
  uint32_t sm[P];   // SCALED: all-ones while the slot holds a scaled state (its d is not delta: k3_filter.cuh)
  unsigned live = 0, expo = 0;
  bool drained = false;
  int j = 0;
  unsigned long long executed = 0;
#pragma unroll
  for (int s = 0; s < P; ++s) { dr[s] = di[s] = er[s] = ei[s] = 0.0; S[s] = 1.0; sc[s] = 0; sm[s] = 0u; }

  for (;;) {
    // ---- re-deal: one group of P same-index pixels per lane (every lane is idle here) ------------
    live = 0; expo = 0;
    if (!drained) {
      unsigned long long base = 0;
      if (lane == 0) {
        if (((volatile unsigned long long*)p.ctr)[CTR_CANCEL]) base = g_total;
        else base = atomicAdd(head, 32ULL);
      }
      base = __shfl_sync(FULL_MASK, base, 0);
      if (base >= g_total) drained = true;
      else {
        unsigned long long g = base + lane;
        if (g < g_total) {
#pragma unroll
          for (int s = 0; s < P; ++s) {
            dr[s] = di[s] = er[s] = ei[s] = 0.0; sc[s] = 0; S[s] = 1.0; sm[s] = 0u;
            int pix = -1, off = -1;
            const unsigned long long u = u0 + g * P + s;
            if (u < u_cur) {
              if (u < n_cur) {
                PixState q = wk.cur[u];
                dr[s] = q.dr; di[s] = q.di; pix = q.pix; off = q.off; j = q.j; sc[s] = SCALED ? q.e : 0;
              }
            } else {
              int w = p.fresh_ids[fresh_begin + (unsigned)(u - u_cur)];
              if (w >= 0) {
                double2 d0 = p.fresh.d[w];
                dr[s] = d0.x; di[s] = d0.y; off = p.fresh.off[w]; j = p.fresh.j[w]; pix = p.fresh.pix[w];
                if (SCALED) sc[s] = p.fresh.e[w];
              }
            }
            if (pix >= 0) {
              EpsVal<SCALED> eps;
              eps.load(p.eps, pix);
              er[s] = eps.re_at(sc[s]);
              ei[s] = eps.im_at(sc[s]);
              if (SCALED) { S[s] = pow2d(sc[s]); sm[s] = sc[s] ? 0xffffffffu : 0u; }
              slots.pix[s * K3F_THREADS + tid] = pix;
              slots.off[s * K3F_THREADS + tid] = off;
              live |= 1u << s;
            }
          }
        }
      }
    }
    if (!__any_sync(FULL_MASK, live != 0)) break;

    // ---- one warp-synchronous pass through the chunk -----------------------------------------------
    const int j_in = j;  // every slot of this lane's group entered at this index
    // The largest high word (sign and, for a scaled slot, everything masked off) among the lane's deltas: what both
    // per-segment decisions look at — "could a slot have escaped" (>= the escape word of the index reached) and "is
    // the lane quiet in the coming segment" (< the segment's bound). Exact when computed (here, after a
    // re-normalisation, at every segment end); in between it can only be too high (exported slots are zeroed),
    // which errs on the loud side.
    auto hi_max = [&]() {
      int m = 0;
#pragma unroll
      for (int s = 0; s < P; ++s) {
        const int keep = SCALED ? (int)(0x7fffffffu & ~sm[s]) : 0x7fffffff;
        m = max(m, max(__double2hiint(dr[s]) & keep, __double2hiint(di[s]) & keep));
      }
      return m;
    };
    // Segment end: export (checkpoint state, index j_ck) every live slot whose sticky glitch flag is set or whose
    // last delta passes the escape filter. slots.ck already holds the state to hand over.
    auto export_flagged = [&](const bool (&bad)[P], int esc_hi, int j_ck) {
#pragma unroll
      for (int s = 0; s < P; ++s) {
        const int keep = SCALED ? (int)(0x7fffffffu & ~sm[s]) : 0x7fffffff;
        const bool flagged = bad[s] | ((__double2hiint(dr[s]) & keep) >= esc_hi) | ((__double2hiint(di[s]) & keep) >= esc_hi);
        if ((live & (1u << s)) && flagged) {
          live &= ~(1u << s); expo |= 1u << s;
          slots.evj[s * K3F_THREADS + tid] = j_ck;
          executed += (unsigned long long)(j_ck - j_in);
          dr[s] = di[s] = er[s] = ei[s] = 0.0;
        }
      }
    };
    int m_hi = hi_max();
    for (;;) {
      // Export the live slots that cannot take another whole block before the chunk end / table end /
      // their iteration limit — unless they simply reached the chunk end with room beyond it: those
      // move on to the next level after the loop. nb = whole blocks every remaining live slot can take.
      int nb = 0x7fffffff;
      bool any_room = false;
#pragma unroll
      for (int s = 0; s < P; ++s)
        if (live & (1u << s)) {
          const int jN = p.N - 1 - slots.off[s * K3F_THREADS + tid];
          const int lim = jcap < jN ? jcap : jN;
          const int room = (lim - j) >> 2;
          if (room > 0) { any_room = true; if (room < nb) nb = room; }
          else if (!(j == jend && j < p.Jmax && jN > j)) {
            live &= ~(1u << s); expo |= 1u << s;
            slots.ck[s * K3F_THREADS + tid] = make_double2(dr[s], di[s]);
            slots.evj[s * K3F_THREADS + tid] = j;
            executed += (unsigned long long)(j - j_in);
            dr[s] = di[s] = er[s] = ei[s] = 0.0;
          }
        }
      if (!any_room) nb = 0;
      if (!__any_sync(FULL_MASK, nb > 0)) break;

      // nb blocks in segments (see above); a flagged slot is rolled back to the segment's checkpoint, exported
      // and parked, its lane-mates keep going (their limits can only be farther away, so nb stays valid)
      for (int b = 0; __any_sync(FULL_MASK, b < nb);) {
        const bool act = b < nb;
        int n4 = 0;
        // quiet: no slot of this lane can glitch anywhere in the coming (whole) segment, decided from the size of
        // its deltas against the per-segment bound of k3_filter.cuh (k3_seg_bound). Lanes that sit this round out
        // do not object. If the whole warp is quiet the segment runs without the per-iteration filter.
        bool quiet = true;
        if (act) {
          if (SCALED && (j & RENORM_MASK) == 0) {
#pragma unroll
            for (int s = 0; s < P; ++s)
              if (live & (1u << s)) {
                pstate ps; ps.dr = dr[s]; ps.di = di[s]; ps.e = sc[s];
                state_renorm(ps);
                if (ps.e != sc[s]) {
                  EpsVal<SCALED> eps;
                  eps.load(p.eps, slots.pix[s * K3F_THREADS + tid]);
                  er[s] = eps.re_at(ps.e); ei[s] = eps.im_at(ps.e);
                  S[s] = pow2d(ps.e);
                }
                dr[s] = ps.dr; di[s] = ps.di; sc[s] = ps.e; sm[s] = ps.e ? 0xffffffffu : 0u;
              }
            m_hi = hi_max();
          }
          n4 = 4 - ((j >> 2) & 3);   // blocks up to the next index = 0 (mod 16)
          if (n4 > nb - b) n4 = nb - b;
          quiet = K3F_QUIET && n4 == 4 && m_hi < __ldg(&p.seg_hi[j >> 4]);
        }
        const bool warp_quiet = K3F_QUIET && __all_sync(FULL_MASK, quiet);
        if (act) {
          const int j_ck = j;
#pragma unroll
          for (int s = 0; s < P; ++s)
            if (live & (1u << s)) slots.ck[s * K3F_THREADS + tid] = make_double2(dr[s], di[s]);
          if (warp_quiet) {   // n4 == 4 in every active lane; no glitch flag can be set
#pragma unroll
            for (int q = 0; q < 4; ++q) k3_block_quiet<P, SCALED>(dr, di, er, ei, S, sZ2, j - jbase + 4 * q);
            j += 16;
            b += 4;
            m_hi = hi_max();
            const int esc_hi = sE[j - jbase];   // k3_filter_escape on the segment's last delta
            if (m_hi >= esc_hi) {
              bool none[P];
#pragma unroll
              for (int s = 0; s < P; ++s) none[s] = false;
              export_flagged(none, esc_hi, j_ck);
            }
          } else {
            bool bad[P];
#pragma unroll
            for (int s = 0; s < P; ++s) bad[s] = false;
            if (n4 == 4) {   // straight-line: the sticky flags stay in predicate registers
#pragma unroll
              for (int q = 0; q < 4; ++q) k3_block<P, SCALED>(dr, di, er, ei, S, sm, sZ2, sF, j - jbase + 4 * q, bad);
              j += 16;
            } else {
              for (int q = 0; q < n4; ++q) {
                k3_block<P, SCALED>(dr, di, er, ei, S, sm, sZ2, sF, j - jbase, bad);
                j += 4;
              }
            }
            b += n4;
            m_hi = hi_max();
            const int esc_hi = sE[j - jbase];
            bool any_bad = m_hi >= esc_hi;
#pragma unroll
            for (int s = 0; s < P; ++s) any_bad = any_bad | bad[s];
            if (any_bad) export_flagged(bad, esc_hi, j_ck);
          }
        }
      }
    }

    // ---- hand over (converged; warp-aggregated appends) --------------------------------------------
#pragma unroll
    for (int s = 0; s < P; ++s) {
      const bool toNext = (live >> s) & 1u;     // reached the chunk end alive
      const bool toEvents = (expo >> s) & 1u;
      if (toNext) executed += (unsigned long long)(j - j_in);
      unsigned long long slot = warp_reserve(wk.next_count, toNext);
      if (toNext) {
        PixState q; q.dr = dr[s]; q.di = di[s]; q.pix = slots.pix[s * K3F_THREADS + tid]; q.j = j;
        q.off = slots.off[s * K3F_THREADS + tid]; q.e = SCALED ? sc[s] : 0;
        wk.next[slot] = q;
      }
      slot = warp_reserve(&p.ctr[CTR_EVENTS], toEvents);
      if (toEvents) {
        const double2 d = slots.ck[s * K3F_THREADS + tid];
        PixState q; q.dr = d.x; q.di = d.y; q.pix = slots.pix[s * K3F_THREADS + tid]; q.j = slots.evj[s * K3F_THREADS + tid];
        q.off = slots.off[s * K3F_THREADS + tid]; q.e = SCALED ? sc[s] : 0;  // an exported slot is parked: its exponent is never re-normalised
        events[slot] = q;
      }
    }
  }

  for (int o = 16; o; o >>= 1) executed += __shfl_xor_sync(FULL_MASK, executed, o);
  if (lane == 0 && executed) atomicAdd(&p.ctr[CTR_EXECUTED], executed);
}

template <int P, bool SCALED>
__global__ void __launch_bounds__(K3F_THREADS, K3F_MIN_CTAS(P, SCALED))
k3_fast(K3Params p, PixState* events) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const K3Work wk = k3f_work(p);
  if (!wk.run) return;
  const unsigned long long u_cur = (wk.n_cur + P - 1) / P * P;   // the input queue, padded to whole groups
  const unsigned long long u_total = u_cur + wk.n_fresh;
  if (u_total == 0) return;

  // this launch's slice of the chunk tables: one bulk-async copy (TMA) per table and CTA
  {
    const int CH = p.CH, jbase = p.k * CH;
    const int first = wk.jlo - jbase;          // entries [first, first + nload) of the chunk's tables are needed
    int nload = p.Jmax + 1 - wk.jlo;
    if (nload > wk.jhi - wk.jlo + 1) nload = wk.jhi - wk.jlo + 1;
    const int nload4 = (nload + 3) & ~3;  // bulk copies move multiples of 16 bytes
    double2* sZ2 = (double2*)smem_raw;
    int4* sF = (int4*)(smem_raw + (size_t)(CH + 4) * sizeof(double2));
    int32_t* sE = (int32_t*)(smem_raw + (size_t)(CH + 4) * (sizeof(double2) + sizeof(int4)));
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t bz = (uint32_t)nload * (uint32_t)sizeof(double2);
      uint32_t be = (uint32_t)nload4 * (uint32_t)sizeof(int32_t);
      mbar_expect_tx(&bar, 2 * bz + be);
      bulk_g2s(sZ2 + first, p.Z2 + wk.jlo, bz, &bar);
      bulk_g2s((void*)(sF + first), p.filt + wk.jlo, bz, &bar);
      bulk_g2s((void*)(sE + first), p.esc_hi + wk.jlo, be, &bar);
    }
    mbar_wait(&bar, 0);
  }

  // Whole waves of lane groups (grid x 256 lanes x P states) run with P states per lane. What is left over
  // would cost a whole pass of its own however few lanes it fills; if it is at most 3/4 of a wave it is dealt
  // one state per lane instead: such a pass issues a quarter of the instructions and is latency bound at about
  // a quarter of the time (profiles/r01h_cfg2_levels.txt: 31 against 146 ns per iteration), and at most three
  // of them are needed. (Levels with less than 3/4 of a wave altogether run entirely that way.)
  unsigned long long u_a = u_total;
  if (P > 1) {
    const unsigned long long wave = (unsigned long long)gridDim.x * K3F_THREADS * P;
    const unsigned long long rem = u_total % wave;
    if (rem * 4 <= wave * 3) u_a = u_total - rem;
  }
  if (u_a > 0) k3_fast_body<P, P, SCALED>(p, wk, events, 0, u_a, u_cur, wk.head);
  if (P > 1 && u_a < u_total) k3_fast_body<1, P, SCALED>(p, wk, events, u_a, u_total, u_cur, wk.head + K3F_SUBS);
}

}  // namespace nm
