// K3 (run-to-completion form) — a frame, or the remainder of one, with so few states that the chunk-by-chunk
// level launches of k3_perturb.cuh / k3_fast.cuh are nothing but launch latency: the glitch re-queue rounds
// (a few hundred to a few thousand samples against a secondary reference), the final rebasing pass, the
// probe-search frames, and — with EVENTS — every state a k3_fast sweep exported (k3_fast.cuh).
// cfg2 spent ~800 of its launches and ~5 of its 41 ms per frame on such sweeps (profiles/r01l_*: every level
// of the orbit is launched because the host cannot know where the last state dies).
//
// One thread per state, straight through the orbit table in global memory (L1/L2 resident: 24 B per index,
// every thread walks it sequentially), same step, same decisions, same order as k3_level<MODE> and the oracle
// (oracle/oracle_p.c:355-396): escape, then glitch (MODE_REQUEUE) or rebase-when-|z|<|delta| (MODE_REBASE),
// then the iteration limit, then the end of the orbit (rebase onto Z[0] = 0). Latency bound by construction
// (one dependent chain per thread); the launch is sized so that every state has a thread.
#pragma once
#include "k3_checked.cuh"

namespace nm {

constexpr int K3_FINISH_THREADS = 128;
constexpr int K3_FINISH_BURST = 8;
constexpr long long K3_FINISH_MAX_STATES = 65536;   // above this the level kernels win (measured: DESIGN.md)

// EVENTS = true: the states are the PixState records k3_fast exported (k3_fast.cuh) — slots a filter fired for,
// slots at a limit, and the slots exported early because their delta came within reach of |Z| — and every one of
// them is finished here; nothing is carried into another sweep.
template <int MODE, bool SCALED, bool EVENTS = false>
__global__ void __launch_bounds__(K3_FINISH_THREADS) k3_finish(CheckedParams p, EpsTab eps_tab, FreshArrays f,
                                                              const unsigned long long* count, long long n_max,
                                                              const PixStateRec* events = nullptr) {
  __shared__ double2 park[K3_FINISH_BURST][K3_FINISH_THREADS];   // the burst's deltas, one column per thread
  const unsigned long long n = count ? *count : (unsigned long long)n_max;
  unsigned long long executed = 0, rebased = 0;
  for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < n;
       w += (unsigned long long)gridDim.x * blockDim.x) {
    int j, pix_, off, e;
    double dr, di;
    if (EVENTS) {
      const PixStateRec q = events[w];
      j = q.j; pix_ = q.pix; off = q.off; dr = q.dr; di = q.di; e = SCALED ? q.e : 0;
    } else {
      j = f.j[w];
      if (j < 0) continue;   // K2 finished this sample itself
      pix_ = f.pix[w];
      off = f.off[w];
      const double2 d0 = f.d[w];
      dr = d0.x; di = d0.y;
      e = SCALED ? f.e[w] : 0;
    }
    const int pix = pix_;
    if (EVENTS && j + off + 1 >= p.N) {   // exported AT its iteration limit: the last step's decision, in the oracle's order
      nm_escape v; v.iterations = p.N; v.smoothing = 0.0f;
      p.out[pix] = v;
      continue;
    }
    EpsVal<SCALED> eps;
    eps.load(eps_tab, pix);
    double S = 1.0, er = eps.r0, ei = eps.i0;
    if (SCALED) { S = pow2d(e); er = eps.re_at(e); ei = eps.im_at(e); }
    if (j >= p.Jmax) {   // defensive: a state handed over AT the end of the orbit rebases before its first step
      const double2 xj = p.Z[j];
      const double zr = SCALED ? __fma_rn(S, dr, xj.x) : xj.x + dr, zi = SCALED ? __fma_rn(S, di, xj.y) : xj.y + di;
      ++rebased; off = j + off; j = 0; dr = zr; di = zi;
      if (SCALED) { e = 0; S = 1.0; er = eps.r0; ei = eps.i0; }
    }
    // Bursts of up to K3_FINISH_BURST steps in which nothing is decided: the delta recurrence (3 dependent FP64
    // operations per step — at one warp per scheduler the ~40-cycle FP64 latency is all there is,
    // profiles/r01m_*) runs ahead while z, |z|^2 and the comparisons trail off the critical path into one bit
    // per step, and every step's delta is parked in shared memory. The first set bit is the step at which the
    // step-by-step loop would have stopped (all steps before it took no decision, so its operands are the same
    // bits): its delta is fetched back and the decision taken there, in the loop's order. Nothing is replayed.
    // (v1 of this kernel decided every step: 4x slower; v2 replayed flagged bursts: 1.6x slower, because in
    // MODE_REBASE a sample whose delta has grown to the size of z rebases every few steps.)
    for (unsigned burst = 0;; ++burst) {
      if ((burst & 63u) == 63u && ((volatile unsigned long long*)p.ctr)[CTR_CANCEL]) break;   // ~ every 500 steps
      int nb = p.Jmax - j;
      const int nN = p.N - 1 - off - j;   // steps until it = j + off reaches N - 1
      if (nN < nb) nb = nN;
      if (SCALED) { const int nr = RENORM_MASK + 1 - (j & RENORM_MASK); if (nr < nb) nb = nr; }
      if (nb > K3_FINISH_BURST) nb = K3_FINISH_BURST;
      if (SCALED && (j & RENORM_MASK) == 0) {
        pstate ps; ps.dr = dr; ps.di = di; ps.e = e;
        state_renorm(ps);
        dr = ps.dr; di = ps.di; e = ps.e;
        S = pow2d(e); er = eps.re_at(e); ei = eps.im_at(e);
      }
      const int j0 = j;
      double zr = 0.0, zi = 0.0;
      unsigned mask = 0;
      {
        double2 x = p.Z[j];
        // one undecided step; t = its position in the burst
        auto step = [&](int t) {
          const double2 y = p.Z[j + 1];
          double wr, wi;
          if (SCALED) { wr = __fma_rn(S, dr, 2.0 * x.x); wi = __fma_rn(S, di, 2.0 * x.y); }
          else { wr = __fma_rn(2.0, x.x, dr); wi = __fma_rn(2.0, x.y, di); }
          const double ndr = __fma_rn(-di, wi, __fma_rn(dr, wr, er));
          const double ndi = __fma_rn(di, wr, __fma_rn(dr, wi, ei));
          dr = ndr; di = ndi;
          park[t][threadIdx.x] = make_double2(dr, di);
          x = y;
          ++j;
          if (SCALED) { zr = __fma_rn(S, dr, y.x); zi = __fma_rn(S, di, y.y); }
          else { zr = y.x + dr; zi = y.y + di; }
          const double zmag = __fma_rn(zi, zi, zr * zr);
          bool c = !(zmag <= BAILOUT2);   // escape, or not a number any more (a burst may run past an escape)
          if (MODE == NM_MODE_REQUEUE) c = c | ((j != p.Jmax) & (zmag < p.gb[j]));
          else {
            double dmag;
            if (SCALED) { const double tr = S * dr, ti = S * di; dmag = __fma_rn(ti, ti, tr * tr); }
            else dmag = __fma_rn(di, di, dr * dr);
            c = c | (zmag < dmag);
          }
          mask |= (c ? 1u : 0u) << t;
        };
        if (nb == K3_FINISH_BURST) {
          // straight-line: a warp issues in order, so the trailing test of one step only gets out of the next
          // step's way if the scheduler can interleave them in the instruction stream (a rolled loop ran at
          // ~400 cycles per step: delta chain + test chain back to back; profiles/r01n_*)
#pragma unroll
          for (int t = 0; t < K3_FINISH_BURST; ++t) step(t);
        } else {
          for (int t = 0; t < nb; ++t) step(t);
        }
      }
      int ev = 0;
      double r2 = 0.0;
      if (mask) {   // back to the first step that decides something
        const int t = __ffs(mask) - 1;
        const double2 d = park[t][threadIdx.x];
        dr = d.x; di = d.y;
        j = j0 + t + 1;
        executed += (unsigned long long)(t + 1);
        const double2 y = p.Z[j];
        if (SCALED) { zr = __fma_rn(S, dr, y.x); zi = __fma_rn(S, di, y.y); }
        else { zr = y.x + dr; zi = y.y + di; }
        const double zmag = __fma_rn(zi, zi, zr * zr);
        if (zmag > BAILOUT2) { ev = 1; r2 = zr * zr + zi * zi; }   // sqMag as the reference forms it (complex.h:23)
        else if (MODE == NM_MODE_REQUEUE) { if (j != p.Jmax && zmag < p.gb[j]) ev = 2; }
        else {
          double dmag;
          if (SCALED) { const double tr = S * dr, ti = S * di; dmag = __fma_rn(ti, ti, tr * tr); }
          else dmag = __fma_rn(di, di, dr * dr);
          if (zmag < dmag) ev = 3;
        }
      } else {
        executed += (unsigned long long)nb;
      }
      if (ev == 1) {
        bool unc;
        const float sm = smoothing_f32(r2, p.log_bailout, &unc);
        nm_escape v; v.iterations = j + off; v.smoothing = sm;
        p.out[pix] = v;
        if (unc) push_fixup(p.ctr, p.fix, p.fix_cap, pix, r2);
        break;
      }
      if (ev == 2) {   // MODE_REQUEUE only
        const unsigned long long slot = atomicAdd(&p.ctr[CTR_REQUEUE], 1ULL);
        p.rq_pix[slot] = pix;
        p.rq_iter[slot] = j + off;
        nm_escape v; v.iterations = -1; v.smoothing = 0.0f;
        p.out[pix] = v;
        break;
      }
      if (j + off + 1 >= p.N) {   // iteration limit (mandelbrot.cpp:226-228)
        nm_escape v; v.iterations = p.N; v.smoothing = 0.0f;
        p.out[pix] = v;
        break;
      }
      if (ev == 3 || j == p.Jmax) {   // |z| < |delta| (MODE_REBASE), or the sample outlived the orbit:
        ++rebased;                    // continue from Z[0] = 0 with delta = z
        off = j + off; j = 0;
        dr = zr; di = zi;
        if (SCALED) { e = 0; S = 1.0; er = eps.r0; ei = eps.i0; }
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    executed += __shfl_xor_sync(FULL_MASK, executed, o);
    rebased += __shfl_xor_sync(FULL_MASK, rebased, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (executed) { atomicAdd(&p.ctr[CTR_EXECUTED], executed); atomicAdd(&p.ctr[CTR_CHECKED], executed); }
    if (rebased) atomicAdd(&p.ctr[CTR_REBASED], rebased);
  }
}

}  // namespace nm
