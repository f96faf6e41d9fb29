// Checked (one step at a time, exact double comparisons) form of the K3 iteration: used by K2 for the
// <= 3 steps that bring a fresh pixel to an index that is a multiple of 4, and by k3_events to resolve
// the pixels the branch-free kernel exported. Same operation order as k3_perturb.cuh / the oracle.
#pragma once
#include "nm_common.cuh"

namespace nm {

// Exact single step + decisions, shared by K2's alignment steps and k3_events. Returns:
//   0 continue, 1 escaped (r2 set), 2 glitched. State (dr, di, j) is advanced in place.
__device__ __forceinline__ int checked_step(const double2* __restrict__ Z, const double* __restrict__ gb, int Jmax,
                                            double er, double ei, double& dr, double& di, int& j, double& r2) {
  const double2 x = Z[j];
  const double2 y = Z[j + 1];
  double wr = __fma_rn(2.0, x.x, dr);
  double wi = __fma_rn(2.0, x.y, di);
  double ndr = __fma_rn(-di, wi, __fma_rn(dr, wr, er));
  double ndi = __fma_rn(di, wr, __fma_rn(dr, wi, ei));
  dr = ndr; di = ndi;
  ++j;
  double zr = y.x + dr, zi = y.y + di;
  double zmag = __fma_rn(zi, zi, zr * zr);
  if (zmag > BAILOUT2) { r2 = zr * zr + zi * zi; return 1; }  // sqMag as the reference forms it (complex.h:23)
  if (j != Jmax && zmag < gb[j]) return 2;
  return 0;
}

struct CheckedParams {  // what the checked path needs (K2 alignment steps, k3_events)
  const double2* Z;
  const double* gb;
  int Jmax, N;
  nm_escape* out;
  unsigned long long* ctr;
  FixupRec* fix;
  unsigned long long fix_cap;
  int32_t* rq_pix;
  int32_t* rq_iter;
  double log_bailout;
};

// Advance a state by up to `max_steps` checked steps (stopping at the iteration limit / end of the
// orbit table) and classify it exactly like k3_perturb.cuh. Returns true if the pixel continues
// (state updated: j is 0 after a rebase, else j_in + steps); false if it is finished (result, glitch
// marker or (N,0) written). *steps = delta updates performed.
__device__ __forceinline__ bool advance_checked(const CheckedParams& p, int pix, double er, double ei, int& off,
                                                double& dr, double& di, int& j, int max_steps, int* steps) {
  int n = 0;
  while (n < max_steps && j < p.Jmax && j + off + 1 < p.N) {
    double r2;
    int ev = checked_step(p.Z, p.gb, p.Jmax, er, ei, dr, di, j, r2);
    ++n;
    if (ev == 1) {
      bool unc;
      float s = smoothing_f32(r2, p.log_bailout, &unc);
      nm_escape v; v.iterations = j + off; v.smoothing = s;
      p.out[pix] = v;
      if (unc) push_fixup(p.ctr, p.fix, p.fix_cap, pix, r2);
      *steps = n;
      return false;
    }
    if (ev == 2) {
      unsigned long long slot = atomicAdd(&p.ctr[CTR_REQUEUE], 1ULL);
      p.rq_pix[slot] = pix;
      p.rq_iter[slot] = j + off;
      nm_escape v; v.iterations = -1; v.smoothing = 0.0f;
      p.out[pix] = v;
      *steps = n;
      return false;
    }
  }
  *steps = n;
  if (j + off + 1 >= p.N) {  // iteration limit (mandelbrot.cpp:226-228)
    nm_escape v; v.iterations = p.N; v.smoothing = 0.0f;
    p.out[pix] = v;
    return false;
  }
  if (j == p.Jmax) {  // outlived the orbit: continue from the virtual iterate Z[0] = 0 with delta = z
    const double2 xj = p.Z[j];
    dr = xj.x + dr;
    di = xj.y + di;
    off = j + off;
    j = 0;
    atomicAdd(&p.ctr[CTR_REBASED], 1ULL);
  }
  return true;
}

}  // namespace nm
