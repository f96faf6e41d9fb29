// Checked (one step at a time, exact double comparisons) form of the K3 iteration: used by K2 for the
// <= 3 steps that bring a fresh pixel to an index that is a multiple of 4, and (round 1) by k3_events to resolve
// the pixels the branch-free kernel exported. Same operation order as k3_perturb.cuh / the oracle.
// SCALED = true: the frame carries floatexp eps and (dr, di, e) states (floatexp.cuh "scaled
// perturbation state"); SCALED = false compiles to exactly the plain-double code.
#pragma once
#include "floatexp.cuh"
#include "nm_common.cuh"

namespace nm {

struct __align__(16) PixState {   // a K3 state in a queue
  double dr, di;
  int32_t pix;
  int32_t j;
  int32_t off;
  int32_t e;   // scale exponent: delta = (dr, di) * 2^e (0 in plain frames)
};
typedef PixState PixStateRec;

struct EpsTab {            // separable pixel offsets: per column / per row
  const double* re;        // [nc] doubles, or mantissas (0.5 <= |m| < 1) when re_e != nullptr
  const double* im;        // [nr]
  const int32_t* re_e;     // [nc] binary exponents (floatexp eps => scaled frames), else nullptr
  const int32_t* im_e;     // [nr]
  int nc;
};

template <bool SCALED> struct EpsVal;
template <> struct EpsVal<false> {
  double r0, i0;  // eps as the plain iteration uses it
  __device__ __forceinline__ void load(const EpsTab& t, int pix) {
    const int r = pix / t.nc, c = pix - r * t.nc;
    r0 = t.re[c]; i0 = t.im[r];
  }
  __device__ __forceinline__ double re_at(int) const { return r0; }
  __device__ __forceinline__ double im_at(int) const { return i0; }
};
template <> struct EpsVal<true> {
  fe r, i;
  double r0, i0;
  __device__ __forceinline__ void load(const EpsTab& t, int pix) {
    const int rr = pix / t.nc, c = pix - rr * t.nc;
    r = fe_norm(t.re[c], t.re_e[c]); i = fe_norm(t.im[rr], t.im_e[rr]);
    r0 = fe_to_double(r); i0 = fe_to_double(i);
  }
  __device__ __forceinline__ double re_at(int e) const { return eps_scaled(r, r0, e); }
  __device__ __forceinline__ double im_at(int e) const { return eps_scaled(i, i0, e); }
};

// Exact single step + decisions, shared by K2's alignment steps and the checked paths. Returns:
//   0 continue, 1 escaped (r2 set), 2 glitched (MODE_REQUEUE), 3 |z|^2 < |delta|^2: rebase (MODE_REBASE).
// State (dr, di, j) is advanced in place; (zr, zi) = z.
// S = 2^e and (er, ei) = eps / 2^e of the state (1 and eps for a plain state).
template <bool SCALED, int MODE = NM_MODE_REQUEUE>
__device__ __forceinline__ int checked_step(const double2* __restrict__ Z, const double* __restrict__ gb, int Jmax,
                                            double er, double ei, double S, double& dr, double& di, int& j, double& r2,
                                            double& zr, double& zi) {
  const double2 x = Z[j];
  const double2 y = Z[j + 1];
  double wr, wi;
  if (SCALED) { wr = __fma_rn(S, dr, 2.0 * x.x); wi = __fma_rn(S, di, 2.0 * x.y); }
  else { wr = __fma_rn(2.0, x.x, dr); wi = __fma_rn(2.0, x.y, di); }
  double ndr = __fma_rn(-di, wi, __fma_rn(dr, wr, er));
  double ndi = __fma_rn(di, wr, __fma_rn(dr, wi, ei));
  dr = ndr; di = ndi;
  ++j;
  if (SCALED) { zr = __fma_rn(S, dr, y.x); zi = __fma_rn(S, di, y.y); }
  else { zr = y.x + dr; zi = y.y + di; }
  double zmag = __fma_rn(zi, zi, zr * zr);
  if (zmag > BAILOUT2) { r2 = zr * zr + zi * zi; return 1; }  // sqMag as the reference forms it (complex.h:23)
  if (MODE == NM_MODE_REQUEUE) {
    if (j != Jmax && zmag < gb[j]) return 2;
  } else {
    double dmag;   // |delta|^2 of the value itself (k3_perturb.cuh: exact when S == 1)
    if (SCALED) { const double tr = S * dr, ti = S * di; dmag = __fma_rn(ti, ti, tr * tr); }
    else dmag = __fma_rn(di, di, dr * dr);
    if (zmag < dmag) return 3;
  }
  return 0;
}

struct CheckedParams {  // what the checked path needs (K2 alignment steps, k3_finish)
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
template <bool SCALED>
__device__ __forceinline__ bool advance_checked(const CheckedParams& p, int pix, const EpsVal<SCALED>& eps, int& off,
                                                double& dr, double& di, int& e, int& j, int max_steps, int* steps) {
  int n = 0;
  double S = 1.0, er = eps.r0, ei = eps.i0, zr = 0.0, zi = 0.0;
  if (SCALED) { S = pow2d(e); er = eps.re_at(e); ei = eps.im_at(e); }
  while (n < max_steps && j < p.Jmax && j + off + 1 < p.N) {
    if (SCALED && (j & RENORM_MASK) == 0) {
      pstate ps; ps.dr = dr; ps.di = di; ps.e = e;
      state_renorm(ps);
      dr = ps.dr; di = ps.di; e = ps.e;
      S = pow2d(e); er = eps.re_at(e); ei = eps.im_at(e);
    }
    double r2;
    int ev = checked_step<SCALED>(p.Z, p.gb, p.Jmax, er, ei, S, dr, di, j, r2, zr, zi);
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
    if (n == 0) {     // arrived here without a step of ours: form z = Z[j] + delta now
      const double2 xj = p.Z[j];
      if (SCALED) { zr = __fma_rn(S, dr, xj.x); zi = __fma_rn(S, di, xj.y); }
      else { zr = xj.x + dr; zi = xj.y + di; }
    }
    dr = zr; di = zi; e = 0;
    off = j + off;
    j = 0;
    atomicAdd(&p.ctr[CTR_REBASED], 1ULL);
  }
  return true;
}

}  // namespace nm
