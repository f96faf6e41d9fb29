// Shared device-side definitions for the newman_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/newman_b200.h"

#ifndef __CUDA_ARCH__
#define NM_HD __host__ __device__
#else
#define NM_HD __host__ __device__
#endif

namespace nm {

constexpr unsigned FULL_MASK = 0xffffffffu;
// bailout = 1024, bailout2 = 2^20 (reference mandelbrot.cpp:58-59)
constexpr double BAILOUT2 = 1048576.0;
constexpr long long BAILOUT2_BITS = 0x4130000000000000LL;  // bit pattern of 2^20

// Device-side counters, one u64 each (cleared per frame).
enum Counter {
  CTR_NEXT = 0,       // K1 / K2 work cursor
  CTR_EXECUTED,       // executed iterations
  CTR_SKIPPED,        // cardioid/bulb pixels
  CTR_AMBIG,          // ambiguous-cardioid list length
  CTR_FIXUP,          // smoothing fix-up list length
  CTR_REQUEUE,        // glitch re-queue list length
  CTR_REBASED,        // rebase events
  CTR_SERIES,         // exact isUnstable evaluations
  CTR_CANCEL,         // nonzero => abandon frame
  CTR_Q_HEAD,         // K3 level kernel: input cursor
  CTR_Q_NEXT,         // K3: survivors appended for the next chunk
  CTR_Q_RESTART,      // K3: rebased pixels appended for the next sweep
  CTR_Q_CUR,          // K3: size of the current input queue
  CTR_DONE,           // K3: pixels finished
  CTR_CHECKED,        // K3 fast kernel: lane-steps taken in the checked (non-block) path
  CTR_EVENTS,         // K3 fast kernel: exported pixels waiting for k3_finish<.., EVENTS>
  CTR_CARRY,          // K3 fast path: states carried into the next sweep
  CTR_MINJ,           // lowest table index any K3 state of the coming sweep starts at (levels below are skipped)
  CTR_COUNT = 24
};

// Hand-over records between K2 / carry-over and K3 ("fresh" entries), SoA in HBM.
struct FreshArrays {
  double2* d;     // delta
  int32_t* j;     // table index, or -1: not a K3 state
  int32_t* off;   // it = j + off
  int32_t* pix;   // pixel id
  int32_t* e;     // scale exponent of delta (scaled frames only; 0 = plain state)
};

struct FixupRec {  // smoothing value to be re-evaluated with the host libm
  int32_t pix;
  int32_t pad;
  double r2;
};

// Smoothing (reference mandelbrot.cpp:133-136): 1 - log2(0.5*log(r2)/log(bailout)), narrowed to
// float32 on store (grid.h:11). CUDA's log/log2 are <=1 ulp but not bit-identical to glibc, so the
// double result s carries an absolute error of a few 1e-16 relative to the host's. If s lies that
// close to a float32 rounding boundary the narrowed value could differ; those pixels (about one in
// 1e6-1e7) are reported and re-evaluated by the host with the same libm the reference uses.
__device__ __forceinline__ float smoothing_f32(double r2, double log_bailout, bool* uncertain) {
  double s = 1.0 - log2((0.5 * log(r2)) / log_bailout);
  const double E = 4e-15;
  float f = (float)s;
  *uncertain = ((float)(s - E) != (float)(s + E));
  return f;
}

__device__ __forceinline__ void push_fixup(unsigned long long* ctr, FixupRec* list, unsigned long long cap,
                                           int pix, double r2) {
  unsigned long long k = atomicAdd(&ctr[CTR_FIXUP], 1ULL);
  if (k < cap) {
    list[k].pix = pix;
    list[k].pad = 0;
    list[k].r2 = r2;
  }
}

// Warp-aggregated reservation of `want`-many slots per lane (0 or 1) from a global cursor.
// Returns this lane's slot index (valid only if want).
__device__ __forceinline__ unsigned long long warp_reserve(unsigned long long* cursor, bool want) {
  unsigned m = __ballot_sync(FULL_MASK, want);
  int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (m) {
    int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
  }
  return base + __popc(m & ((1u << lane) - 1u));
}

}  // namespace nm
