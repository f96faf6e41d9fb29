// FP64 issue-rate probes: MEASUREMENT code, not part of the hot path. bench.py needs a measured FP64 instruction-issue
// peak for the roofline denominator (MEASURED_PEAKS.json carries HBM and bf16 figures only); these kernels provide it
// through nm_fp64_peak (include/newman_b200.h, "measurement helpers"). Own translation unit: nothing here is reachable
// from a render call.
#include <cuda_runtime.h>

#include "fp64_probe.h"

namespace nm {

// FP64 pipe peak probe (roofline denominator): 8 independent dependent-chains per thread.
template <int KIND>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double b, double c) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    if (KIND == 0) {
      a0 = __fma_rn(a0, b, c); a1 = __fma_rn(a1, b, c); a2 = __fma_rn(a2, b, c); a3 = __fma_rn(a3, b, c);
      a4 = __fma_rn(a4, b, c); a5 = __fma_rn(a5, b, c); a6 = __fma_rn(a6, b, c); a7 = __fma_rn(a7, b, c);
    } else if (KIND == 1) {
      a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c);
      a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c);
    } else {
      a0 = __dmul_rn(a0, b); a1 = __dmul_rn(a1, b); a2 = __dmul_rn(a2, b); a3 = __dmul_rn(a3, b);
      a4 = __dmul_rn(a4, b); a5 = __dmul_rn(a5, b); a6 = __dmul_rn(a6, b); a7 = __dmul_rn(a7, b);
    }
  }
  double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

// Issue-port probe: 8 independent DFMA chains interleaved with NI integer-pipe operations (IADD / LOP on 8
// independent accumulators) per loop iteration. Reports the DFMA rate: if it stays at the pure-DFMA rate the
// integer work rides along for free, if it drops the warp instructions share dispatch cycles (the model
// DESIGN.md uses for k3_fast: a warp-wide FP64 instruction holds the dispatch port for 2 cycles).
template <int NI>
__global__ void __launch_bounds__(256) fp64_int_mix_kernel(double* sink, int iters, double b, double c, int m0, int m1) {
  double a[8];
  int k[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { a[q] = threadIdx.x * 1e-9 + q; k[q] = threadIdx.x + q; }
#pragma unroll 2
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      a[q] = __fma_rn(a[q], b, c);
      if (NI >= 8) k[q] += m0;
      if (NI >= 16) k[q] ^= m1;
      if (NI >= 24) k[q] = max(k[q], m0 + q);
    }
  }
  double s = 0.0; int t = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) { s += a[q]; t ^= k[q]; }
  if (s == 123.456 || t == 0x12345678) sink[0] = s + t;
}

// Register-operand probe: 8 DFMA (or DADD) chains whose instructions read 3 (2, 1) DISTINCT 64-bit register
// operands that no neighbouring instruction shares, so the operand-reuse caches cannot help — unlike
// fp64_peak_kernel, whose b and c are the same registers in every instruction. OPS: 3 = fma(a[q], b[q], c[q]),
// 2 = fma(a[q], b, c[q]), 1 = fma(a[q], b, c) (== kind 0), 0 = a[q] + c[q] (DADD, 2 distinct).
template <int OPS>
__global__ void __launch_bounds__(256) fp64_operand_kernel(double* sink, int iters, double b0, double c0) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { a[q] = threadIdx.x * 1e-9 + q; b[q] = b0 + 1e-12 * (q + threadIdx.x); c[q] = c0 * (q + 1); }
#pragma unroll 2
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (OPS == 3) a[q] = __fma_rn(a[q], b[q], c[q]);
      else if (OPS == 2) a[q] = __fma_rn(a[q], b[0], c[q]);
      else if (OPS == 1) a[q] = __fma_rn(a[q], b[0], c[0]);
      else a[q] = __dadd_rn(a[q], c[q]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += a[q] + b[q] + c[q];
  if (s == 123.456) sink[0] = s;
}

// The K3 iteration body itself (4 pixels per thread, all operands in distinct registers, no memory):
// what the FP64 pipe sustains for this exact instruction mix (7 DFMA + 2 DADD + 1 DMUL per pixel-
// iteration with three different 64-bit register operands per DFMA).
__global__ void __launch_bounds__(256) fp64_k3mix_kernel(double* sink, int iters, double xr, double xi, double yr, double yi) {
  double dr[4], di[4], er[4], ei[4];
  int acc = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) { dr[s] = 1e-9 * (threadIdx.x + s); di[s] = -1e-9 * (s + 1); er[s] = 1e-12 * s; ei[s] = 2e-12; }
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      double wr = __fma_rn(2.0, xr, dr[s]);
      double wi = __fma_rn(2.0, xi, di[s]);
      double ndr = __fma_rn(-di[s], wi, __fma_rn(dr[s], wr, er[s]));
      double ndi = __fma_rn(di[s], wr, __fma_rn(dr[s], wi, ei[s]));
      dr[s] = ndr; di[s] = ndi;
      double zr = yr + ndr, zi = yi + ndi;
      acc |= __double2hiint(__fma_rn(zi, zi, zr * zr));  // integer pipe, like the kernel's candidate compare
    }
  }
  if (acc == 123456) sink[0] = acc + dr[0] + di[1];
}

}  // namespace nm

using namespace nm;

// kinds: see nm_fp64_peak in include/newman_b200.h. Returns a cudaError_t as int (0 = ok).
int nm_probe_run(void* stream_, int sm_count, int kind, int iters, double* sink, double* inst_per_s, double* ms_out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  // kinds 4-7: DFMA + 0/8/16/24 integer operations per 8 DFMA at full occupancy (64 warps/SM); 8-11: the
  // same at 16 warps/SM (k3_fast's occupancy). The rate returned counts the FP64 instructions only.
  // kinds 12-15: DFMA with 3 / 2 / 1 distinct register operands per instruction, DADD with 2 (64 warps/SM).
  const unsigned blocks = (unsigned)sm_count * ((kind >= 8 && kind < 12) ? 2 : 8);
  cudaEvent_t a, b;
  cudaError_t e;
  if ((e = cudaEventCreate(&a)) != cudaSuccess) return (int)e;
  if ((e = cudaEventCreate(&b)) != cudaSuccess) { cudaEventDestroy(a); return (int)e; }
  float best = 1e30f;
  for (int rep = 0; rep < 4 && e == cudaSuccess; rep++) {
    cudaEventRecord(a, stream);
    if (kind == 0) fp64_peak_kernel<0><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9);
    else if (kind == 1) fp64_peak_kernel<1><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9);
    else if (kind == 2) fp64_peak_kernel<2><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9);
    else if (kind == 3) fp64_k3mix_kernel<<<blocks, 256, 0, stream>>>(sink, iters / 4 + 1, 0.3, -0.2, 0.31, -0.19);
    else if (kind == 12) fp64_operand_kernel<3><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9);
    else if (kind == 13) fp64_operand_kernel<2><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9);
    else if (kind == 14) fp64_operand_kernel<1><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9);
    else if (kind == 15) fp64_operand_kernel<0><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9);
    else if ((kind & 3) == 0) fp64_int_mix_kernel<0><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9, 3, 5);
    else if ((kind & 3) == 1) fp64_int_mix_kernel<8><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9, 3, 5);
    else if ((kind & 3) == 2) fp64_int_mix_kernel<16><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9, 3, 5);
    else fp64_int_mix_kernel<24><<<blocks, 256, 0, stream>>>(sink, iters, 1.0000001, 1e-9, 3, 5);
    e = cudaGetLastError();
    cudaEventRecord(b, stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  if (e != cudaSuccess) return (int)e;
  double inst = (double)blocks * 256.0 * (double)iters * 8.0;
  if (kind == 3) inst = (double)blocks * 256.0 * (double)(iters / 4 + 1) * 4.0 * 10.0;  // 10 counted FP64 inst per pixel-iteration
  if (inst_per_s) *inst_per_s = inst / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return 0;
}
