// K4 — colour resolve: palette lookup, smooth interpolation, sc x sc RGB-space box average, clip.
// Replaces FractalViewer::getColor / colorLine / recolor (reference viewer.cpp:84-124).
//
// libbyteimage (Color, interp, clip) is not vendored by the reference, so these conventions are
// OURS and palette/RGB parity is unpinned: interp(a,b,t) = trunc(clamp((1-t)*a + t*b))
// per channel in float32; clip(x) = trunc(clamp(x, 0, 255)); `pal[it-1]` with it == 0
// (viewer.cpp:92 reads out of bounds) is clamped to pal[0].
// HBM-bound: 8 B read per grid sample, 3 B written per image pixel.
#pragma once
#include "nm_common.cuh"

namespace nm {

struct K4Params {
  const nm_escape* grid;  // nr x nc samples
  int nr, nc;             // grid size (multiples of sc)
  const uint8_t* pal;     // 3*n_pal
  int n_pal, N, sc, smooth;
  uint8_t* rgb;           // (nr/sc) x (nc/sc) x 3
};

__device__ __forceinline__ float clampf(float x) { return fminf(fmaxf(x, 0.0f), 255.0f); }

__device__ __forceinline__ void get_color(const K4Params& p, nm_escape e, float& r, float& g, float& b) {
  if (e.iterations >= p.N || e.iterations < 0) { r = g = b = 0.0f; return; }
  int i1 = e.iterations < p.n_pal ? e.iterations : p.n_pal - 1;
  const uint8_t* c1 = p.pal + 3 * i1;
  if (!p.smooth) { r = c1[0]; g = c1[1]; b = c1[2]; return; }
  int i0 = i1 > 0 ? i1 - 1 : 0;
  const uint8_t* c0 = p.pal + 3 * i0;
  float t = e.smoothing, u = 1.0f - t;
  r = truncf(clampf(u * c0[0] + t * c1[0]));
  g = truncf(clampf(u * c0[1] + t * c1[1]));
  b = truncf(clampf(u * c0[2] + t * c1[2]));
}

__global__ void __launch_bounds__(256) k4_resolve(K4Params p) {
  const int onr = p.nr / p.sc, onc = p.nc / p.sc;
  long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long total = (long long)onr * onc;
  for (; o < total; o += stride) {
    int r = (int)(o / onc), c = (int)(o - (long long)r * onc);
    float sr = 0.0f, sg = 0.0f, sb = 0.0f;
    for (int r1 = r * p.sc; r1 < (r + 1) * p.sc; ++r1) {
      const nm_escape* row = p.grid + (size_t)r1 * p.nc;
      for (int c1 = c * p.sc; c1 < (c + 1) * p.sc; ++c1) {
        float cr, cg, cb;
        get_color(p, row[c1], cr, cg, cb);
        sr += cr; sg += cg; sb += cb;
      }
    }
    if (p.sc > 1) {
      float n = (float)(p.sc * p.sc);
      sr = truncf(clampf(sr / n)); sg = truncf(clampf(sg / n)); sb = truncf(clampf(sb / n));
    }
    uint8_t* out = p.rgb + 3 * o;
    out[0] = (uint8_t)sr; out[1] = (uint8_t)sg; out[2] = (uint8_t)sb;
  }
}

}  // namespace nm
