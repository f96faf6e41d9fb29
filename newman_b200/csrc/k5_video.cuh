// K5 — zoom-video in-betweening: VideoZoom::nextFrame (reference video.cpp:14-34). Between two key
// frames (the previous one P and the new one Q, rendered 1.5x deeper; both H x W = 1.5x the video frame,
// viewer.cpp:271-274) the reference writes `rate` frames: frame i shows P enlarged by lg = v^i and,
// blended over it with weight t = i/rate, Q shrunk by sm = (2/3) v^i, v = 1.5^(1/rate), both centred
// on an nr x nc canvas:   large = P.scaled(int(H lg), int(W lg));   small = Q.scaled(int(H sm), int(W sm));
//                         canvas.blit(large, (nr - large.nr)/2, (nc - large.nc)/2);  canvas.blend(small, t, ...).
// One launch writes all `rate` canvases (HBM-bound: 2 small reads per output pixel from L2-resident key
// frames, 3 bytes written).
//
// scaled / blit / blend live in libbyteimage, which the reference does not vendor: the conventions here
// are ours and unpinned: scaled = bilinear with pixel centres aligned
// (src = (dst + 0.5) * src_size / dst_size - 0.5, clamped), channel = trunc(value); blend = trunc((1-t) a + t b)
// in float32; canvas pixels outside `large` stay 0. Restated by oracle/oracle_p.c: oraclep_video_inbetween.
#pragma once
#include "nm_common.cuh"

namespace nm {

struct K5Params {
  const uint8_t* prev;  // [H*W*3]
  const uint8_t* next;
  int H, W, nr, nc, rate;
  float v;              // 1.5^(1/rate), evaluated by the host like the reference does (video.cpp:17: double pow, narrowed)
  uint8_t* out;         // [rate*nr*nc*3]
};

// bilinear sample of channel-interleaved src (H x W) scaled to (sh x sw), at scaled pixel (r, c)
__device__ __forceinline__ void k5_sample(const uint8_t* src, int H, int W, int sh, int sw, int r, int c, float* rgb) {
  float fy = (r + 0.5f) * ((float)H / (float)sh) - 0.5f;
  float fx = (c + 0.5f) * ((float)W / (float)sw) - 0.5f;
  fy = fy < 0.0f ? 0.0f : (fy > (float)(H - 1) ? (float)(H - 1) : fy);
  fx = fx < 0.0f ? 0.0f : (fx > (float)(W - 1) ? (float)(W - 1) : fx);
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + 1 < H ? y0 + 1 : H - 1, x1 = x0 + 1 < W ? x0 + 1 : W - 1;
  float wy = fy - (float)y0, wx = fx - (float)x0;
  const uint8_t* p00 = src + 3 * ((size_t)y0 * W + x0);
  const uint8_t* p01 = src + 3 * ((size_t)y0 * W + x1);
  const uint8_t* p10 = src + 3 * ((size_t)y1 * W + x0);
  const uint8_t* p11 = src + 3 * ((size_t)y1 * W + x1);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float top = (1.0f - wx) * p00[k] + wx * p01[k];
    float bot = (1.0f - wx) * p10[k] + wx * p11[k];
    rgb[k] = truncf((1.0f - wy) * top + wy * bot);   // the scaled image holds bytes
  }
}

__global__ void __launch_bounds__(256) k5_inbetween(K5Params p) {
  const long long per = (long long)p.nr * p.nc;
  const long long total = per * p.rate;
  const float v = p.v;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / per);
    const int rc = (int)(idx - (long long)i * per);
    const int r = rc / p.nc, c = rc - r * p.nc;
    float sm = 2.0f / 3.0f, lg = 1.0f;
    for (int k = 0; k < i; k++) { sm *= v; lg *= v; }   // the reference's running products (video.cpp:27-28)
    const float t = (float)i / (float)p.rate;
    const int lh = (int)(p.H * lg), lw = (int)(p.W * lg), sh = (int)(p.H * sm), sw = (int)(p.W * sm);
    float out[3] = {0.0f, 0.0f, 0.0f};
    {  // blit(large, (nr - lh)/2, (nc - lw)/2)
      const int lr = r - (p.nr - lh) / 2, lc = c - (p.nc - lw) / 2;
      if (lr >= 0 && lr < lh && lc >= 0 && lc < lw) k5_sample(p.prev, p.H, p.W, lh, lw, lr, lc, out);
    }
    {  // blend(small, t, (nr - sh)/2, (nc - sw)/2)
      const int sr = r - (p.nr - sh) / 2, sc = c - (p.nc - sw) / 2;
      if (sr >= 0 && sr < sh && sc >= 0 && sc < sw) {
        float s[3];
        k5_sample(p.next, p.H, p.W, sh, sw, sr, sc, s);
#pragma unroll
        for (int k = 0; k < 3; k++) out[k] = truncf((1.0f - t) * out[k] + t * s[k]);
      }
    }
    uint8_t* o = p.out + 3 * idx;
    o[0] = (uint8_t)out[0]; o[1] = (uint8_t)out[1]; o[2] = (uint8_t)out[2];
  }
}

}  // namespace nm
