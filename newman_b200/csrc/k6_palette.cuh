// K6 — multiwave palette table on the device: MultiWaveGenerator::cache (reference multiwave.cpp:75-116;
// FloatCycle::value 5-12, FloatWave::value 14-17), one table entry per thread. SURVEY.md 8f-2: a
// palette edit then costs one tiny launch + the colour resolve (K4) of the raster that is already
// resident, without the N-entry table crossing PCIe.
//
// Same conventions as the host builder (multiwave_host.cpp; libbyteimage is not vendored, so RGB parity
// is unpinned): CSS hsl2rgb with round(255 v), interp = trunc(clamp((1-t) a + t b)) in float32. sin and
// exp are evaluated in double and narrowed, which reproduces glibc's float results except where those
// are not correctly rounded; the tests allow <= 1 LSB per channel (north star) and report the count.
#pragma once
#include "nm_common.cuh"

namespace nm {

struct K6Params {
  int n_cycles;
  const int* hue_counts;     // [n_cycles]
  const int* hue_offsets;    // [n_cycles] start of cycle k in hue_values
  const float* hue_values;   // concatenated hue nodes, degrees
  const int* hue_periods;    // [n_cycles]
  int hue_period;
  int n_sat;
  const float* sat_values;
  int sat_period;
  int n_lum;
  const float* lum_amp;
  const int* lum_period;
  int N;
  uint8_t* rgb;              // [3*N]
};

__device__ __forceinline__ uint8_t k6_to_byte(float v) {
  float x = v * 255.0f + 0.5f;
  if (!(x > 0.0f)) return 0;
  if (x >= 255.0f) return 255;
  return (uint8_t)x;
}
__device__ __forceinline__ float k6_hue_channel(float p, float q, float t) {
  if (t < 0.0f) t += 1.0f;
  if (t > 1.0f) t -= 1.0f;
  if (t < 1.0f / 6.0f) return p + (q - p) * 6.0f * t;
  if (t < 0.5f) return q;
  if (t < 2.0f / 3.0f) return p + (q - p) * (2.0f / 3.0f - t) * 6.0f;
  return p;
}
__device__ __forceinline__ void k6_hsl2rgb(float h_deg, float s, float l, uint8_t* rgb) {
  float h = fmodf(h_deg, 360.0f);
  if (h < 0.0f) h += 360.0f;
  h /= 360.0f;
  if (s <= 0.0f) { rgb[0] = rgb[1] = rgb[2] = k6_to_byte(l); return; }
  float q = l < 0.5f ? l * (1.0f + s) : l + s - l * s;
  float p = 2.0f * l - q;
  rgb[0] = k6_to_byte(k6_hue_channel(p, q, h + 1.0f / 3.0f));
  rgb[1] = k6_to_byte(k6_hue_channel(p, q, h));
  rgb[2] = k6_to_byte(k6_hue_channel(p, q, h - 1.0f / 3.0f));
}
__device__ __forceinline__ void k6_interp(const uint8_t* a, const uint8_t* b, float t, uint8_t* o) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float v = (1.0f - t) * a[k] + t * b[k];
    v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
    o[k] = (uint8_t)v;
  }
}
__device__ __forceinline__ float k6_cycle(const float* v, int n, int period, int step, int* lo_out, int* hi_out) {
  float pos = (size_t)n * (step % period) / (float)period;   // size_t product like the reference's values.size() * ...
  int lo = (int)pos, hi = (lo + 1) % n;
  *lo_out = lo; *hi_out = hi;
  return pos - lo;
}

__global__ void __launch_bounds__(256) k6_palette(K6Params p) {
  const float tau = (float)(2.0 * 3.14159265358979);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.N; i += gridDim.x * blockDim.x) {
    int lo, hi;
    float t = k6_cycle(p.sat_values, p.n_sat, p.sat_period, i, &lo, &hi);
    const float sat = (float)((1.0 - t) * p.sat_values[lo] + t * p.sat_values[hi]);
    float lum = 0.0f;
    for (int k = 0; k < p.n_lum; k++) lum += p.lum_amp[k] * (float)sin((double)(i * tau / p.lum_period[k]));
    lum = (float)(1.0 / (1.0 + (double)(float)exp((double)(-lum))));
    float ty = (size_t)p.n_cycles * (i % p.hue_period) / (float)p.hue_period;
    const int y0 = (int)ty, y1 = (y0 + 1) % p.n_cycles;
    ty -= y0;
    uint8_t col[2][3];
    const int ys[2] = {y0, y1};
#pragma unroll
    for (int w = 0; w < 2; w++) {
      const float* hv = p.hue_values + p.hue_offsets[ys[w]];
      const float tx = k6_cycle(hv, p.hue_counts[ys[w]], p.hue_periods[ys[w]], i, &lo, &hi);
      uint8_t a[3], b[3];
      k6_hsl2rgb(hv[lo], sat, lum, a);
      k6_hsl2rgb(hv[hi], sat, lum, b);
      k6_interp(a, b, tx, col[w]);
    }
    uint8_t out[3];
    k6_interp(col[0], col[1], ty, out);
    p.rgb[3 * i] = out[0]; p.rgb[3 * i + 1] = out[1]; p.rgb[3 * i + 2] = out[2];
  }
}

}  // namespace nm
