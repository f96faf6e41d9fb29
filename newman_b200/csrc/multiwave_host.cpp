// Host-side multiwave palette (include/newman_b200/multiwave.h) + its C-ABI (nmp_*).
// O(N) table build on the host, exactly where the reference does it (viewer.cpp:66-68, 317, 381);
// the per-sample lookup/interpolation is the GPU's job (k4_resolve.cuh).
#include "../../include/newman_b200/multiwave.h"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <string>

#include "../../include/newman_b200.h"

#ifndef NEWMAN_B200_HAVE_BYTEIMAGE
namespace byteimage {

namespace {
unsigned char to_byte(float v) {
  float x = v * 255.0f + 0.5f;
  if (!(x > 0.0f)) return 0;
  if (x >= 255.0f) return 255;
  return (unsigned char)x;
}
float hue_channel(float p, float q, float t) {  // CSS Color 3, section 4.2.4
  if (t < 0.0f) t += 1.0f;
  if (t > 1.0f) t -= 1.0f;
  if (t < 1.0f / 6.0f) return p + (q - p) * 6.0f * t;
  if (t < 0.5f) return q;
  if (t < 2.0f / 3.0f) return p + (q - p) * (2.0f / 3.0f - t) * 6.0f;
  return p;
}
}  // namespace

void hsl2rgb(float h_deg, float s, float l, unsigned char& r, unsigned char& g, unsigned char& b) {
  float h = std::fmod(h_deg, 360.0f);
  if (h < 0.0f) h += 360.0f;
  h /= 360.0f;
  if (s <= 0.0f) { r = g = b = to_byte(l); return; }
  float q = l < 0.5f ? l * (1.0f + s) : l + s - l * s;
  float p = 2.0f * l - q;
  r = to_byte(hue_channel(p, q, h + 1.0f / 3.0f));
  g = to_byte(hue_channel(p, q, h));
  b = to_byte(hue_channel(p, q, h - 1.0f / 3.0f));
}

Color interp(const Color& a, const Color& b, float t) {
  auto mix = [t](unsigned char x, unsigned char y) {
    float v = (1.0f - t) * x + t * y;
    v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
    return (unsigned char)v;  // truncation, like k4_resolve.cuh
  };
  return Color(mix(a.r, b.r), mix(a.g, b.g), mix(a.b, b.b));
}

}  // namespace byteimage
#endif

using byteimage::Color;
using byteimage::hsl2rgb;
using byteimage::interp;

float MultiWaveGenerator::FloatCycle::value(int step) const {
  // position inside the cycle, in units of table entries; neighbours wrap around
  const int n = (int)values.size();
  float pos = values.size() * (step % period) / (float)period;
  int lo = (int)pos;
  int hi = (lo + 1) % n;
  pos -= lo;
  return (float)((1.0 - pos) * values[lo] + pos * values[hi]);
}

float MultiWaveGenerator::FloatWave::value(int step) const {
  const float tau = (float)(2.0 * 3.14159265358979);
  return amplitude * std::sin(step * tau / period);
}

void MultiWaveGenerator::load_filename(const char* fn) {
  std::ifstream in(fn);
  if (!in) return;
  int n = 0;
  in >> n;
  hue_cycles.assign(n > 0 ? n : 0, FloatCycle());
  for (FloatCycle& hc : hue_cycles) {
    in >> n;
    hc.values.assign(n > 0 ? n : 0, 0.0f);
    for (float& v : hc.values) in >> v;
    in >> hc.period;
  }
  in >> hue_period;
  in >> n;
  sat_cycle.values.assign(n > 0 ? n : 0, 0.0f);
  for (float& v : sat_cycle.values) in >> v;
  in >> sat_cycle.period;
  in >> n;
  lum_waves.assign(n > 0 ? n : 0, FloatWave());
  for (FloatWave& w : lum_waves) in >> w.amplitude >> w.period;
}

void MultiWaveGenerator::save_filename(const char* fn) const {
  FILE* fp = fopen(fn, "w");
  if (!fp) return;
  fprintf(fp, "%d\n", (int)hue_cycles.size());
  for (const FloatCycle& hc : hue_cycles) {
    fprintf(fp, "%d ", (int)hc.values.size());
    for (float v : hc.values) fprintf(fp, "%f ", v);
    fprintf(fp, "%d\n", hc.period);
  }
  fprintf(fp, "%d\n\n", hue_period);
  fprintf(fp, "%d\n", (int)sat_cycle.values.size());
  for (float v : sat_cycle.values) fprintf(fp, "%f ", v);
  fprintf(fp, "\n%d\n\n", sat_cycle.period);
  fprintf(fp, "%d\n", (int)lum_waves.size());
  for (const FloatWave& w : lum_waves) fprintf(fp, "%f %d\n", w.amplitude, w.period);
  fclose(fp);
}

CachedPalette MultiWaveGenerator::cache(int N) const {
  CachedPalette pal(N);
  if (hue_cycles.empty() || sat_cycle.values.empty()) return pal;
  // colour of hue cycle `which` at step i: lerp in RGB bytes between its two neighbouring hue nodes
  auto cycle_color = [this](int which, int i, float sat, float lum) {
    const FloatCycle& hc = hue_cycles[which];
    const int n = (int)hc.values.size();
    float pos = hc.values.size() * (i % hc.period) / (float)hc.period;
    int lo = (int)pos, hi = (lo + 1) % n;
    pos -= lo;
    Color a, b;
    hsl2rgb(hc.values[lo], sat, lum, a.r, a.g, a.b);
    hsl2rgb(hc.values[hi], sat, lum, b.r, b.g, b.b);
    return interp(a, b, pos);
  };
  for (int i = 0; i < N; i++) {
    float sat = sat_cycle.value(i);
    float lum = 0.0f;
    for (const FloatWave& w : lum_waves) lum += w.value(i);
    lum = (float)(1.0 / (1.0 + std::exp(-lum)));  // logistic squash of the wave sum
    // blend between two neighbouring hue cycles
    float pos = hue_cycles.size() * (i % hue_period) / (float)hue_period;
    int c0 = (int)pos, c1 = (c0 + 1) % (int)hue_cycles.size();
    pos -= c0;
    pal[i] = interp(cycle_color(c0, i, sat, lum), cycle_color(c1, i, sat, lum), pos);
  }
  return pal;
}

// ---- C-ABI ---------------------------------------------------------------------------------------
struct nmp_palette {
  MultiWaveGenerator mw;
};

extern "C" {
nmp_palette* nmp_create(void) { return new nmp_palette(); }
void nmp_destroy(nmp_palette* p) { delete p; }
int nmp_load_file(nmp_palette* p, const char* fn) {
  if (!p || !fn) return NM_EINVAL;
  std::ifstream probe(fn);
  if (!probe) return NM_EINVAL;
  p->mw.load_filename(fn);
  return NM_OK;
}
int nmp_save_file(const nmp_palette* p, const char* fn) {
  if (!p || !fn) return NM_EINVAL;
  p->mw.save_filename(fn);
  return NM_OK;
}
int nmp_clear(nmp_palette* p) {
  if (!p) return NM_EINVAL;
  p->mw = MultiWaveGenerator();
  return NM_OK;
}
int nmp_add_hue_cycle(nmp_palette* p, const float* hues_deg, int n, int period) {
  if (!p || !hues_deg || n < 1 || period < 1) return NM_EINVAL;
  MultiWaveGenerator::FloatCycle c;
  c.values.assign(hues_deg, hues_deg + n);
  c.period = period;
  p->mw.hue_cycles.push_back(c);
  return NM_OK;
}
int nmp_set_hue_period(nmp_palette* p, int period) {
  if (!p || period < 1) return NM_EINVAL;
  p->mw.hue_period = period;
  return NM_OK;
}
int nmp_set_sat_cycle(nmp_palette* p, const float* sats, int n, int period) {
  if (!p || !sats || n < 1 || period < 1) return NM_EINVAL;
  p->mw.sat_cycle.values.assign(sats, sats + n);
  p->mw.sat_cycle.period = period;
  return NM_OK;
}
int nmp_add_lum_wave(nmp_palette* p, float amplitude, int period) {
  if (!p || period < 1) return NM_EINVAL;
  MultiWaveGenerator::FloatWave w;
  w.amplitude = amplitude;
  w.period = period;
  p->mw.lum_waves.push_back(w);
  return NM_OK;
}
int nmp_cache_device(const nmp_palette* p, nm_ctx* ctx, int N, uint8_t* rgb_out) {
  if (!p || !ctx || N < 0) return NM_EINVAL;
  const MultiWaveGenerator& mw = p->mw;
  if (mw.hue_cycles.empty() || mw.sat_cycle.values.empty()) return NM_ESTATE;
  std::vector<int> counts, periods, lper;
  std::vector<float> hues, lamp;
  for (const MultiWaveGenerator::FloatCycle& hc : mw.hue_cycles) {
    counts.push_back((int)hc.values.size());
    periods.push_back(hc.period);
    hues.insert(hues.end(), hc.values.begin(), hc.values.end());
  }
  for (const MultiWaveGenerator::FloatWave& w : mw.lum_waves) { lamp.push_back(w.amplitude); lper.push_back(w.period); }
  return nm_palette_cache(ctx, (int)counts.size(), counts.data(), hues.data(), periods.data(), mw.hue_period,
                          (int)mw.sat_cycle.values.size(), mw.sat_cycle.values.data(), mw.sat_cycle.period, (int)lamp.size(),
                          lamp.empty() ? nullptr : lamp.data(), lper.empty() ? nullptr : lper.data(), N, rgb_out);
}

int nmp_cache(const nmp_palette* p, int N, uint8_t* rgb_out) {
  if (!p || N < 0 || (N > 0 && !rgb_out)) return NM_EINVAL;
  if (p->mw.hue_cycles.empty() || p->mw.sat_cycle.values.empty()) return NM_ESTATE;
  CachedPalette pal = p->mw.cache(N);
  for (int i = 0; i < N; i++) {
    rgb_out[3 * i] = pal[i].r;
    rgb_out[3 * i + 1] = pal[i].g;
    rgb_out[3 * i + 2] = pal[i].b;
  }
  return NM_OK;
}
}
