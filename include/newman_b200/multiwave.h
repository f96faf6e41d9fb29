// newman_b200/multiwave.h — drop-in for newman's "multiwave" palette generator
// (reference multiwave.h:8-37, multiwave.cpp): hue cycles x saturation cycle x luminance waves,
// `.pal` text I/O, and cache(N) -> N-entry RGB table that the colour resolve (K4) indexes.
//
// The reference builds on libbyteimage's CachedPalette / Color / hsl2rgb / interp, which it does not
// vendor (README.md:24). When that library is available define NEWMAN_B200_HAVE_BYTEIMAGE and its
// types are used; otherwise the minimal stand-ins below apply, with THIS repo's colour conventions
// (DESIGN.md "colour"): hsl2rgb = the CSS HSL algorithm, channel = round(255 v); interp(a,b,t) =
// trunc(clamp((1-t) a + t b)) per channel in float32. Palette/RGB parity is therefore unpinned.
#ifndef NEWMAN_B200_MULTIWAVE_H
#define NEWMAN_B200_MULTIWAVE_H

#include <vector>

#ifndef NEWMAN_B200_CXX_API
#define NEWMAN_B200_CXX_API __attribute__((visibility("default")))   // exported from libnewman_b200.so (see mandelbrot.h)
#endif

#ifdef NEWMAN_B200_HAVE_BYTEIMAGE
#include <byteimage/palette.h>
using byteimage::CachedPalette;
#else
namespace byteimage {
struct Color {
  unsigned char r, g, b;
  Color() : r(0), g(0), b(0) {}
  explicit Color(unsigned char v) : r(v), g(v), b(v) {}
  Color(unsigned char r_, unsigned char g_, unsigned char b_) : r(r_), g(g_), b(b_) {}
};
class CachedPalette {
  std::vector<Color> c_;
public:
  CachedPalette() {}
  explicit CachedPalette(int n) : c_(n > 0 ? n : 0) {}
  int size() const { return (int)c_.size(); }
  Color& operator[](int i) { return c_[i]; }
  const Color& operator[](int i) const { return c_[i]; }
  const unsigned char* bytes() const { return c_.empty() ? nullptr : &c_[0].r; }  // 3*size(), r g b r g b ...
};
NEWMAN_B200_CXX_API void hsl2rgb(float h_deg, float s, float l, unsigned char& r, unsigned char& g, unsigned char& b);
NEWMAN_B200_CXX_API Color interp(const Color& a, const Color& b, float t);
}  // namespace byteimage
using byteimage::CachedPalette;
#endif

class NEWMAN_B200_CXX_API MultiWaveGenerator {
public:
  struct NEWMAN_B200_CXX_API FloatCycle {  // piecewise-linear cyclic table (multiwave.cpp:5-12)
    std::vector<float> values;
    int period = 1;
    float value(int step) const;
  };
  struct NEWMAN_B200_CXX_API FloatWave {   // amplitude * sin(step * tau / period) (multiwave.cpp:14-17)
    float amplitude = 1.0f;
    int period = 1;
    float value(int step) const;
  };

  std::vector<FloatCycle> hue_cycles;  // hue nodes in degrees
  int hue_period = 1;
  FloatCycle sat_cycle;
  std::vector<FloatWave> lum_waves;

  void load_filename(const char* fn);        // multiwave.cpp:19-48 (silently ignores a missing file)
  void save_filename(const char* fn) const;  // multiwave.cpp:50-73
  CachedPalette cache(int N) const;          // multiwave.cpp:75-116
};

#endif
