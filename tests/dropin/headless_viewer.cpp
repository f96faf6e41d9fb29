// A headless caller written the way newman's viewer uses the engine (reference viewer.cpp:71-124 reset / getColor /
// colorLine, 157-184 render, 186-253 beautyRender, 271-285 video key frames, 329-354 and 429-448 navigation) — compiled
// against include/newman_b200/ and linked with libnewman_b200.so, nothing else. It is the proof that the drop-in is
// one at the C++ source level: the same member functions, public fields and value semantics, exported from the
// library. tests/test_dropin_cpp.py builds and runs it and compares what it prints with the C-ABI path.
//
//   headless_viewer state                      no GPU needed: navigation, copies, precision policy, palette table
//   headless_viewer render H W N [view file]   needs a GPU: frames through precompute()/computeRow()/at()
//   headless_viewer beauty H W SC N NGPU view [Z]  the beauty render (viewer.cpp:186-253) of a saved view on NGPU GPUs
//                                                  (Z: zoom in by 10^Z before loading, so that a deep file keeps its digits)
#include "mandelbrot.h"   // the include path decides: newman_b200's, not the reference's
#include "multiwave.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

using byteimage::Color;

static uint64_t fnv(const void* p, size_t n, uint64_t h = 0xcbf29ce484222325ULL) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ULL; }
  return h;
}

static std::string str(const mpf_class& f) {
  mp_exp_t e;
  char* s = mpf_get_str(nullptr, &e, 10, 0, f.get_mpf_t());
  std::string r = std::string(s) + "@" + std::to_string((long)e);
  free(s);
  return r;
}

struct HeadlessViewer {
  Mandelbrot mandel;
  MultiWaveGenerator mw;
  CachedPalette pal;
  int h, w, sc;
  bool smoothflag;
  std::vector<unsigned char> img;  // h x w x 3, interleaved

  HeadlessViewer(int h_, int w_) : h(h_), w(w_), sc(1), smoothflag(true), img((size_t)h_ * w_ * 3) { reset(); }

  void reset() {
    mandel = Mandelbrot(h, w);   // move-assignment of a temporary
    sc = 1;
    mw.load_filename("default.pal");
    pal = mw.cache(mandel.N);
  }
  Color getColor(const RenderGrid::EscapeValue& e) const {
    if (e.iterations >= mandel.N || e.iterations < 0) return Color(0);
    if (!smoothflag) return pal[e.iterations];
    return byteimage::interp(pal[e.iterations > 0 ? e.iterations - 1 : 0], pal[e.iterations], e.smoothing);
  }
  void colorLine(int r) {
    for (int c = 0; c < w; c++) {
      float x = 0, y = 0, z = 0;
      for (int r1 = r * sc; r1 < (r + 1) * sc; r1++)
        for (int c1 = c * sc; c1 < (c + 1) * sc; c1++) {
          Color k = getColor(mandel.at(r1, c1));
          x += k.r; y += k.g; z += k.b;
        }
      unsigned char* px = &img[((size_t)r * w + c) * 3];
      if (sc == 1) { px[0] = (unsigned char)x; px[1] = (unsigned char)y; px[2] = (unsigned char)z; }
      else {
        const float d = (float)(sc * sc);
        float v[3] = {x / d, y / d, z / d};
        for (int k = 0; k < 3; k++) px[k] = v[k] <= 0.f ? 0 : v[k] >= 255.f ? 255 : (unsigned char)v[k];
      }
    }
  }
  void render() {
    mandel.precompute();
    for (int r = 0; r < h; r++) {
      for (int r1 = r * sc; r1 < (r + 1) * sc; r1++) mandel.computeRow(r1);
      colorLine(r);
    }
  }
  uint64_t rasterHash() {
    uint64_t hh = 0xcbf29ce484222325ULL;
    for (int r = 0; r < mandel.rows(); r++)
      for (int c = 0; c < mandel.cols(); c++) {
        const RenderGrid::EscapeValue& e = mandel.at(r, c);
        hh = fnv(&e.iterations, 4, hh);
        hh = fnv(&e.smoothing, 4, hh);
      }
    return hh;
  }
  void printView(const char* tag) {
    printf("%s rows=%d cols=%d N=%d hw=%d c_re=%s c_im=%s sz_re=%s sz_im=%s\n", tag, mandel.rows(), mandel.cols(), mandel.N,
           (int)mandel.useHardware(), str(mandel.center.re).c_str(), str(mandel.center.im).c_str(), str(mandel.sz.re).c_str(),
           str(mandel.sz.im).c_str());
  }
};

static int run_state() {
  HeadlessViewer v(48, 64);
  v.printView("reset");
  printf("pal n=%d hash=%016llx\n", v.pal.size(), (unsigned long long)fnv(v.pal.bytes(), (size_t)3 * v.pal.size()));
  // navigation the way the mouse / key handlers do it
  v.mandel.zoomAt(2.0f, 10, 40);
  v.printView("zoomAt");
  v.mandel.translate(3, -5);
  v.printView("translate");
  for (int i = 0; i < 100; i++) v.mandel.zoom(1.5f);   // deep enough to leave the hardware path
  v.printView("zoom100");
  HPComplex p = v.mandel.pointAt(7, 9);
  printf("pointAt re=%s im=%s\n", str(p.re).c_str(), str(p.im).c_str());
  // the beauty render's choreography: save a copy, replace, restore
  Mandelbrot saved = v.mandel;
  v.mandel = Mandelbrot(96, 128);
  v.mandel.N = saved.N;
  v.mandel.center = saved.center;
  v.mandel.sz.re = saved.sz.re * ((double)saved.rows() / v.mandel.rows());
  v.mandel.sz.im = saved.sz.im * ((double)saved.rows() / v.mandel.rows());
  v.printView("beauty");
  v.mandel = saved;
  v.printView("restored");
  // multisample level switch on an (unrendered) raster
  v.mandel.scaleUp(2);
  v.printView("scaleUp");
  v.mandel.scaleDown(2);
  v.printView("scaleDown");
  // no CPU fallback: without a device the render entry points fail loudly
  try {
    v.mandel.precompute();
    v.mandel.computeRow(0);
    printf("precompute ok\n");
  } catch (const std::runtime_error& e) {
    printf("precompute runtime_error\n");
  }
  return 0;
}

static int run_render(int h, int w, int N, const char* view_file) {
  HeadlessViewer v(h, w);
  v.mandel.N = N;
  v.pal = v.mw.cache(v.mandel.N);
  v.render();                                    // the start-up view: plain double path
  printf("frame0 hw=%d raster=%016llx rgb=%016llx\n", (int)v.mandel.frameInfo().hardware, (unsigned long long)v.rasterHash(),
         (unsigned long long)fnv(v.img.data(), v.img.size()));
  if (view_file) {
    v.mandel.loadLegacy(view_file);              // F3 in the viewer
    v.pal = v.mw.cache(v.mandel.N);
    v.printView("loaded");
    v.render();
    printf("frame1 hw=%d M=%d refs=%d raster=%016llx rgb=%016llx\n", (int)v.mandel.frameInfo().hardware, v.mandel.frameInfo().orbit_len,
           v.mandel.frameInfo().references, (unsigned long long)v.rasterHash(), (unsigned long long)fnv(v.img.data(), v.img.size()));
    // 2x multisampling the way the viewer switches levels, then a fresh render of the finer grid
    v.mandel.scaleUp(2);
    v.sc = 2;
    v.render();
    std::vector<unsigned char> host_rgb = v.img;
    printf("frame2 rows=%d raster=%016llx rgb=%016llx\n", v.mandel.rows(), (unsigned long long)v.rasterHash(),
           (unsigned long long)fnv(v.img.data(), v.img.size()));
    // the same colours from the raster that is still resident on the GPU (K4)
    std::vector<unsigned char> dev_rgb(host_rgb.size());
    v.mandel.resolveRGB(v.pal.bytes(), v.pal.size(), v.sc, v.smoothflag, dev_rgb.data());
    printf("k4 equal=%d\n", (int)(dev_rgb == host_rgb));
    // a copy shares the frame; changing a public field invalidates it on the next computeRow
    Mandelbrot copy = v.mandel;
    printf("copy at=%d\n", copy.at(5, 7).iterations == v.mandel.at(5, 7).iterations);
    v.mandel.N = v.mandel.N / 2;
    v.mandel.computeRow(0);
    int over = 0;
    for (int c = 0; c < v.mandel.cols(); c++) over += v.mandel.at(0, c).iterations > v.mandel.N;
    printf("rerender N=%d over=%d\n", v.mandel.N, over);
  }
  return 0;
}

// FractalViewer::beautyRender (viewer.cpp:186-253) the way the viewer does it — save a copy, replace `mandel` by a
// supersampled one, copy N / centre / scaled sz over, precompute(), computeRow() for every row, recolor(), restore —
// with the one line a multi-GPU box adds: `mandel.devices`. (The reference forgets setPrecision after assigning the
// centre; a caller that wants the deep centre sets the precision first, as here.)
static int run_beauty(int h, int w, int sc, int N, int ngpu, const char* view_file, int prezoom) {
  HeadlessViewer v(h, w);
  v.mandel.N = N;
  // loadLegacy parses at the CURRENT precision of the fields (mandelbrot.cpp:24-27): a viewer that is shallower than the
  // file loses the centre's digits. A deep location is loaded by a viewer that is already deep: zoom in first.
  for (int i = 0; i < prezoom; i++) v.mandel.zoom(10.0f);
  v.mandel.loadLegacy(view_file);
  v.pal = v.mw.cache(v.mandel.N);
  uint64_t raster[2] = {0, 0}, rgb[2] = {0, 0};
  double secs[2] = {0, 0};
  for (int pass = 0; pass < 2; pass++) {           // pass 0: one GPU; pass 1: all of them — the frames must be identical
    Mandelbrot saved = v.mandel;                    // viewer.cpp:193
    v.mandel = Mandelbrot(h * sc, w * sc);          // :195-196
    v.mandel.N = saved.N;                           // :197
    v.mandel.sz.re = saved.sz.re / sc;              // :199-200 (scaled sz first, so that ...
    v.mandel.sz.im = saved.sz.im / sc;
    v.mandel.zoom(1.0f);                            // ... setPrecision runs before the centre is assigned)
    v.mandel.center = saved.center;                 // :198
    if (pass == 1) {
      v.mandel.devices.clear();
      for (int d = 0; d < ngpu; d++) v.mandel.devices.push_back(d);
      v.mandel.band_rows = sc;
    }
    v.sc = sc;
    v.mandel.precompute();                          // :202
    for (int r = 0; r < v.mandel.rows(); r++) v.mandel.computeRow(r);   // :209-210
    for (int r = 0; r < h; r++) v.colorLine(r);     // recolor(), :236-238
    raster[pass] = v.rasterHash();
    rgb[pass] = fnv(v.img.data(), v.img.size());
    secs[pass] = v.mandel.frameInfo().frame_s;
    printf("beauty pass=%d gpus=%d rows=%d cols=%d M=%d refs=%d executed=%llu frame_s=%.3f device_ms=%.2f raster=%016llx rgb=%016llx\n", pass,
           pass ? ngpu : 1, v.mandel.rows(), v.mandel.cols(), v.mandel.frameInfo().orbit_len, v.mandel.frameInfo().references,
           (unsigned long long)v.mandel.frameInfo().executed_iters, secs[pass], v.mandel.frameInfo().device_ms,
           (unsigned long long)raster[pass], (unsigned long long)rgb[pass]);
    v.mandel = saved;                               // :250-252
    v.sc = 1;
  }
  printf("beauty identical=%d\n", (int)(raster[0] == raster[1] && rgb[0] == rgb[1]));
  return raster[0] == raster[1] && rgb[0] == rgb[1] ? 0 : 3;
}

int main(int argc, char** argv) {
  try {
    if (argc >= 2 && !strcmp(argv[1], "state")) return run_state();
    if (argc >= 8 && !strcmp(argv[1], "beauty"))
      return run_beauty(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), argv[7], argc > 8 ? atoi(argv[8]) : 0);
    if (argc >= 5 && !strcmp(argv[1], "render")) return run_render(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argc > 5 ? argv[5] : nullptr);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 2;
  }
  fprintf(stderr, "usage: headless_viewer state | render H W N [view file] | beauty H W SC N NGPU view-file\n");
  return 1;
}
