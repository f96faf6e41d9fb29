// newman_b200/mandelbrot.h — source-compatible drop-in for newman's fractal engine headers
// (reference mandelbrot.h:22-53, grid.h:6-25, complex.h:6-35), backed by the B200 CUDA path.
//
// A caller written against the reference (viewer.cpp:76, 157-184, 186-253, 441-453) compiles
// unchanged when this directory precedes the reference's on the include path: same class names,
// same public fields (`error_tolerance, N, center, sz`), same member functions with the same
// argument meaning, same value semantics (copyable / movable; GPU state sits behind a shared
// handle). `grid.h` and `complex.h` in this directory forward here.
//
// What differs underneath:
//   * precompute() does the reference's host-side arbitrary-precision work (probe search, orbit,
//     series: mandelbrot.cpp:73-131, multi-threaded over probes) AND renders the whole raster on
//     the GPU through the C-ABI in <newman_b200.h>; computeRow(r) then only validates the frame.
//   * the per-pixel arbitrary-precision continuation (mandelbrot.cpp:209-224) is replaced by an
//     FP64 perturbation iteration with glitch detection and secondary reference orbits.
//   * there is no CPU fallback: without a CUDA device precompute()/computeRow() throw
//     std::runtime_error.
#ifndef NEWMAN_B200_MANDELBROT_H
#define NEWMAN_B200_MANDELBROT_H

#include <gmpxx.h>

#include <memory>
#include <vector>

// The classes below are implemented in libnewman_b200.so, which is otherwise built with hidden visibility: a C++
// caller (the reference's viewer.o, video.o) links against the library and finds exactly these.
#ifndef NEWMAN_B200_CXX_API
#define NEWMAN_B200_CXX_API __attribute__((visibility("default")))
#endif

// ---- value types (complex.h) ------------------------------------------------------------------
struct LPComplex {
  double re, im;
  constexpr LPComplex() : re(0.0), im(0.0) {}
  constexpr LPComplex(double r, double i) : re(r), im(i) {}
};
struct HPComplex {
  mpf_class re, im;
};
// Operation order is part of the contract (complex.h:19-31): the device kernels issue the same ops.
constexpr LPComplex operator+(const LPComplex& p, const LPComplex& q) { return LPComplex(p.re + q.re, p.im + q.im); }
constexpr LPComplex operator*(const LPComplex& p, const LPComplex& q) {
  return LPComplex(p.re * q.re - p.im * q.im, p.re * q.im + p.im * q.re);
}
constexpr LPComplex sq(const LPComplex& p) { return LPComplex(p.re * p.re - p.im * p.im, 2.0 * p.re * p.im); }
constexpr double sqMag(const LPComplex& p) { return p.re * p.re + p.im * p.im; }
inline LPComplex descend(const HPComplex& p) { return LPComplex(p.re.get_d(), p.im.get_d()); }  // truncating

// ---- output raster (grid.h) ---------------------------------------------------------------------
class RenderGrid {
public:
  struct EscapeValue {
    int iterations;
    float smoothing;
    EscapeValue(int it = 0, float sm = 0.0f) : iterations(it), smoothing(sm) {}
  };
  int nr, nc;
  std::vector<EscapeValue> values;  // row-major nr x nc, 8 bytes per sample == nm_escape

  RenderGrid() : nr(0), nc(0) {}
  RenderGrid(int rows, int cols) : nr(rows), nc(cols), values((size_t)rows * cols) {}
  EscapeValue& at(int r, int c) { return values[(size_t)r * nc + c]; }
  const EscapeValue& at(int r, int c) const { return values[(size_t)r * nc + c]; }
};

namespace newman_b200 {
class Engine;  // GPU context + last frame; shared between copies of a Mandelbrot
class Group;   // one context + one host thread + one NCCL rank per GPU (set `devices`), shared between copies
struct FrameInfo {
  bool hardware = false;
  int floatexp = 0;                 // 0 double series, 1 floatexp series, 2 floatexp series + scaled deltas
  int precision_bits = 64, orbit_len = 0, probe_row = -1, probe_col = -1;
  int references = 0;               // reference orbits used (1 + secondary rounds)
  unsigned long long executed_iters = 0, series_evals = 0, skipped_pixels = 0, glitched = 0, rebased = 0,
                     fixups = 0, kernel_launches = 0, ambiguous = 0;
  double host_precompute_s = 0, device_ms = 0, frame_s = 0;
  // probe search (GPU-assisted): delta updates spent on the candidates, candidates measured in mpf
  unsigned long long probe_iters = 0, probe_exact = 0;
  // 1: the exact lengths of the short-list agree with the perturbation counts that ranked it (or the list was widened
  // until they do); 0: they do not — the GPU-assisted probe may differ from the exhaustive search's (probe_search = 0)
  int probe_consistent = 1;
  unsigned long long refined = 0, refine_iters = 0;   // exact mode: samples repeated in double-double, their iterations
  double refine_ms = 0;             // exact mode: device time of the probe rendering + the double-double pass
  bool cancelled = false;           // the frame was abandoned by Mandelbrot::cancel(): the raster is not current
};
}  // namespace newman_b200

class NEWMAN_B200_CXX_API Mandelbrot {
protected:
  RenderGrid grid;
  std::shared_ptr<newman_b200::Engine> engine_;
  std::shared_ptr<newman_b200::Group> group_;
  struct Signature;  // view parameters the current raster was rendered for
  std::shared_ptr<Signature> rendered_;
  newman_b200::FrameInfo info_;

  void setPrecision();  // mandelbrot.cpp:37-56
  bool frameCurrent() const;
  void renderFrame();
  void renderFrameImpl();

public:
  double error_tolerance;
  int N;
  HPComplex center, sz;

  // Extensions (defaults reproduce the reference wherever the reference is defined).
  double glitch_tolerance;  // K3 glitch rule |X_n+d_n|^2 < glitch_tolerance*|X_n|^2
  int max_secondary;        // secondary reference rounds before the final rebasing pass
  int device;               // CUDA device ordinal
  std::vector<int> devices; // not empty: the frame is split over these GPUs — bands of `band_rows` grid rows
                            // dealt round-robin, one host thread per GPU, tables by ncclBroadcast, every GPU returning its
                            // bands over its own PCIe link (the beauty render, viewer.cpp:193-238, on all GPUs of the box);
                            // the raster is byte-identical to the one-GPU one
  int band_rows;            // rows per band (a multiple of the multisampling factor keeps colour-resolve blocks on one GPU)
  int host_threads;         // probe-search threads (0 = hardware concurrency)
  int probe_search;         // findProbe: 1 (default) GPU-assisted short-list + exact mpf check of the short-list (with a
                            // consistency guard: FrameInfo::probe_consistent); 0 the reference's exhaustive
                            // arbitrary-precision search (mandelbrot.cpp:73-95) — the only mode GUARANTEED to pick the
                            // reference's probe on every view; the two agree on every view tested down to 1e-97
  int exact;                // 1: exact mode — after the frame, the samples FP64 perturbation cannot resolve (found by rendering
                            // the frame a second time against the truncated orbit: counts that differ) are repeated in
                            // double-double arithmetic. ~2.3x the frame time; every sampled count of the 1e-50 bench view then
                            // equals the reference's. One GPU, frames with plain (not scaled) states.
  int force_floatexp;       // 0 automatic; 1 floatexp series, 2 also floatexp eps + scaled deltas, even where
                            // doubles suffice (verification: same raster wherever both are defined)

  Mandelbrot();
  Mandelbrot(int nr, int nc);

  void loadLegacy(const char* fn);  // 5-line view file (viewer.cpp:12-23)

  inline int rows() const { return grid.nr; }
  inline int cols() const { return grid.nc; }

  bool useHardware();
  void precompute();
  void computeRow(int r);
  // Not in the reference, whose viewer abandons a frame by no longer calling computeRow (viewer.cpp:177, 221-231): here
  // precompute() renders the whole frame, so abandoning it is a call of its own, from any other thread (the UI's).
  // precompute() then returns early with frameInfo().cancelled set; the next precompute() starts afresh.
  void cancel();

  HPComplex pointAt(int r, int c, int sc = 1) const;
  void translate(int dr, int dc, int sc = 1);
  void zoom(float scale);
  void zoomAt(float scale, int r, int c, int sc = 1);

  const RenderGrid::EscapeValue& at(int r, int c);
  RenderGrid::EscapeValue at(int r, int c, int sc);  // sc x sc average in iteration space

  void scaleUp(int sc);
  void scaleDown(int sc);

  void load(const char* fn);  // empty in the reference (mandelbrot.cpp:362-364); here: loadLegacy format
  void save(const char* fn);

  // Not in the reference: introspection for tests/bench, and the fused colour resolve
  // (viewer.cpp:84-124) on the raster that is already resident on the GPU.
  const newman_b200::FrameInfo& frameInfo() const { return info_; }
  const RenderGrid& raster() const { return grid; }
  void resolveRGB(const unsigned char* pal_rgb, int n_pal, int sc, bool smooth, unsigned char* rgb_out);
  void setPrecisionNow() { setPrecision(); }
  // findProbe on its own (probe_search selects the method); n_exact = candidates measured in mpf
  void findProbe(int& row, int& col, int& length, int* n_exact = nullptr);
};

#endif
