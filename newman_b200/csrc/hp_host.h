// Host-side arbitrary-precision work for a frame (C++11, GMP mpf through the C API with explicit
// precision so it is thread-safe): the reference's probe search, reference orbit and series
// coefficients (reference mandelbrot.cpp:73-131), descended once to the flat double tables the
// device consumes, plus the separable per-column / per-row epsilon and coordinate arrays
// (mandelbrot.cpp:155-159, 271, 275) and the cardioid classification (63-71).
//
// Every mpf operation below is issued with the operand order and the temporary precisions the
// reference's gmpxx expressions produce (see compat/gmpxx.h for the rules), so the descended tables
// are bit-identical to the reference's by construction; tests/test_host_tables.py checks that
// against the compiled reference.
#ifndef NEWMAN_B200_HP_HOST_H
#define NEWMAN_B200_HP_HOST_H

#include <gmp.h>

#include <cstdint>
#include <utility>
#include <vector>

namespace newman_b200 {

struct ViewHP {  // borrowed view parameters (all at the working precision)
  mpf_srcptr center_re, center_im, sz_re, sz_im;
  int nr, nc, N;
  mp_bitcnt_t prec;  // working precision in bits (what setPrecision chose)
};

struct DeepTablesHost {
  int M = 0;
  bool has_escape = false;
  int probe_row = -1, probe_col = -1;
  std::vector<double> x_hi, x_lo, a, b, c;  // interleaved re,im
  // the same coefficients as mantissa (0.5 <= |m| < 1, truncated like descend) * 2^exponent: the
  // floatexp form the device uses once `finite` is false
  std::vector<double> a_m, b_m, c_m;
  std::vector<int32_t> a_e, b_e, c_e;
  std::vector<double> eps_re, eps_im;
  // low parts of the pixel offsets, trunc((pixel - X[0]) - eps): the double-double refinement pass (NM_MODE_DD) adds them
  std::vector<double> eps_re_lo, eps_im_lo;
  // eps as mantissa * 2^exponent (mpf_get_d_2exp, truncating): what scaled frames hand to the device
  std::vector<double> eps_re_m, eps_im_m;
  std::vector<int32_t> eps_re_e, eps_im_e;
  int pitch_exp = 0;   // binary exponent of the pixel pitch sz.re
  bool finite = true;  // false: some descended coefficient is inf/nan (reference would SIGFPE)
};

// mandelbrot.cpp:37-45: bits = max(64, (int)(64 - e + log2(1e-20))), e = binary exponent of sz.re
int precision_bits_for(mpf_srcptr sz_re);

// Pixel coordinate arrays in double (truncated): c_re[nc], c_im[nr] (mandelbrot.cpp:271, 275, 234)
void pixel_coords(const ViewHP& v, std::vector<double>& c_re, std::vector<double>& c_im);

// inCardioid (mandelbrot.cpp:63-71) for pixel (r, c), in mpf at the working precision.
bool in_cardioid_pixel(const ViewHP& v, int r, int c);

// Classify the whole view for the deep path: returns NM_CARDIOID_NONE / _ALL, or _MASK with `mask`
// filled (nr*nc bytes) when the cardioid/bulb boundary may cross the view.
int classify_cardioid(const ViewHP& v, int threads, std::vector<uint8_t>& mask);

// findProbe (mandelbrot.cpp:73-95): "first in scan order with the longest orbit". Lengths are found
// in parallel; returns the winning probe's (row, col) and its orbit length.
void find_probe(const ViewHP& v, int threads, int& row, int& col, int& length);

// The candidate list of findProbe in the reference's scan order (mandelbrot.cpp:77-83): rows nr/4,
// nr/2, 3nr/4 x every 2nd column, then every 2nd row x column nc/2. (row, col) pairs.
void probe_candidates(const ViewHP& v, std::vector<std::pair<int, int> >& cand);

// Exact orbit lengths (X.size() of computeOrbit) of the candidates listed in `which` (indices into
// cand), in parallel. len[k] belongs to which[k].
void probe_lengths(const ViewHP& v, const std::vector<std::pair<int, int> >& cand, const std::vector<int>& which,
                   int threads, std::vector<int>& len);

// computeOrbit + computeSeries for the reference point at pixel (row, col), descended; eps arrays
// relative to that orbit's X[0]. threads: 0 = hardware concurrency; with 4 or more the orbit and the three
// coefficient recurrences run as a pipeline on four threads, with 6 or more the products A^2 and A B get two stages of
// their own (same operations, bit-identical tables), else everything runs serially.
void build_tables(const ViewHP& v, int row, int col, DeepTablesHost& out, int threads = 0);

// true if the pooled, hand-laid-out mpf values of the pipelined table build behave exactly like mpf_init2 values under
// the running libgmp (checked once per process; if not, the build uses mpf_init2 per value)
bool host_mpf_layout_ok();
}  // namespace newman_b200
#endif
