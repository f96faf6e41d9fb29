/* Stand-in for libbyteimage's <byteimage/types.h> (not vendored by the reference; see
 * /root/reference/README.md:24). The escape path uses exactly one type from it: the (row, col)
 * pair `Pt` built at mandelbrot.cpp:78-83. TEST INFRASTRUCTURE ONLY (Oracle-R build). */
#ifndef NEWMAN_ORACLE_BYTEIMAGE_TYPES_H
#define NEWMAN_ORACLE_BYTEIMAGE_TYPES_H
#include <cmath>
#include <cstdio>
#include <vector>
namespace byteimage {
struct Pt {
  int r, c;
  Pt() : r(0), c(0) {}
  Pt(int r_, int c_) : r(r_), c(c_) {}
};
}
#endif
