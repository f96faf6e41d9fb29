/* Minimal <gmpxx.h> stand-in: an *eager* mpf_class with the evaluation rules newman's sources
 * rely on (reference: complex.h:14-17 HPComplex, mandelbrot.cpp throughout, viewer.cpp:15-21).
 * Used only where the real GMP C++ headers are absent (this image has libgmp.so.10 but no
 * headers/libgmpxx). With real GMP installed, remove this directory from the include path.
 *
 * Semantics mirrored from gmpxx (and validated against libgmp.so.10 in tests/test_gmp_compat.py):
 *   - default ctor: value 0 at the *global default precision*; copy ctor: source precision;
 *   - copy-assign is mpf_set (destination keeps its precision); move = swap;
 *   - `= double` is mpf_set_d; `= const char*` is mpf_set_str base 10 at the destination precision;
 *   - a binary op yields a temporary at max(operand precisions); a builtin operand counts as the
 *     default precision; `mpf (op) double` goes through a 64-bit temporary; `int * mpf` is
 *     mpf_mul_ui (+ mpf_neg for negative ints); `mpf / int` is mpf_div_ui;
 *   - get_d() truncates toward zero (mpf_get_d).
 * Expression templates are not reproduced: results are bit-identical to gmpxx whenever operand
 * precision equals the default precision, which holds on every newman construction path
 * (setPrecision sets both, mandelbrot.cpp:47-51).
 */
#ifndef NEWMAN_B200_COMPAT_GMPXX_H
#define NEWMAN_B200_COMPAT_GMPXX_H

#include <gmp.h>
#include <algorithm>
#include <cstring>
#include <string>
#include <utility>

class mpf_class {
  mpf_t v_;

  static mp_bitcnt_t pmax(mp_bitcnt_t a, mp_bitcnt_t b) { return a > b ? a : b; }
  struct with_prec {};
  mpf_class(with_prec, mp_bitcnt_t p) { mpf_init2(v_, p); }

  template <typename F>
  static mpf_class with_double(const mpf_class& a, double d, F f) {
    mpf_class r(with_prec(), pmax(a.get_prec(), mpf_get_default_prec()));
    mpf_t t;
    mpf_init2(t, 8 * sizeof(double));
    mpf_set_d(t, d);
    f(r.v_, t);
    mpf_clear(t);
    return r;
  }

public:
  mpf_class() { mpf_init(v_); }
  mpf_class(const mpf_class& o) { mpf_init2(v_, mpf_get_prec(o.v_)); mpf_set(v_, o.v_); }
  mpf_class(mpf_class&& o) { mpf_init(v_); mpf_swap(v_, o.v_); }
  mpf_class(double d) { mpf_init(v_); mpf_set_d(v_, d); }
  mpf_class(int i) { mpf_init(v_); mpf_set_si(v_, i); }
  mpf_class(long i) { mpf_init(v_); mpf_set_si(v_, i); }
  mpf_class(double d, mp_bitcnt_t prec) { mpf_init2(v_, prec); mpf_set_d(v_, d); }
  mpf_class(const mpf_class& o, mp_bitcnt_t prec) { mpf_init2(v_, prec); mpf_set(v_, o.v_); }
  explicit mpf_class(const char* s) { mpf_init(v_); mpf_set_str(v_, s, 10); }
  mpf_class(const char* s, mp_bitcnt_t prec, int base = 10) { mpf_init2(v_, prec); mpf_set_str(v_, s, base); }
  ~mpf_class() { mpf_clear(v_); }

  mpf_class& operator=(const mpf_class& o) { if (this != &o) mpf_set(v_, o.v_); return *this; }
  mpf_class& operator=(mpf_class&& o) { mpf_swap(v_, o.v_); return *this; }
  mpf_class& operator=(double d) { mpf_set_d(v_, d); return *this; }
  mpf_class& operator=(int i) { mpf_set_si(v_, i); return *this; }
  mpf_class& operator=(long i) { mpf_set_si(v_, i); return *this; }
  mpf_class& operator=(const char* s) { mpf_set_str(v_, s, 10); return *this; }
  int set_str(const char* s, int base) { return mpf_set_str(v_, s, base); }

  mpf_ptr get_mpf_t() { return v_; }
  mpf_srcptr get_mpf_t() const { return v_; }
  mp_bitcnt_t get_prec() const { return mpf_get_prec(v_); }
  void set_prec(mp_bitcnt_t p) { mpf_set_prec(v_, p); }
  double get_d() const { return mpf_get_d(v_); }
  long get_si() const { return mpf_get_si(v_); }

  /* mpf (op) mpf */
  friend mpf_class operator+(const mpf_class& a, const mpf_class& b) {
    mpf_class r(with_prec(), pmax(a.get_prec(), b.get_prec())); mpf_add(r.v_, a.v_, b.v_); return r; }
  friend mpf_class operator-(const mpf_class& a, const mpf_class& b) {
    mpf_class r(with_prec(), pmax(a.get_prec(), b.get_prec())); mpf_sub(r.v_, a.v_, b.v_); return r; }
  friend mpf_class operator*(const mpf_class& a, const mpf_class& b) {
    mpf_class r(with_prec(), pmax(a.get_prec(), b.get_prec())); mpf_mul(r.v_, a.v_, b.v_); return r; }
  friend mpf_class operator/(const mpf_class& a, const mpf_class& b) {
    mpf_class r(with_prec(), pmax(a.get_prec(), b.get_prec())); mpf_div(r.v_, a.v_, b.v_); return r; }
  friend mpf_class operator-(const mpf_class& a) {
    mpf_class r(with_prec(), a.get_prec()); mpf_neg(r.v_, a.v_); return r; }

  /* mpf (op) double: 64-bit temporary, like gmpxx's __gmp_binary_* builtin overloads */
  friend mpf_class operator+(const mpf_class& a, double d) {
    return with_double(a, d, [&](mpf_ptr r, mpf_srcptr t) { mpf_add(r, a.v_, t); }); }
  friend mpf_class operator+(double d, const mpf_class& a) {
    return with_double(a, d, [&](mpf_ptr r, mpf_srcptr t) { mpf_add(r, t, a.v_); }); }
  friend mpf_class operator-(const mpf_class& a, double d) {
    return with_double(a, d, [&](mpf_ptr r, mpf_srcptr t) { mpf_sub(r, a.v_, t); }); }
  friend mpf_class operator-(double d, const mpf_class& a) {
    return with_double(a, d, [&](mpf_ptr r, mpf_srcptr t) { mpf_sub(r, t, a.v_); }); }
  friend mpf_class operator*(const mpf_class& a, double d) {
    return with_double(a, d, [&](mpf_ptr r, mpf_srcptr t) { mpf_mul(r, a.v_, t); }); }
  friend mpf_class operator*(double d, const mpf_class& a) {
    return with_double(a, d, [&](mpf_ptr r, mpf_srcptr t) { mpf_mul(r, t, a.v_); }); }
  friend mpf_class operator/(const mpf_class& a, double d) {
    return with_double(a, d, [&](mpf_ptr r, mpf_srcptr t) { mpf_div(r, a.v_, t); }); }

  /* mpf (op) int: the *_ui entry points, sign handled outside (gmpxx does the same) */
  friend mpf_class operator*(const mpf_class& a, long l) {
    mpf_class r(with_prec(), pmax(a.get_prec(), mpf_get_default_prec()));
    if (l >= 0) mpf_mul_ui(r.v_, a.v_, (unsigned long)l);
    else { mpf_mul_ui(r.v_, a.v_, 0UL - (unsigned long)l); mpf_neg(r.v_, r.v_); }
    return r; }
  friend mpf_class operator*(long l, const mpf_class& a) { return a * l; }
  friend mpf_class operator*(const mpf_class& a, int l) { return a * (long)l; }
  friend mpf_class operator*(int l, const mpf_class& a) { return a * (long)l; }
  friend mpf_class operator/(const mpf_class& a, long l) {
    mpf_class r(with_prec(), pmax(a.get_prec(), mpf_get_default_prec()));
    if (l >= 0) mpf_div_ui(r.v_, a.v_, (unsigned long)l);
    else { mpf_div_ui(r.v_, a.v_, 0UL - (unsigned long)l); mpf_neg(r.v_, r.v_); }
    return r; }
  friend mpf_class operator/(const mpf_class& a, int l) { return a / (long)l; }

  mpf_class& operator+=(const mpf_class& b) { mpf_add(v_, v_, b.v_); return *this; }
  mpf_class& operator-=(const mpf_class& b) { mpf_sub(v_, v_, b.v_); return *this; }
  mpf_class& operator*=(const mpf_class& b) { mpf_mul(v_, v_, b.v_); return *this; }
  mpf_class& operator*=(double d) {
    mpf_t t; mpf_init2(t, 8 * sizeof(double)); mpf_set_d(t, d); mpf_mul(v_, v_, t); mpf_clear(t); return *this; }
  mpf_class& operator+=(double d) {
    mpf_t t; mpf_init2(t, 8 * sizeof(double)); mpf_set_d(t, d); mpf_add(v_, v_, t); mpf_clear(t); return *this; }

  friend bool operator<(const mpf_class& a, const mpf_class& b) { return mpf_cmp(a.v_, b.v_) < 0; }
  friend bool operator>(const mpf_class& a, const mpf_class& b) { return mpf_cmp(a.v_, b.v_) > 0; }
  friend bool operator<=(const mpf_class& a, const mpf_class& b) { return mpf_cmp(a.v_, b.v_) <= 0; }
  friend bool operator>=(const mpf_class& a, const mpf_class& b) { return mpf_cmp(a.v_, b.v_) >= 0; }
  friend bool operator==(const mpf_class& a, const mpf_class& b) { return mpf_cmp(a.v_, b.v_) == 0; }
  friend bool operator<(const mpf_class& a, double d) { return mpf_cmp_d(a.v_, d) < 0; }
  friend bool operator>(const mpf_class& a, double d) { return mpf_cmp_d(a.v_, d) > 0; }
};

#endif
