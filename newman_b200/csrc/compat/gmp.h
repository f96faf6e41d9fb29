/* Minimal <gmp.h> stand-in: hand-declared prototypes for the mpf_* subset of GMP that
 * newman's hot path touches (reference call sites: mandelbrot.cpp:37-56 setPrecision,
 * 63-71 inCardioid, 97-131 orbit/series, 155-159 per-pixel epsilon; viewer.cpp:15-21).
 *
 * This image ships the GMP 6.3.0 *runtime* (libgmp.so.10) but no development headers, so the
 * build links `-l:libgmp.so.10` against these declarations. On a machine with real GMP headers,
 * drop this directory from the include path and everything below resolves to the real <gmp.h>.
 * The struct layout and symbol names are GMP's stable public ABI (they have not changed since 4.x).
 */
#ifndef NEWMAN_B200_COMPAT_GMP_H
#define NEWMAN_B200_COMPAT_GMP_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef unsigned long mp_bitcnt_t;
typedef long mp_exp_t;

typedef struct {
  int _mp_prec;      /* precision in limbs (result carries _mp_prec+1 limbs) */
  int _mp_size;      /* signed limb count; sign of the number */
  mp_exp_t _mp_exp;  /* exponent in limbs */
  mp_limb_t* _mp_d;
} __mpf_struct;

typedef __mpf_struct mpf_t[1];
typedef __mpf_struct* mpf_ptr;
typedef const __mpf_struct* mpf_srcptr;

#define NM_GMP_DECL(ret, name, args) ret __gmpf_##name args
NM_GMP_DECL(void, init, (mpf_ptr));
NM_GMP_DECL(void, init2, (mpf_ptr, mp_bitcnt_t));
NM_GMP_DECL(void, clear, (mpf_ptr));
NM_GMP_DECL(void, set, (mpf_ptr, mpf_srcptr));
NM_GMP_DECL(void, set_d, (mpf_ptr, double));
NM_GMP_DECL(void, set_si, (mpf_ptr, long));
NM_GMP_DECL(void, set_ui, (mpf_ptr, unsigned long));
NM_GMP_DECL(int, set_str, (mpf_ptr, const char*, int));
NM_GMP_DECL(void, set_prec, (mpf_ptr, mp_bitcnt_t));
NM_GMP_DECL(mp_bitcnt_t, get_prec, (mpf_srcptr));
NM_GMP_DECL(void, set_default_prec, (mp_bitcnt_t));
NM_GMP_DECL(mp_bitcnt_t, get_default_prec, (void));
NM_GMP_DECL(void, add, (mpf_ptr, mpf_srcptr, mpf_srcptr));
NM_GMP_DECL(void, sub, (mpf_ptr, mpf_srcptr, mpf_srcptr));
NM_GMP_DECL(void, mul, (mpf_ptr, mpf_srcptr, mpf_srcptr));
NM_GMP_DECL(void, div, (mpf_ptr, mpf_srcptr, mpf_srcptr));
NM_GMP_DECL(void, mul_ui, (mpf_ptr, mpf_srcptr, unsigned long));
NM_GMP_DECL(void, div_ui, (mpf_ptr, mpf_srcptr, unsigned long));
NM_GMP_DECL(void, mul_2exp, (mpf_ptr, mpf_srcptr, mp_bitcnt_t));
NM_GMP_DECL(void, div_2exp, (mpf_ptr, mpf_srcptr, mp_bitcnt_t));
NM_GMP_DECL(void, neg, (mpf_ptr, mpf_srcptr));
NM_GMP_DECL(void, abs, (mpf_ptr, mpf_srcptr));
NM_GMP_DECL(void, swap, (mpf_ptr, mpf_ptr));
NM_GMP_DECL(int, cmp, (mpf_srcptr, mpf_srcptr));
NM_GMP_DECL(int, cmp_d, (mpf_srcptr, double));
NM_GMP_DECL(int, cmp_si, (mpf_srcptr, long));
NM_GMP_DECL(double, get_d, (mpf_srcptr));
NM_GMP_DECL(double, get_d_2exp, (long*, mpf_srcptr));
NM_GMP_DECL(long, get_si, (mpf_srcptr));
NM_GMP_DECL(char*, get_str, (char*, mp_exp_t*, int, size_t, mpf_srcptr));
NM_GMP_DECL(size_t, out_str, (FILE*, int, size_t, mpf_srcptr));
#undef NM_GMP_DECL

extern const char* const __gmp_version;

#define mpf_init __gmpf_init
#define mpf_init2 __gmpf_init2
#define mpf_clear __gmpf_clear
#define mpf_set __gmpf_set
#define mpf_set_d __gmpf_set_d
#define mpf_set_si __gmpf_set_si
#define mpf_set_ui __gmpf_set_ui
#define mpf_set_str __gmpf_set_str
#define mpf_set_prec __gmpf_set_prec
#define mpf_get_prec __gmpf_get_prec
#define mpf_set_default_prec __gmpf_set_default_prec
#define mpf_get_default_prec __gmpf_get_default_prec
#define mpf_add __gmpf_add
#define mpf_sub __gmpf_sub
#define mpf_mul __gmpf_mul
#define mpf_div __gmpf_div
#define mpf_mul_ui __gmpf_mul_ui
#define mpf_div_ui __gmpf_div_ui
#define mpf_mul_2exp __gmpf_mul_2exp
#define mpf_div_2exp __gmpf_div_2exp
#define mpf_neg __gmpf_neg
#define mpf_abs __gmpf_abs
#define mpf_swap __gmpf_swap
#define mpf_cmp __gmpf_cmp
#define mpf_cmp_d __gmpf_cmp_d
#define mpf_cmp_si __gmpf_cmp_si
#define mpf_get_d __gmpf_get_d
#define mpf_get_d_2exp __gmpf_get_d_2exp
#define mpf_get_si __gmpf_get_si
#define mpf_get_str __gmpf_get_str
#define mpf_out_str __gmpf_out_str
#define gmp_version __gmp_version
#define mpf_sgn(F) ((F)->_mp_size < 0 ? -1 : (F)->_mp_size > 0)

#ifdef __cplusplus
}
#endif
#endif
