/* oracle/llz_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the lambda-lanczos hot path: a plain-C (C99) restatement of the reference's Krylov iteration
 * (LambdaLanczos<T>::run / run_iteration, Exponentiator<T>::run / taylor_run, the util:: vector helpers and the
 * implicit-shift QR tridiagonal solver).  Sequential, single-accumulator, same operation order as the reference.
 *
 * PARITY PINNED: this restatement is checked (tests/test_oracle.py) against
 *   (1) the reference's own known-answer tests (test/lambda_lanczos_test.cpp, test/exponentiator_test.cpp), and
 *   (2) the reference itself, compiled from /root/reference by oracle/Makefile into oracle/_ref/libllz_ref.so,
 *       directly where that library is present and through fixtures in tests/golden/ generated from it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (lambda-lanczos_b200/) never links, imports or calls it.
 *
 * Build: see oracle/Makefile  ->  oracle/libllz_oracle.so
 */
#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- real tridiagonal solvers ---- */
#define R float
#define RSFX f32
#define R_EPS FLT_EPSILON
#define R_MIN FLT_MIN
#define R_SQRT sqrtf
#define R_ABS fabsf
#include "llzo_tridiag.inc"
#undef R
#undef RSFX
#undef R_EPS
#undef R_MIN
#undef R_SQRT
#undef R_ABS

#define R double
#define RSFX f64
#define R_EPS DBL_EPSILON
#define R_MIN DBL_MIN
#define R_SQRT sqrt
#define R_ABS fabs
#include "llzo_tridiag.inc"
#undef R
#undef RSFX
#undef R_EPS
#undef R_MIN
#undef R_SQRT
#undef R_ABS

/* ---- float ---- */
#define T float
#define R float
#define SFX f32
#define RSFX f32
#define CONJ(x) (x)
#define RE(x) (x)
#define R_EPS FLT_EPSILON
#define R_SQRT sqrtf
#define R_ABS fabsf
#define T_ABS fabsf
#define T_EXP expf
#include "llzo_impl.inc"
#undef T
#undef R
#undef SFX
#undef RSFX
#undef CONJ
#undef RE
#undef R_EPS
#undef R_SQRT
#undef R_ABS
#undef T_ABS
#undef T_EXP

/* ---- double ---- */
#define T double
#define R double
#define SFX f64
#define RSFX f64
#define CONJ(x) (x)
#define RE(x) (x)
#define R_EPS DBL_EPSILON
#define R_SQRT sqrt
#define R_ABS fabs
#define T_ABS fabs
#define T_EXP exp
#include "llzo_impl.inc"
#undef T
#undef R
#undef SFX
#undef RSFX
#undef CONJ
#undef RE
#undef R_EPS
#undef R_SQRT
#undef R_ABS
#undef T_ABS
#undef T_EXP

/* ---- complex double ---- */
#define T double _Complex
#define R double
#define SFX c128
#define RSFX f64
#define CONJ(x) conj(x)
#define RE(x) creal(x)
#define R_EPS DBL_EPSILON
#define R_SQRT sqrt
#define R_ABS fabs
#define T_ABS cabs
#define T_EXP cexp
#include "llzo_impl.inc"
#undef T
#undef R
#undef SFX
#undef RSFX
#undef CONJ
#undef RE
#undef R_EPS
#undef R_SQRT
#undef R_ABS
#undef T_ABS
#undef T_EXP

/* ---- complex float (the reference template covers it: lambda_lanczos.hpp:109, util/common.hpp:80-102) ---- */
#define T float _Complex
#define R float
#define SFX c64
#define RSFX f32
#define CONJ(x) conjf(x)
#define RE(x) crealf(x)
#define R_EPS FLT_EPSILON
#define R_SQRT sqrtf
#define R_ABS fabsf
#define T_ABS cabsf
#define T_EXP cexpf
#include "llzo_impl.inc"
