/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the POGS graph-form hot path.
 *
 * A plain-C restatement of the reference algorithm (foges/pogs @ 649ba26,
 * /root/reference): prox library (src/include/prox_lib.h, prox_tools.h),
 * equilibration + norm estimate (src/cpu/matrix/matrix_dense.cpp,
 * matrix_sparse.cpp, src/cpu/include/equil_helper.h, gsl/gsl_rand.h), direct
 * and CGLS projectors (src/cpu/projector/ sources, src/cpu/include/cgls.h,
 * gsl/gsl_linalg.h) and the ADMM loop (src/cpu/pogs.cpp:91-581).  The
 * reference's BLAS calls (system CBLAS, un-vendored) are restated as loops.
 *
 * Parity: pinned (see pogs_oracle_t.h header and tests/test_oracle.py).
 * The product never links or calls this file.
 *
 * Build: make -C oracle   ->  oracle/liboracle.so
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* LambertW(exp(x)), always evaluated in double: src/include/prox_tools.h:98-129 */
static double lambert_w_exp(double x) {
  double w;
  if (x > 100.0) {
    double lx = log(x);
    return -0.36962844 + x - 0.97284858 * lx + 1.3437973 / lx;
  } else if (x < 0.0) {
    double p = sqrt(2.0 * (exp(x + 1.0) + 1.0));
    w = -1.0 + p * (1.0 + p * (-1.0 / 3.0 + p * (11.0 / 72.0)));
  } else {
    w = x;
  }
  if (x > 1.098612288668110) w -= log(w);
  for (unsigned i = 0; i < 10u; i++) {
    double e = exp(w);
    double t = w * e - exp(x);
    double p = w + 1.0;
    t /= e * p - 0.5 * (p + 1.0) * t / p;
    w -= t;
    if (fabs(t) < 4e-16 * (1.0 + fabs(w))) break;
  }
  return w;
}
double oracle_lambert_w_exp(double x) { return lambert_w_exp(x); }

/* ---- float instantiation ---- */
#define T float
#define SFX s
#define T_FMAX fmaxf
#define T_FMIN fminf
#define T_FABS fabsf
#define T_SQRT sqrtf
#define T_EXP expf
#define T_LOG logf
#define T_POW powf
#define T_ACOS acosf
#define T_COS cosf
#define T_TOL 1e-5f
#define T_EPS FLT_EPSILON
#define T_MAXVAL FLT_MAX
#define T_RAND_K 1
#define T_NEXT_BELOW_ONE nextafterf(1.0f, 0.0f)
#include "pogs_oracle_t.h"
#undef T
#undef SFX
#undef T_FMAX
#undef T_FMIN
#undef T_FABS
#undef T_SQRT
#undef T_EXP
#undef T_LOG
#undef T_POW
#undef T_ACOS
#undef T_COS
#undef T_TOL
#undef T_EPS
#undef T_MAXVAL
#undef T_RAND_K
#undef T_NEXT_BELOW_ONE

/* ---- double instantiation ---- */
#define T double
#define SFX d
#define T_FMAX fmax
#define T_FMIN fmin
#define T_FABS fabs
#define T_SQRT sqrt
#define T_EXP exp
#define T_LOG log
#define T_POW pow
#define T_ACOS acos
#define T_COS cos
#define T_TOL 1e-10
#define T_EPS DBL_EPSILON
#define T_MAXVAL DBL_MAX
#define T_RAND_K 2
#define T_NEXT_BELOW_ONE nextafter(1.0, 0.0)
#include "pogs_oracle_t.h"
