// Scalar proximal-operator library of the graph-form solver, device side.
//
// Semantics follow the reference's FunctionObj contract
//   phi(v) = c*h(a*v - b) + d*v + (e/2)*v^2
// (/root/reference/src/include/prox_lib.h:23-70 for the descriptor and enum
// values, :83-230 for the proximal maps, :241-349 for the objective terms,
// src/include/prox_tools.h:98-149 for LambertWExp / the cubic root).  The code
// is written for a SIMT lane: the sixteen cases are evaluated through one
// switch on a per-element tag that is uniform in practice (a whole f or g
// vector normally carries a single h), all helpers are branch-light, and the
// Lambert-W evaluations run in fp64 exactly like the reference does.
//
// POGS_HD lets the same source be compiled by g++ for the CPU-side unit test
// of this header (tests/test_prox_header.py); the product only uses it from
// CUDA kernels.
#pragma once

#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define POGS_HD __host__ __device__ __forceinline__
#else
#define POGS_HD inline
#endif

namespace pogs_b200 {

// ABI values of the function tag (== reference enum Function / C enum FUNCTION,
// pinned by the reference's tests/test_c_interface.cpp:149-154).
enum FuncTag : int {
  kAbs = 0, kExp = 1, kHuber = 2, kIdentity = 3, kIndBox01 = 4, kIndEq0 = 5, kIndGe0 = 6,
  kIndLe0 = 7, kLogistic = 8, kMaxNeg0 = 9, kMaxPos0 = 10, kNegEntr = 11, kNegLog = 12,
  kRecipr = 13, kSquare = 14, kZero = 15
};

// ---- precision-dispatched math ------------------------------------------
POGS_HD float  m_exp(float x)   { return expf(x); }
POGS_HD double m_exp(double x)  { return exp(x); }
POGS_HD float  m_log(float x)   { return logf(x); }
POGS_HD double m_log(double x)  { return log(x); }
POGS_HD float  m_sqrt(float x)  { return sqrtf(x); }
POGS_HD double m_sqrt(double x) { return sqrt(x); }
POGS_HD float  m_abs(float x)   { return fabsf(x); }
POGS_HD double m_abs(double x)  { return fabs(x); }
POGS_HD float  m_max(float x, float y)   { return fmaxf(x, y); }
POGS_HD double m_max(double x, double y) { return fmax(x, y); }
POGS_HD float  m_min(float x, float y)   { return fminf(x, y); }
POGS_HD double m_min(double x, double y) { return fmin(x, y); }
POGS_HD float  m_pow(float x, float y)   { return powf(x, y); }
POGS_HD double m_pow(double x, double y) { return pow(x, y); }
POGS_HD float  m_acos(float x)  { return acosf(x); }
POGS_HD double m_acos(double x) { return acos(x); }
POGS_HD float  m_cos(float x)   { return cosf(x); }
POGS_HD double m_cos(double x)  { return cos(x); }

template <typename T> POGS_HD T bisect_tol();
template <> POGS_HD float  bisect_tol<float>()  { return 1e-5f; }   // prox_tools.h:57-62
template <> POGS_HD double bisect_tol<double>() { return 1e-10; }

// W(exp(x)) on the principal branch, fp64 (prox_tools.h:98-129): series start,
// then up to ten Halley steps.
POGS_HD double lambert_w_exp(double x) {
  double w;
  if (x > 100.0) {
    const double lx = log(x);
    return -0.36962844 + x - 0.97284858 * lx + 1.3437973 / lx;
  }
  if (x < 0.0) {
    const double p = sqrt(2.0 * (exp(x + 1.0) + 1.0));
    w = -1.0 + p * (1.0 + p * (-1.0 / 3.0 + p * (11.0 / 72.0)));
  } else {
    w = x;
  }
  if (x > 1.098612288668110) w -= log(w);
  const double ex = exp(x);
  for (int it = 0; it < 10; ++it) {
    const double ew = exp(w);
    double t = w * ew - ex;
    const double p = w + 1.0;
    t /= ew * p - 0.5 * (p + 1.0) * t / p;
    w -= t;
    if (fabs(t) < 4e-16 * (1.0 + fabs(w))) break;
  }
  return w;
}

// Single positive root of x^3 + p x^2 + q x + r (prox_tools.h:134-149).
template <typename T>
POGS_HD T cubic_pos_root(T p, T q, T r) {
  const T s = p / 3, s2 = s * s, s3 = s2 * s;
  const T a = -s2 + q / 3;
  const T b = s3 - s * q / 2 + r / 2;
  const T a3 = a * a * a, b2 = b * b;
  if (a3 + b2 >= 0) {
    const T A = m_pow(m_sqrt(a3 + b2) - b, T(1) / 3);
    return -s - a / A + A;
  }
  const T A = m_sqrt(-a3);
  const T B = m_acos(-b / A);
  const T C = m_pow(A, T(1) / 3);
  return -s + (C - a / C) * m_cos(B / 3);
}

// prox of log(1+e^x): 5 safeguarded Newton steps, then guarded bisection
// (prox_lib.h:132-170).
template <typename T>
POGS_HD T prox_logistic(T v, T rho) {
  const T inv = 1 / rho;
  T x;
  if (v < T(-2.5)) x = v;
  else if (v > T(2.5) + inv) x = v - inv;
  else x = (rho * v - T(0.5)) / (T(0.2) + rho);
  T lo = v - inv, hi = v;
  for (int i = 0; i < 5; ++i) {
    const T s = 1 / (1 + m_exp(-x));
    const T fv = s + rho * (x - v);
    const T gv = s * (1 - s) + rho;
    if (fv < 0) lo = x; else hi = x;
    x = x - fv / gv;
    x = m_min(x, hi);
    x = m_max(x, lo);
  }
  for (int i = 0; hi - lo > bisect_tol<T>() && i < 100; ++i) {
    const T gr = 1 / (rho * (1 + m_exp(-x))) + (x - v);
    if (gr > 0) { lo = m_max(lo, x - gr); hi = x; }
    else        { hi = m_min(hi, x - gr); lo = x; }
    x = (hi + lo) / 2;
  }
  return x;
}

// Base proximal map of h with penalty rho (prox_lib.h:83-203).
template <typename T>
POGS_HD T prox_base(int h, T v, T rho) {
  const T inv = 1 / rho;
  switch (h) {
    case kAbs:      return m_max(T(0), v - inv) - m_max(T(0), -(v + inv));
    case kExp:      return v - static_cast<T>(lambert_w_exp(static_cast<double>(v - m_log(rho))));
    case kHuber:    return m_abs(v) < 1 + inv ? v * rho / (1 + rho) : v - (v >= 0 ? T(1) : T(-1)) / rho;
    case kIdentity: return v - inv;
    case kIndBox01: return v <= 0 ? T(0) : (v >= 1 ? T(1) : v);
    case kIndEq0:   return T(0);
    case kIndGe0:   return v <= 0 ? T(0) : v;
    case kIndLe0:   return v >= 0 ? T(0) : v;
    case kLogistic: return prox_logistic(v, rho);
    case kMaxNeg0:  return v + inv <= 0 ? v + inv : (v >= 0 ? v : T(0));
    case kMaxPos0:  return v >= inv ? v - inv : (v <= 0 ? v : T(0));
    case kNegEntr:  return static_cast<T>(lambert_w_exp(static_cast<double>((rho * v - 1) + m_log(rho)))) / rho;
    case kNegLog:   return (v + m_sqrt(v * v + 4 / rho)) / 2;
    case kRecipr:   return cubic_pos_root(-m_max(v, T(0)), T(0), -inv);
    case kSquare:   return rho * v / (1 + rho);
    default:        return v;  // kZero
  }
}

// prox of c*h(a*.-b) + d*. + (e/2)*.^2 with penalty rho (prox_lib.h:207-230).
// c == 0 makes the inner penalty +inf and every base map degenerates correctly.
template <typename T>
POGS_HD T prox_eval(int h, T a, T b, T c, T d, T e, T v, T rho) {
  v = a * (v * rho - d) / (e + rho) - b;
  rho = (e + rho) / (c * a * a);
  v = prox_base(h, v, rho);
  return (v + b) / a;
}

// Objective term c*h(a*x-b) + d*x + e*x^2/2 (prox_lib.h:241-349).
template <typename T>
POGS_HD T func_eval(int h, T a, T b, T c, T d, T e, T x) {
  const T lin = d * x;
  const T quad = e * x * x / 2;
  x = a * x - b;
  T r;
  switch (h) {
    case kAbs:      r = m_abs(x); break;
    case kExp:      r = m_exp(x); break;
    case kHuber:    { const T xa = m_abs(x); r = xa < T(1) ? xa * xa / 2 : xa - T(0.5); break; }
    case kIdentity: r = x; break;
    case kLogistic: r = m_log(1 + m_exp(x)); break;
    case kMaxNeg0:  r = m_max(T(0), -x); break;
    case kMaxPos0:  r = m_max(T(0), x); break;
    case kNegEntr:  r = x <= 0 ? T(0) : x * m_log(x); break;
    case kNegLog:   r = -m_log(m_max(T(0), x)); break;
    case kRecipr:   r = 1 / m_max(T(0), x); break;
    case kSquare:   r = x * x / 2; break;
    default:        r = 0; break;  // indicators and kZero
  }
  return c * r + lin + quad;
}

}  // namespace pogs_b200
