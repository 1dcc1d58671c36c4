/*
 * TEST INFRASTRUCTURE ONLY -- type-generic body of the CPU oracle.
 *
 * Included twice by pogs_oracle.c (once with T=float / SFX=s, once with
 * T=double / SFX=d).  Every function restates, in plain C, the algorithm of the
 * reference file:line named in its comment (foges/pogs @ 649ba26, tree at
 * /root/reference).  Nothing here is linked into, imported by or called from
 * the product (pogs_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg use it.
 *
 * Parity status: pinned.  tests/test_oracle.py checks this restatement against
 * (i) the known-answer prox values of the reference's tests/test_proximal.cpp,
 * (ii) golden vectors produced by the compiled, unmodified reference
 * (oracle/_ref, tests/golden/make_golden.py) and (iii) live runs of oracle/_ref
 * when it is present.
 */

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(CAT(name, _), SFX)

/* ------------------------------------------------------------------------ */
/* Scalar helpers: src/include/prox_tools.h:12-88                            */
/* ------------------------------------------------------------------------ */
static inline T FN(maxpos)(T x) { return T_FMAX((T)0, x); }
static inline T FN(maxneg)(T x) { return T_FMAX((T)0, -x); }

/* Base proximal operators: src/include/prox_lib.h:83-203.  rho here is the
 * already-transformed penalty (e+rho)/(c a^2) of the wrapper below. */
static T FN(prox_base)(int h, T v, T rho) {
  switch (h) {
    case 0: /* kAbs  :83  */ return FN(maxpos)(v - 1 / rho) - FN(maxneg)(v + 1 / rho);
    case 1: /* kExp  :96  */ return v - (T)lambert_w_exp((double)(v - T_LOG(rho)));
    case 2: /* kHuber:102 */
      return T_FABS(v) < 1 + 1 / rho ? v * rho / (1 + rho) : v - (v >= 0 ? (T)1 : (T)-1) / rho;
    case 3: /* kIdentity :107 */ return v - 1 / rho;
    case 4: /* kIndBox01 :112 */ return v <= 0 ? (T)0 : (v >= 1 ? (T)1 : v);
    case 5: /* kIndEq0   :117 */ return (T)0;
    case 6: /* kIndGe0   :122 */ return v <= 0 ? (T)0 : v;
    case 7: /* kIndLe0   :127 */ return v >= 0 ? (T)0 : v;
    case 8: { /* kLogistic :132-170 */
      T x;
      if (v < (T)-2.5) x = v;
      else if (v > (T)2.5 + 1 / rho) x = v - 1 / rho;
      else x = (rho * v - (T)0.5) / ((T)0.2 + rho);
      T lo = v - 1 / rho, hi = v;
      for (unsigned i = 0; i < 5; ++i) {
        T s = 1 / (1 + T_EXP(-x));
        T fv = s + rho * (x - v);
        T gv = s * (1 - s) + rho;
        if (fv < 0) lo = x; else hi = x;
        x = x - fv / gv;
        x = T_FMIN(x, hi);
        x = T_FMAX(x, lo);
      }
      for (unsigned i = 0; hi - lo > T_TOL && i < 100; ++i) {
        T gr = 1 / (rho * (1 + T_EXP(-x))) + (x - v);
        if (gr > 0) { lo = T_FMAX(lo, x - gr); hi = x; }
        else        { hi = T_FMIN(hi, x - gr); lo = x; }
        x = (hi + lo) / 2;
      }
      return x;
    }
    case 9: { /* kMaxNeg0 :173 */
      T z = v >= 0 ? v : (T)0;
      return v + 1 / rho <= 0 ? v + 1 / rho : z;
    }
    case 10: { /* kMaxPos0 :179 */
      T z = v <= 0 ? v : (T)0;
      return v >= 1 / rho ? v - 1 / rho : z;
    }
    case 11: /* kNegEntr :88 */
      return (T)lambert_w_exp((double)((rho * v - 1) + T_LOG(rho))) / rho;
    case 12: /* kNegLog :185 */ return (v + T_SQRT(v * v + 4 / rho)) / 2;
    case 13: { /* kRecipr :190 + CubicSolve prox_tools.h:134-149 with p=-max(v,0), q=0, r=-1/rho */
      T p = -T_FMAX(v, (T)0), q = 0, r = -1 / rho;
      T s = p / 3, s2 = s * s, s3 = s2 * s;
      T a = -s2 + q / 3;
      T b = s3 - s * q / 2 + r / 2;
      T a3 = a * a * a, b2 = b * b;
      if (a3 + b2 >= 0) {
        T A = T_POW(T_SQRT(a3 + b2) - b, (T)1 / 3);
        return -s - a / A + A;
      } else {
        T A = T_SQRT(-a3);
        T B = T_ACOS(-b / A);
        T C = T_POW(A, (T)1 / 3);
        return -s + (C - a / C) * T_COS(B / 3);
      }
    }
    case 14: /* kSquare :196 */ return rho * v / (1 + rho);
    default: /* kZero :201 */ return v;
  }
}

/* Wrapper c*h(a*x-b)+d*x+e*x^2/2: src/include/prox_lib.h:207-230 */
static T FN(prox_eval)(int h, T a, T b, T c, T d, T e, T v, T rho) {
  v = a * (v * rho - d) / (e + rho) - b;
  rho = (e + rho) / (c * a * a);
  v = FN(prox_base)(h, v, rho);
  return (v + b) / a;
}

/* Objective term: src/include/prox_lib.h:241-349 */
static T FN(func_eval)(int h, T a, T b, T c, T d, T e, T x) {
  T dx = d * x;
  T ex = e * x * x / 2;
  x = a * x - b;
  T r;
  switch (h) {
    case 0: r = T_FABS(x); break;
    case 1: r = T_EXP(x); break;
    case 2: { T xa = T_FABS(x); r = xa < (T)1 ? xa * xa / 2 : xa - (T)0.5; break; }
    case 3: r = x; break;
    case 8: r = T_LOG(1 + T_EXP(x)); break;
    case 9: r = FN(maxneg)(x); break;
    case 10: r = FN(maxpos)(x); break;
    case 11: r = x <= 0 ? (T)0 : x * T_LOG(x); break;
    case 12: x = T_FMAX((T)0, x); r = -T_LOG(x); break;
    case 13: x = T_FMAX((T)0, x); r = 1 / x; break;
    case 14: r = x * x / 2; break;
    default: r = 0; break; /* indicators + kZero */
  }
  return c * r + dx + ex;
}

/* Exposed for unit tests: base prox on raw (v, rho) -- the calls made by the
 * reference's tests/test_proximal.cpp. */
T FN(oracle_prox_base)(int h, T v, T rho) { return FN(prox_base)(h, v, rho); }

/* Vector ProxEval / FuncEval: src/include/prox_lib.h:504-511, 521-529 */
void FN(oracle_prox_vec)(size_t n, const int *h, const T *a, const T *b, const T *c, const T *d,
                         const T *e, T rho, const T *in, T *out) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i)
    out[i] = FN(prox_eval)(h[i], a[i], b[i], c[i], d[i], e[i], in[i], rho);
}

T FN(oracle_func_vec)(size_t n, const int *h, const T *a, const T *b, const T *c, const T *d,
                      const T *e, const T *in) {
  T sum = 0;
  for (size_t i = 0; i < n; ++i) sum += FN(func_eval)(h[i], a[i], b[i], c[i], d[i], e[i], in[i]);
  return sum;
}

/* ------------------------------------------------------------------------ */
/* BLAS-1 restatements (the reference calls CBLAS: gsl_blas.h:17-87)         */
/* ------------------------------------------------------------------------ */
static T FN(nrm2)(size_t n, const T *x) {
  double s = 0; /* wide accumulator: stands in for OpenBLAS' scaled/extended nrm2 */
  for (size_t i = 0; i < n; ++i) s += (double)x[i] * (double)x[i];
  return (T)sqrt(s);
}
static T FN(dot)(size_t n, const T *x, const T *y) {
  T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  size_t i = 0;
  for (; i + 4 <= n; i += 4) {
    s0 += x[i] * y[i]; s1 += x[i + 1] * y[i + 1]; s2 += x[i + 2] * y[i + 2]; s3 += x[i + 3] * y[i + 3];
  }
  for (; i < n; ++i) s0 += x[i] * y[i];
  return (s0 + s1) + (s2 + s3);
}
static void FN(axpy)(size_t n, T a, const T *x, T *y) { for (size_t i = 0; i < n; ++i) y[i] += a * x[i]; }
static void FN(scal)(size_t n, T a, T *x) { for (size_t i = 0; i < n; ++i) x[i] *= a; }

/* ------------------------------------------------------------------------ */
/* Matrix operator: dense (matrix_dense.cpp:76-113) or CSR+CSC sparse        */
/* (matrix_sparse.cpp:97-155, gsl_spblas.h:10-40, gsl_spmat.h:32-58)          */
/* ------------------------------------------------------------------------ */
typedef struct {
  int sparse;     /* 0 dense, 1 sparse */
  int rowmaj;     /* dense: 1 = row-major */
  size_t m, n, nnz;
  T *data;        /* dense: m*n copy.  sparse: 2*nnz (CSR copy then CSC copy) */
  int *ind;       /* sparse: 2*nnz */
  int *ptr;       /* sparse: (m+1) + (n+1) */
} FN(mat);

/* y = alpha*op(A)x + beta*y ; `sq` squares the entries on the fly (used by the
 * equilibration, which the reference does on a squared copy of A). */
static void FN(mat_mul)(const FN(mat) *A, char op, T alpha, const T *x, T beta, T *y, int sq) {
  size_t m = A->m, n = A->n;
  if (!A->sparse) {
    int row_dot = (A->rowmaj && op == 'n') || (!A->rowmaj && op == 't');
    size_t rows = op == 'n' ? m : n;   /* output length */
    size_t cols = op == 'n' ? n : m;   /* input length */
    if (row_dot) {
      /* output i = dot(contiguous line i, x) */
#pragma omp parallel for schedule(static)
      for (long i = 0; i < (long)rows; ++i) {
        const T *r = A->data + (size_t)i * cols;
        T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        size_t j = 0;
        if (sq) {
          for (; j + 4 <= cols; j += 4) {
            s0 += r[j] * r[j] * x[j]; s1 += r[j + 1] * r[j + 1] * x[j + 1];
            s2 += r[j + 2] * r[j + 2] * x[j + 2]; s3 += r[j + 3] * r[j + 3] * x[j + 3];
          }
          for (; j < cols; ++j) s0 += r[j] * r[j] * x[j];
        } else {
          for (; j + 4 <= cols; j += 4) {
            s0 += r[j] * x[j]; s1 += r[j + 1] * x[j + 1]; s2 += r[j + 2] * x[j + 2]; s3 += r[j + 3] * x[j + 3];
          }
          for (; j < cols; ++j) s0 += r[j] * x[j];
        }
        T s = (s0 + s1) + (s2 + s3);
        y[i] = beta == 0 ? alpha * s : alpha * s + beta * y[i];
      }
    } else {
      /* output = sum over contiguous lines l of x[l] * line_l (axpy form),
       * accumulated in blocks of lines to keep fp32 sums short. */
      size_t lines = cols, len = rows;
      T *acc = (T *)calloc(len, sizeof(T));
      enum { BLK = 256 };
      T *part = (T *)malloc(len * sizeof(T));
      for (size_t l0 = 0; l0 < lines; l0 += BLK) {
        size_t l1 = l0 + BLK < lines ? l0 + BLK : lines;
        memset(part, 0, len * sizeof(T));
        for (size_t l = l0; l < l1; ++l) {
          const T *r = A->data + l * len;
          T xl = x[l];
          if (sq) { for (size_t j = 0; j < len; ++j) part[j] += r[j] * r[j] * xl; }
          else    { for (size_t j = 0; j < len; ++j) part[j] += r[j] * xl; }
        }
        for (size_t j = 0; j < len; ++j) acc[j] += part[j];
      }
      for (size_t j = 0; j < len; ++j) y[j] = beta == 0 ? alpha * acc[j] : alpha * acc[j] + beta * y[j];
      free(acc); free(part);
    }
  } else {
    /* pick whichever compressed copy makes the product a row-gather */
    int first = (A->rowmaj && op == 'n') || (!A->rowmaj && op == 't');
    size_t len0 = A->rowmaj ? m + 1 : n + 1;
    const T *val = first ? A->data : A->data + A->nnz;
    const int *ind = first ? A->ind : A->ind + A->nnz;
    const int *ptr = first ? A->ptr : A->ptr + len0;
    size_t rows = op == 'n' ? m : n;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)rows; ++i) {
      T s = 0;
      if (sq) { for (int j = ptr[i]; j < ptr[i + 1]; ++j) s += val[j] * val[j] * x[ind[j]]; }
      else    { for (int j = ptr[i]; j < ptr[i + 1]; ++j) s += val[j] * x[ind[j]]; }
      y[i] = alpha * s + beta * y[i];
    }
  }
}

/* CSR -> CSC (counting sort), gsl_spmat.h:32-58 */
static void FN(transpose_compressed)(size_t rows, size_t cols, size_t nnz, const T *a, const int *rp,
                                     const int *ci, T *at, int *ri, int *cp) {
  memset(cp, 0, (cols + 1) * sizeof(int));
  for (size_t k = 0; k < nnz; ++k) cp[ci[k] + 1]++;
  for (size_t j = 0; j < cols; ++j) cp[j + 1] += cp[j];
  int *next = (int *)malloc((cols + 1) * sizeof(int));
  memcpy(next, cp, (cols + 1) * sizeof(int));
  for (size_t i = 0; i < rows; ++i)
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
      int dst = next[ci[k]]++;
      ri[dst] = (int)i;
      at[dst] = a[k];
    }
  free(next);
}

static void FN(mat_free)(FN(mat) *A) { free(A->data); free(A->ind); free(A->ptr); }

static void FN(mat_init_dense)(FN(mat) *A, int rowmaj, size_t m, size_t n, const T *src) {
  memset(A, 0, sizeof(*A));
  A->rowmaj = rowmaj; A->m = m; A->n = n;
  A->data = (T *)malloc(m * n * sizeof(T));
  memcpy(A->data, src, m * n * sizeof(T));
}

static void FN(mat_init_sparse)(FN(mat) *A, int rowmaj, size_t m, size_t n, size_t nnz, const T *val,
                                const int *ptr, const int *ind) {
  memset(A, 0, sizeof(*A));
  A->sparse = 1; A->rowmaj = rowmaj; A->m = m; A->n = n; A->nnz = nnz;
  size_t len0 = rowmaj ? m + 1 : n + 1;
  A->data = (T *)malloc(2 * nnz * sizeof(T));
  A->ind = (int *)malloc(2 * nnz * sizeof(int));
  A->ptr = (int *)malloc((m + n + 2) * sizeof(int));
  memcpy(A->data, val, nnz * sizeof(T));
  memcpy(A->ind, ind, nnz * sizeof(int));
  memcpy(A->ptr, ptr, len0 * sizeof(int));
  if (rowmaj) FN(transpose_compressed)(m, n, nnz, val, ptr, ind, A->data + nnz, A->ind + nnz, A->ptr + len0);
  else        FN(transpose_compressed)(n, m, nnz, val, ptr, ind, A->data + nnz, A->ind + nnz, A->ptr + len0);
}

/* ------------------------------------------------------------------------ */
/* Equilibration: matrix_dense.cpp:116-200, matrix_sparse.cpp:158-242,        */
/* equil_helper.h:141-164 (modified Sinkhorn-Knopp on A.^2, 50 sweeps)        */
/* ------------------------------------------------------------------------ */
static void FN(equilibrate)(FN(mat) *A, T *d, T *e) {
  size_t m = A->m, n = A->n;
  for (size_t i = 0; i < m; ++i) d[i] = 1;
  for (size_t j = 0; j < n; ++j) e[j] = 1;
  T ce = (T)1e-4 * (T)(m + n) / (T)m;
  T cd = (T)1e-4 * (T)(m + n) / (T)n;
  for (unsigned k = 0; k < 50; ++k) {
    FN(mat_mul)(A, 't', 1, d, 0, e, 1);
    for (size_t j = 0; j < n; ++j) e[j] = (T)m / (e[j] + ce);
    FN(mat_mul)(A, 'n', 1, e, 0, d, 1);
    for (size_t i = 0; i < m; ++i) d[i] = (T)n / (d[i] + cd);
  }
  for (size_t i = 0; i < m; ++i) d[i] = T_SQRT(d[i]);
  for (size_t j = 0; j < n; ++j) e[j] = T_SQRT(e[j]);
  /* A := D A E, then Frobenius-normalise (norm over one copy for sparse). */
  double fro = 0;
  if (!A->sparse) {
    for (size_t i = 0; i < m; ++i)
      for (size_t j = 0; j < n; ++j) {
        size_t t = A->rowmaj ? i * n + j : j * m + i;
        A->data[t] *= d[i] * e[j];
        fro += (double)A->data[t] * (double)A->data[t];
      }
  } else {
    size_t len0 = A->rowmaj ? m + 1 : n + 1;
    for (int copy = 0; copy < 2; ++copy) {
      int is_csr = (A->rowmaj != 0) == (copy == 0);
      T *val = A->data + (copy ? A->nnz : 0);
      const int *ind = A->ind + (copy ? A->nnz : 0);
      const int *ptr = A->ptr + (copy ? len0 : 0);
      size_t lines = is_csr ? m : n;
      for (size_t t = 0; t < lines; ++t)
        for (int k = ptr[t]; k < ptr[t + 1]; ++k) {
          val[k] *= is_csr ? d[t] * e[ind[k]] : d[ind[k]] * e[t];
          if (copy == 0) fro += (double)val[k] * (double)val[k];
        }
    }
  }
  size_t mn = m < n ? m : n;
  T normA = (T)(sqrt(fro) / sqrt((double)mn));
  size_t tot = A->sparse ? 2 * A->nnz : m * n;
  T inv = 1 / normA;
  for (size_t t = 0; t < tot; ++t) A->data[t] *= inv;
  T invs = 1 / T_SQRT(normA);
  for (size_t i = 0; i < m; ++i) d[i] *= invs;
  for (size_t j = 0; j < n; ++j) e[j] *= invs;
}

/* Start vector of the norm estimate: gsl_rand.h:9-16 =
 * std::default_random_engine (libstdc++: minstd_rand0, seed 1) driving
 * std::uniform_real_distribution<T>(0,1) (libstdc++ generate_canonical). */
void FN(oracle_rand)(T *x, size_t size) {
  unsigned long long st = 1;
  const long double R = 2147483646.0L; /* max - min + 1 */
  int k = T_RAND_K;                    /* draws per variate: 1 (float) / 2 (double) */
  for (size_t i = 0; i < size; ++i) {
    T sum = 0, tmp = 1;
    for (int j = 0; j < k; ++j) {
      st = (st * 16807ULL) % 2147483647ULL;
      sum += (T)(st - 1ULL) * tmp;
      tmp *= (T)R;
    }
    T r = sum / tmp;
    if (r >= (T)1) r = T_NEXT_BELOW_ONE;
    x[i] = r;
  }
}

/* Norm2Est: equil_helper.h:108-135 */
static T FN(norm2est)(const FN(mat) *A) {
  size_t m = A->m, n = A->n;
  T *x = (T *)calloc(n, sizeof(T)), *Sx = (T *)calloc(m, sizeof(T));
  FN(oracle_rand)(x, n);
  T est = 0, last;
  for (unsigned i = 0; i < 50; ++i) {
    last = est;
    FN(mat_mul)(A, 'n', 1, x, 0, Sx, 0);
    FN(mat_mul)(A, 't', 1, Sx, 0, x, 0);
    T nx = FN(nrm2)(n, x), nSx = FN(nrm2)(m, Sx);
    FN(scal)(n, 1 / nx, x);
    est = nx / nSx;
    if (T_FABS(last - est) < (T)1e-4 * est) break;
  }
  free(x); free(Sx);
  return est;
}

/* ------------------------------------------------------------------------ */
/* Direct projector: projector_direct_dense.cpp:45-84 (Gram matrix),          */
/* :116-135 (factor once, then gemv / 2 x trsv / gemv); Cholesky + solve      */
/* gsl_linalg.h:14-61.                                                        */
/* ------------------------------------------------------------------------ */
typedef struct { size_t k; T *L; int factored; } FN(direct);

static void FN(direct_init)(FN(direct) *P, const FN(mat) *A) {
  size_t m = A->m, n = A->n, k = m < n ? m : n;
  P->k = k; P->factored = 0;
  P->L = (T *)calloc(k * k, sizeof(T));
  /* lower triangle of A^T A (m>n) or A A^T, blocked over the long dimension */
  int gram_cols = m > n;
  size_t inner = gram_cols ? m : n;
#pragma omp parallel for schedule(dynamic, 8)
  for (long i = 0; i < (long)k; ++i)
    for (size_t j = 0; j <= (size_t)i; ++j) {
      double s = 0;
      for (size_t t = 0; t < inner; ++t) {
        size_t r1, c1, r2, c2;
        if (gram_cols) { r1 = t; c1 = i; r2 = t; c2 = j; } else { r1 = i; c1 = t; r2 = j; c2 = t; }
        T a1 = A->rowmaj ? A->data[r1 * n + c1] : A->data[c1 * m + r1];
        T a2 = A->rowmaj ? A->data[r2 * n + c2] : A->data[c2 * m + r2];
        s += (double)a1 * (double)a2;
      }
      P->L[(size_t)i * k + j] = (T)s;
    }
}

static void FN(direct_factor)(FN(direct) *P, T s) {
  size_t k = P->k; T *L = P->L;
  for (size_t i = 0; i < k; ++i) L[i * k + i] += s;
  /* row-oriented Cholesky, lower triangle in place */
  for (size_t i = 0; i < k; ++i) {
    for (size_t j = 0; j <= i; ++j) {
      T sum = L[i * k + j];
      const T *ri = L + i * k, *rj = L + j * k;
      for (size_t t = 0; t < j; ++t) sum -= ri[t] * rj[t];
      L[i * k + j] = (i == j) ? T_SQRT(sum) : sum / L[j * k + j];
    }
  }
  P->factored = 1;
}

static void FN(chol_solve)(const FN(direct) *P, T *v) {
  size_t k = P->k; const T *L = P->L;
  for (size_t i = 0; i < k; ++i) {           /* L w = v */
    T s = v[i];
    for (size_t t = 0; t < i; ++t) s -= L[i * k + t] * v[t];
    v[i] = s / L[i * k + i];
  }
  for (size_t ii = k; ii-- > 0;) {           /* L^T u = w */
    T s = v[ii] / L[ii * k + ii];
    v[ii] = s;
    for (size_t t = 0; t < ii; ++t) v[t] -= L[ii * k + t] * s;
  }
}

static void FN(project_direct)(FN(direct) *P, const FN(mat) *A, const T *x0, const T *y0, T s, T *x, T *y) {
  size_t m = A->m, n = A->n;
  if (!P->factored) FN(direct_factor)(P, s);
  memcpy(x, x0, n * sizeof(T));
  memcpy(y, y0, m * sizeof(T));
  if (m > n) {
    FN(mat_mul)(A, 't', 1, y, 1, x, 0);
    FN(chol_solve)(P, x);
    FN(mat_mul)(A, 'n', 1, x, 0, y, 0);
  } else {
    FN(mat_mul)(A, 'n', 1, x, -1, y, 0);
    FN(chol_solve)(P, y);
    FN(mat_mul)(A, 't', -1, y, 1, x, 0);
    FN(axpy)(m, 1, y0, y);
  }
}

/* ------------------------------------------------------------------------ */
/* CGLS projector: projector_cgls.cpp:52-88, cgls.h:201-323 (scalars double)  */
/* ------------------------------------------------------------------------ */
static int FN(cgls)(const FN(mat) *A, const T *b, T *x, double shift, double tol, int maxit, int *iters) {
  size_t m = A->m, n = A->n;
  T *p = (T *)calloc(n, sizeof(T)), *q = (T *)calloc(m, sizeof(T));
  T *r = (T *)malloc(m * sizeof(T)), *s = (T *)malloc(n * sizeof(T));
  memcpy(r, b, m * sizeof(T));
  memcpy(s, x, n * sizeof(T));
  const double eps = T_EPS;
  int flag = 0, indefinite = 0, k = 0;
  double normx = FN(nrm2)(n, x);
  if (normx > 0.) FN(mat_mul)(A, 'n', -1, x, 1, r, 0);
  FN(mat_mul)(A, 't', 1, r, (T)(-shift), s, 0);
  memcpy(p, s, n * sizeof(T));
  double norms = FN(nrm2)(n, s), norms0 = norms, gamma = norms0 * norms0;
  normx = FN(nrm2)(n, x);
  double xmax = normx;
  if (norms < eps) flag = 1;
  for (k = 0; k < maxit && !flag; ++k) {
    FN(mat_mul)(A, 'n', 1, p, 0, q, 0);
    double normp = FN(nrm2)(n, p), normq = FN(nrm2)(m, q);
    double delta = normq * normq + shift * normp * normp;
    if (delta <= 0.) indefinite = 1;
    if (delta == 0.) delta = eps;
    T alpha = (T)(gamma / delta), nalpha = (T)(-gamma / delta);
    FN(axpy)(n, alpha, p, x);
    FN(axpy)(m, nalpha, q, r);
    memcpy(s, x, n * sizeof(T));
    FN(mat_mul)(A, 't', 1, r, (T)(-shift), s, 0);
    norms = FN(nrm2)(n, s);
    double gamma1 = gamma;
    gamma = norms * norms;
    T beta = (T)(gamma / gamma1);
    FN(axpy)(n, beta, p, s);
    memcpy(p, s, n * sizeof(T));
    normx = FN(nrm2)(n, x);
    if (normx > xmax) xmax = normx;
    if ((norms <= norms0 * tol) || (normx * tol >= 1.)) break;
  }
  if (iters) *iters = k;
  double shrink = normx / xmax;
  if (k == maxit) flag = 2; else if (indefinite) flag = 3; else if (shrink * shrink <= tol) flag = 4;
  free(p); free(q); free(r); free(s);
  return flag;
}

static void FN(project_cgls)(const FN(mat) *A, const T *x0, const T *y0, T s, T *x, T *y, T tol, long *inner) {
  size_t m = A->m, n = A->n;
  FN(axpy)(n, -1, x0, x);                 /* x holds the warm start -> delta */
  memcpy(y, y0, m * sizeof(T));
  FN(mat_mul)(A, 'n', -1, x0, 1, y, 0);   /* y = y0 - A x0 */
  int it = 0;
  FN(cgls)(A, y, x, (double)s, (double)tol, 500, &it);
  if (inner) *inner += it;
  FN(axpy)(n, 1, x0, x);
  FN(mat_mul)(A, 'n', 1, x, 0, y, 0);
}

/* ------------------------------------------------------------------------ */
/* Persistent solver object == PogsImplementation state, pogs.h:55-75         */
/* ------------------------------------------------------------------------ */
typedef struct {
  FN(mat) A;
  FN(direct) P;
  int direct;          /* 1: ProjectorDirect, 0: ProjectorCgls */
  int done_init;
  T *de, *z, *zt;
  T rho, nrmA;
  T *x, *y, *mu, *lambda;
  T optval;
  unsigned final_iter;
  long cgls_iters;     /* total CGLS inner iterations of the last solve */
  T abs_tol, rel_tol;
  unsigned max_iter;
  int adaptive_rho, gap_stop, init_x, init_lambda;
} FN(solver);

void *FN(oracle_create_dense)(int rowmaj, size_t m, size_t n, const T *A, int direct) {
  FN(solver) *S = (FN(solver) *)calloc(1, sizeof(FN(solver)));
  FN(mat_init_dense)(&S->A, rowmaj, m, n, A);
  S->direct = direct;
  S->rho = 1; S->abs_tol = (T)1e-4; S->rel_tol = (T)1e-3; S->max_iter = 2500;
  S->adaptive_rho = 1; S->gap_stop = 0;
  S->x = (T *)calloc(n, sizeof(T)); S->mu = (T *)calloc(n, sizeof(T));
  S->y = (T *)calloc(m, sizeof(T)); S->lambda = (T *)calloc(m, sizeof(T));
  return S;
}

void *FN(oracle_create_sparse)(int rowmaj, size_t m, size_t n, size_t nnz, const T *val, const int *ptr,
                               const int *ind) {
  FN(solver) *S = (FN(solver) *)calloc(1, sizeof(FN(solver)));
  FN(mat_init_sparse)(&S->A, rowmaj, m, n, nnz, val, ptr, ind);
  S->direct = 0;
  S->rho = 1; S->abs_tol = (T)1e-4; S->rel_tol = (T)1e-3; S->max_iter = 2500;
  S->adaptive_rho = 1; S->gap_stop = 0;
  S->x = (T *)calloc(n, sizeof(T)); S->mu = (T *)calloc(n, sizeof(T));
  S->y = (T *)calloc(m, sizeof(T)); S->lambda = (T *)calloc(m, sizeof(T));
  return S;
}

void FN(oracle_destroy)(void *h) {
  FN(solver) *S = (FN(solver) *)h;
  FN(mat_free)(&S->A);
  free(S->P.L); free(S->de); free(S->z); free(S->zt);
  free(S->x); free(S->y); free(S->mu); free(S->lambda);
  free(S);
}

void FN(oracle_set_params)(void *h, T rho, T abs_tol, T rel_tol, unsigned max_iter, int adaptive_rho,
                           int gap_stop) {
  FN(solver) *S = (FN(solver) *)h;
  S->rho = rho; S->abs_tol = abs_tol; S->rel_tol = rel_tol; S->max_iter = max_iter;
  S->adaptive_rho = adaptive_rho; S->gap_stop = gap_stop;
}
void FN(oracle_set_init)(void *h, const T *x, const T *lambda) {
  FN(solver) *S = (FN(solver) *)h;
  if (x) { memcpy(S->x, x, S->A.n * sizeof(T)); S->init_x = 1; }
  if (lambda) { memcpy(S->lambda, lambda, S->A.m * sizeof(T)); S->init_lambda = 1; }
}

/* _Init: pogs.cpp:59-88 */
static void FN(solver_init)(FN(solver) *S) {
  size_t m = S->A.m, n = S->A.n;
  S->done_init = 1;
  S->de = (T *)calloc(m + n, sizeof(T));
  S->z = (T *)calloc(m + n, sizeof(T));
  S->zt = (T *)calloc(m + n, sizeof(T));
  FN(equilibrate)(&S->A, S->de, S->de + m);
  S->nrmA = FN(norm2est)(&S->A);
  if (S->direct) FN(direct_init)(&S->P, &S->A);
}

/* Run the setup only and export what it produced (unit tests of the device setup). */
void FN(oracle_setup)(void *h, T *d_out, T *e_out, T *nrmA_out, T *Aeq_out) {
  FN(solver) *S = (FN(solver) *)h;
  if (!S->done_init) FN(solver_init)(S);
  size_t m = S->A.m, n = S->A.n;
  if (d_out) memcpy(d_out, S->de, m * sizeof(T));
  if (e_out) memcpy(e_out, S->de + m, n * sizeof(T));
  if (nrmA_out) *nrmA_out = S->nrmA;
  if (Aeq_out && !S->A.sparse) memcpy(Aeq_out, S->A.data, m * n * sizeof(T));
}

/* One projection in the equilibrated space (unit tests of the device projector). */
void FN(oracle_project)(void *h, const T *x0, const T *y0, T *x, T *y, T tol) {
  FN(solver) *S = (FN(solver) *)h;
  if (!S->done_init) FN(solver_init)(S);
  if (S->direct) FN(project_direct)(&S->P, &S->A, x0, y0, 1, x, y);
  else FN(project_cgls)(&S->A, x0, y0, 1, x, y, tol, 0);
}

/* Solve: pogs.cpp:91-581 with the separable objective of :591-621. */
int FN(oracle_solve)(void *h, const int *f_h, const T *f_a, const T *f_b, const T *f_c, const T *f_d,
                     const T *f_e, const int *g_h, const T *g_a, const T *g_b, const T *g_c,
                     const T *g_d, const T *g_e) {
  FN(solver) *S = (FN(solver) *)h;
  const T kDeltaMin = (T)1.05, kGamma = (T)1.01, kTau = (T)0.8, kRhoMin = (T)1e-4, kRhoMax = (T)1e4;
  const T kKappa = (T)0.9, kAlpha = (T)1.7, kProjTolMax = (T)1e-8, kProjTolMin = (T)1e-2;
  if (!S->done_init) FN(solver_init)(S);
  size_t m = S->A.m, n = S->A.n, N = m + n;
  T *d = S->de, *e = S->de + m;
  T *z = S->z, *zt = S->zt;
  T *zprev = (T *)calloc(N, sizeof(T)), *ztemp = (T *)calloc(N, sizeof(T)), *z12 = (T *)calloc(N, sizeof(T));
  T *x = z, *y = z + n, *x12 = z12, *y12 = z12 + n, *xprev = zprev;
  T *xtemp = ztemp, *ytemp = ztemp + n;

  /* clamp c,e >= 0 (FunctionObj ctor, prox_lib.h:62-69) and rescale (pogs.cpp:608-617) */
  T *fa = (T *)malloc(m * sizeof(T)), *fc = (T *)malloc(m * sizeof(T)), *fd = (T *)malloc(m * sizeof(T)),
    *fe = (T *)malloc(m * sizeof(T));
  T *ga = (T *)malloc(n * sizeof(T)), *gc = (T *)malloc(n * sizeof(T)), *gd = (T *)malloc(n * sizeof(T)),
    *ge = (T *)malloc(n * sizeof(T));
  for (size_t i = 0; i < m; ++i) {
    fc[i] = T_FMAX(f_c[i], (T)0);
    fa[i] = f_a[i] / d[i]; fd[i] = f_d[i] / d[i]; fe[i] = T_FMAX(f_e[i], (T)0) / (d[i] * d[i]);
  }
  for (size_t j = 0; j < n; ++j) {
    gc[j] = T_FMAX(g_c[j], (T)0);
    ga[j] = g_a[j] * e[j]; gd[j] = g_d[j] * e[j]; ge[j] = T_FMAX(g_e[j], (T)0) * (e[j] * e[j]);
  }

  /* explicit warm start, pogs.cpp:144-156 (both x and lambda, else the reference aborts) */
  if (S->init_x && S->init_lambda) {
    for (size_t j = 0; j < n; ++j) xtemp[j] = S->x[j] / e[j];
    FN(mat_mul)(&S->A, 'n', 1, xtemp, 0, ytemp, 0);
    memcpy(z, ztemp, N * sizeof(T));
    for (size_t i = 0; i < m; ++i) ytemp[i] = S->lambda[i] / d[i];
    FN(mat_mul)(&S->A, 't', -1, ytemp, 0, xtemp, 0);
    FN(scal)(N, -1 / S->rho, ztemp);
    memcpy(zt, ztemp, N * sizeof(T));
  } else if (S->init_x || S->init_lambda) {
    return 6; /* reference: ASSERT(false) -> exit(1) */
  }
  S->init_x = S->init_lambda = 0;

  T sqrtn_atol = T_SQRT((T)n) * S->abs_tol, sqrtm_atol = T_SQRT((T)m) * S->abs_tol;
  T sqrtmn_atol = T_SQRT((T)(m + n)) * S->abs_tol;
  T delta = kDeltaMin, xi = 1;
  unsigned k = 0, kd = 0, ku = 0;
  int converged = 0;
  T nrm_r = 0, nrm_s = 0, gap, eps_gap, eps_pri, eps_dua;
  T prev_nrm_r = T_MAXVAL;
  S->cgls_iters = 0;

  for (;; ++k) {
    memcpy(zprev, z, N * sizeof(T));
    FN(axpy)(N, -1, zt, z);
    FN(oracle_prox_vec)(n, g_h, ga, g_b, gc, gd, ge, S->rho, x, x12);
    FN(oracle_prox_vec)(m, f_h, fa, f_b, fc, fd, fe, S->rho, y, y12);

    FN(axpy)(N, -1, z12, z);
    gap = T_FABS(FN(dot)(N, z, z12));
    eps_gap = sqrtmn_atol + S->rel_tol * FN(nrm2)(N, z) * FN(nrm2)(N, z12);
    eps_pri = sqrtm_atol + S->rel_tol * FN(nrm2)(m, y12);
    eps_dua = S->rho * (sqrtn_atol + S->rel_tol * FN(nrm2)(n, x));

    memcpy(ztemp, zt, N * sizeof(T));
    FN(axpy)(N, kAlpha, z12, ztemp);
    FN(axpy)(N, 1 - kAlpha, zprev, ztemp);

    memcpy(x, xprev, n * sizeof(T));
    T proj_tol = kProjTolMin * T_POW(T_FMIN(prev_nrm_r, (T)1), (T)0.5);
    proj_tol = T_FMAX(proj_tol, kProjTolMax);
    if (S->direct) FN(project_direct)(&S->P, &S->A, xtemp, ytemp, 1, x, y);
    else FN(project_cgls)(&S->A, xtemp, ytemp, 1, x, y, proj_tol, &S->cgls_iters);

    memcpy(ztemp, zprev, N * sizeof(T));
    FN(axpy)(N, -1, z, ztemp);
    nrm_s = S->rho * (S->nrmA * FN(nrm2)(m, ytemp) + FN(nrm2)(n, xtemp));
    memcpy(ztemp, z12, N * sizeof(T));
    FN(axpy)(N, -1, z, ztemp);
    nrm_r = S->nrmA * FN(nrm2)(n, xtemp) + FN(nrm2)(m, ytemp);

    int exact = 0;
    if (nrm_r < 10 * eps_pri && nrm_s < 10 * eps_dua) {
      memcpy(ztemp, z12, N * sizeof(T));
      FN(mat_mul)(&S->A, 'n', 1, x12, -1, ytemp, 0);
      nrm_r = FN(nrm2)(m, ytemp);
      memcpy(ztemp, z12, N * sizeof(T));
      FN(axpy)(N, 1, zt, ztemp);
      FN(axpy)(N, -1, zprev, ztemp);
      FN(mat_mul)(&S->A, 't', 1, ytemp, 1, xtemp, 0);
      nrm_s = S->rho * FN(nrm2)(n, xtemp);
      exact = 1;
    }
    converged = exact && nrm_r < eps_pri && nrm_s < eps_dua && (!S->gap_stop || gap < eps_gap);
    if (converged || k == S->max_iter - 1) { S->final_iter = k; break; }

    FN(axpy)(N, kAlpha, z12, zt);
    FN(axpy)(N, 1 - kAlpha, zprev, zt);
    FN(axpy)(N, -1, z, zt);

    if (S->adaptive_rho) {
      if (k > 0 && k % 50 == 0 && eps_pri > 0 && eps_dua > 0) {
        T pn = nrm_r / eps_pri, dn = nrm_s / eps_dua;
        if (pn > 0 && dn > 0) {
          T imb = pn / dn;
          if (imb > 10 || imb < (T)1 / 10) {
            T ratio = T_SQRT(imb);
            ratio = T_FMAX((T)0.67, T_FMIN((T)1.5, ratio));
            T rho_new = S->rho * ratio;
            rho_new = T_FMAX(kRhoMin, T_FMIN(kRhoMax, rho_new));
            if (T_FABS(rho_new - S->rho) / S->rho > (T)0.05) {
              T sc = S->rho / rho_new;
              S->rho = rho_new;
              FN(scal)(N, sc, zt);
            }
          }
        }
      } else if (nrm_s < xi * eps_dua && nrm_r > xi * eps_pri && kTau * (T)k > (T)kd) {
        if (S->rho < kRhoMax) { S->rho *= delta; FN(scal)(N, 1 / delta, zt); delta = kGamma * delta; ku = k; }
      } else if (nrm_s > xi * eps_dua && nrm_r < xi * eps_pri && kTau * (T)k > (T)ku) {
        if (S->rho > kRhoMin) { S->rho /= delta; FN(scal)(N, delta, zt); delta = kGamma * delta; kd = k; }
      } else if (nrm_s < xi * eps_dua && nrm_r < xi * eps_pri) {
        xi *= kKappa;
      } else {
        delta = kDeltaMin;
      }
    }
    prev_nrm_r = nrm_r;
  }

  S->optval = FN(oracle_func_vec)(m, f_h, fa, f_b, fc, fd, fe, y12) +
              FN(oracle_func_vec)(n, g_h, ga, g_b, gc, gd, ge, x12);
  int status = converged ? 0 : 3;

  memcpy(ztemp, zt, N * sizeof(T));
  FN(axpy)(N, -1, zprev, ztemp);
  FN(axpy)(N, 1, z12, ztemp);
  FN(scal)(N, -S->rho, ztemp);
  for (size_t i = 0; i < m; ++i) { ytemp[i] *= d[i]; y12[i] /= d[i]; }
  for (size_t j = 0; j < n; ++j) { xtemp[j] /= e[j]; x12[j] *= e[j]; }
  memcpy(S->x, x12, n * sizeof(T)); memcpy(S->y, y12, m * sizeof(T));
  memcpy(S->mu, xtemp, n * sizeof(T)); memcpy(S->lambda, ytemp, m * sizeof(T));
  memcpy(z, zprev, N * sizeof(T));

  free(zprev); free(ztemp); free(z12);
  free(fa); free(fc); free(fd); free(fe); free(ga); free(gc); free(gd); free(ge);
  return status;
}

void FN(oracle_get)(void *h, T *x, T *y, T *lambda, T *mu, T *optval, unsigned *final_iter, T *rho,
                    long *cgls_iters) {
  FN(solver) *S = (FN(solver) *)h;
  if (x) memcpy(x, S->x, S->A.n * sizeof(T));
  if (y) memcpy(y, S->y, S->A.m * sizeof(T));
  if (lambda) memcpy(lambda, S->lambda, S->A.m * sizeof(T));
  if (mu) memcpy(mu, S->mu, S->A.n * sizeof(T));
  if (optval) *optval = S->optval;
  if (final_iter) *final_iter = S->final_iter;
  if (rho) *rho = S->rho;
  if (cgls_iters) *cgls_iters = S->cgls_iters;
}

#undef CAT2
#undef CAT
#undef FN
