/*
 * pogs_b200 -- C ABI of the B200-native graph-form solver.
 *
 *   minimize  sum_i f_i(y_i) + sum_j g_j(x_j)   subject to  y = A x
 *   f_i(v) = c_i h_i(a_i v - b_i) + d_i v + (e_i/2) v^2      (same for g_j)
 *
 * Part 1 is the drop-in boundary: the four graph-form entry points of the
 * reference C interface, with identical names, argument order, enum values,
 * ownership rules (all pointers are HOST pointers owned by the caller, inputs
 * are copied, nothing is retained) and return codes.  A build of this library
 * installed as libpogs_cpu.so is loadable by the reference's own
 * python/pogs/graph.py unchanged (it binds PogsD and PogsSparseD at import,
 * graph.py:167-233).
 *
 * Part 2 is additive: a persistent solver handle that exposes what the
 * reference only offers in C++ (pogs::PogsDirect / PogsIndirect objects,
 * src/include/pogs.h:55-131) -- cached equilibration and factor, warm starts,
 * lambda paths -- plus device-pointer inputs and timing read-outs.
 *
 * Part 3 are unit-level hooks used by the parity tests.
 *
 * No CPU fallback exists: every entry point needs a CUDA device and returns
 * POGS_ERROR (6) with a message on stderr when none is usable.
 */
#ifndef POGS_B200_H_
#define POGS_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == reference src/interface_c/pogs_c.h:51 */
enum ORD { COL_MAJ, ROW_MAJ };

/* == reference src/interface_c/pogs_c.h:54-69 (values pinned by the reference's
 * tests/test_c_interface.cpp:149-154) and enum Function, src/include/prox_lib.h:23-38 */
enum FUNCTION { ABS, EXP, HUBER, IDENTITY, INDBOX01, INDEQ0, INDGE0, INDLE0, LOGISTIC, MAXNEG0, MAXPOS0,
                NEGENTR, NEGLOG, RECIPR, SQUARE, ZERO };

/* == PogsStatus, reference src/include/pogs.h:31-37 (the int every entry point returns) */
enum POGS_STATUS { POGS_SUCCESS, POGS_INFEASIBLE, POGS_UNBOUNDED, POGS_MAX_ITER, POGS_NAN_FOUND,
                   POGS_INVALID_CONE, POGS_ERROR };

/* ---------------------------------------------------------------------------
 * Part 1 -- reference entry points.
 * Replaces PogsD / PogsS (reference src/interface_c/pogs_c.h:75-91,
 * pogs_c.cpp:8-55,111-160): dense A (m x n, ord), direct projector.
 * final_iter is the zero-based index of the last iteration.
 * ------------------------------------------------------------------------- */
int PogsD(enum ORD ord, size_t m, size_t n, const double *A,
          const double *f_a, const double *f_b, const double *f_c, const double *f_d, const double *f_e,
          const enum FUNCTION *f_h,
          const double *g_a, const double *g_b, const double *g_c, const double *g_d, const double *g_e,
          const enum FUNCTION *g_h,
          double rho, double abs_tol, double rel_tol, unsigned int max_iter, unsigned int verbose,
          int adaptive_rho, int gap_stop,
          double *x, double *y, double *l, double *optval, unsigned int *final_iter);

int PogsS(enum ORD ord, size_t m, size_t n, const float *A,
          const float *f_a, const float *f_b, const float *f_c, const float *f_d, const float *f_e,
          const enum FUNCTION *f_h,
          const float *g_a, const float *g_b, const float *g_c, const float *g_d, const float *g_e,
          const enum FUNCTION *g_h,
          float rho, float abs_tol, float rel_tol, unsigned int max_iter, unsigned int verbose,
          int adaptive_rho, int gap_stop,
          float *x, float *y, float *l, float *optval, unsigned int *final_iter);

/* Replaces PogsSparseD / PogsSparseS (reference src/interface_c/pogs_c.h:93-119,
 * pogs_c.cpp:57-108,162-203): CSR (ROW_MAJ, ptr has m+1 entries) or CSC
 * (COL_MAJ, n+1 entries), int32 indices, CGLS projector. */
int PogsSparseD(enum ORD ord, size_t m, size_t n, size_t nnz,
                const double *data, const int *ptr, const int *ind,
                const double *f_a, const double *f_b, const double *f_c, const double *f_d, const double *f_e,
                const enum FUNCTION *f_h,
                const double *g_a, const double *g_b, const double *g_c, const double *g_d, const double *g_e,
                const enum FUNCTION *g_h,
                double rho, double abs_tol, double rel_tol, unsigned int max_iter, unsigned int verbose,
                int adaptive_rho, int gap_stop,
                double *x, double *y, double *l, double *optval, unsigned int *final_iter);

int PogsSparseS(enum ORD ord, size_t m, size_t n, size_t nnz,
                const float *data, const int *ptr, const int *ind,
                const float *f_a, const float *f_b, const float *f_c, const float *f_d, const float *f_e,
                const enum FUNCTION *f_h,
                const float *g_a, const float *g_b, const float *g_c, const float *g_d, const float *g_e,
                const enum FUNCTION *g_h,
                float rho, float abs_tol, float rel_tol, unsigned int max_iter, unsigned int verbose,
                int adaptive_rho, int gap_stop,
                float *x, float *y, float *l, float *optval, unsigned int *final_iter);

/* ---------------------------------------------------------------------------
 * Part 1b -- cone-form entry points, first slice (reference: src/interface_c/pogs_c.h:120-243,
 * PogsCone<T,M,P>::Solve src/cpu/pogs.cpp:1925-1976 with PogsObjectiveCone :642-790).
 *     minimize c^T x   subject to   b - A x in K_y,  x in K_x
 * Supported here: the separable cones CONE_ZERO / CONE_NON_NEG / CONE_NON_POS (linear programs), for
 * which the reference's cone objective is exactly a separable graph-form objective
 *     f_i = indicator{ b_i - y_i in K },   g_j = c_j x_j + indicator{ x_j in K }
 * (the encoding of the reference's own examples/cpp/lp_eq.cpp, lp_ineq.cpp), run on the device path
 * above, with the reference's objective normalisation (c scaled to unit norm after equilibration,
 * pogs.cpp:737-754).  PogsCone* use the CGLS projector, PogsConeDirect* the cached factor, like the
 * reference.  SOC / SDP / exponential cones and the quadratic (Q) variants are not implemented:
 * POGS_ERROR with a message; index sets out of range or overlapping: POGS_INVALID_CONE (5).
 * Enum values pinned by the reference's tests/test_c_interface.cpp:157-162.
 * ------------------------------------------------------------------------- */
enum CONE { CONE_ZERO, CONE_NON_NEG, CONE_NON_POS, CONE_SOC, CONE_SDP, CONE_EXP_PRIMAL, CONE_EXP_DUAL };
struct ConeConstraintC {
  enum CONE cone;
  unsigned int *indices;
  unsigned int size;
};
int PogsConeD(enum ORD ord, size_t m, size_t n, const double *A, const double *b, const double *c,
              const struct ConeConstraintC *cones_x, size_t num_cones_x,
              const struct ConeConstraintC *cones_y, size_t num_cones_y,
              double rho, double abs_tol, double rel_tol, unsigned int max_iter, unsigned int verbose,
              int adaptive_rho, int gap_stop, double *x, double *y, double *l, double *optval,
              unsigned int *final_iter);
int PogsConeS(enum ORD ord, size_t m, size_t n, const float *A, const float *b, const float *c,
              const struct ConeConstraintC *cones_x, size_t num_cones_x,
              const struct ConeConstraintC *cones_y, size_t num_cones_y,
              float rho, float abs_tol, float rel_tol, unsigned int max_iter, unsigned int verbose,
              int adaptive_rho, int gap_stop, float *x, float *y, float *l, float *optval,
              unsigned int *final_iter);
int PogsConeDirectD(enum ORD ord, size_t m, size_t n, const double *A, const double *b, const double *c,
                    const struct ConeConstraintC *cones_x, size_t num_cones_x,
                    const struct ConeConstraintC *cones_y, size_t num_cones_y,
                    double rho, double abs_tol, double rel_tol, unsigned int max_iter, unsigned int verbose,
                    int adaptive_rho, int gap_stop, double *x, double *y, double *l, double *optval,
                    unsigned int *final_iter);
int PogsConeDirectS(enum ORD ord, size_t m, size_t n, const float *A, const float *b, const float *c,
                    const struct ConeConstraintC *cones_x, size_t num_cones_x,
                    const struct ConeConstraintC *cones_y, size_t num_cones_y,
                    float rho, float abs_tol, float rel_tol, unsigned int max_iter, unsigned int verbose,
                    int adaptive_rho, int gap_stop, float *x, float *y, float *l, float *optval,
                    unsigned int *final_iter);
/* Quadratic-objective variants: bound by python/pogs_cone.py at import; not implemented (POGS_ERROR). */
int PogsConeQD(enum ORD ord, size_t m, size_t n, const double *A, const double *b, const double *c, const double *P,
               const struct ConeConstraintC *cones_x, size_t num_cones_x,
               const struct ConeConstraintC *cones_y, size_t num_cones_y,
               double rho, double abs_tol, double rel_tol, unsigned int max_iter, unsigned int verbose,
               int adaptive_rho, int gap_stop, double *x, double *y, double *l, double *optval,
               unsigned int *final_iter);
int PogsConeDirectQD(enum ORD ord, size_t m, size_t n, const double *A, const double *b, const double *c, const double *P,
                     const struct ConeConstraintC *cones_x, size_t num_cones_x,
                     const struct ConeConstraintC *cones_y, size_t num_cones_y,
                     double rho, double abs_tol, double rel_tol, unsigned int max_iter, unsigned int verbose,
                     int adaptive_rho, int gap_stop, double *x, double *y, double *l, double *optval,
                     unsigned int *final_iter);

/* ---------------------------------------------------------------------------
 * Part 2 -- persistent solver handle (C view of pogs::PogsDirect<T,MatrixDense<T>>
 * and pogs::PogsIndirect<T,MatrixSparse<T>>, reference src/include/pogs.h:55-131,
 * 155-158).  The matrix is uploaded and set up once (lazily, on the first
 * solve); z, z~ and rho persist between solves, so a second solve with changed
 * f/g is warm-started exactly like the reference's examples/cpp/lasso_path.cpp.
 * Functions ending in _s take float, _d take double.  Return 0 / handle on
 * success; NULL or POGS_ERROR on failure (see pogs_b200_last_error).
 * ------------------------------------------------------------------------- */
typedef struct pogs_b200_handle pogs_b200_handle;

/* a_on_device != 0: A is a CUDA device pointer on the current device.  A handle lives on the
 * device that was current when it was created; later calls may come from any thread (they make
 * that device current for their duration and serialise on a per-device mutex). */
pogs_b200_handle *pogs_b200_create_dense_s(enum ORD ord, size_t m, size_t n, const float *A, int a_on_device);
pogs_b200_handle *pogs_b200_create_dense_d(enum ORD ord, size_t m, size_t n, const double *A, int a_on_device);
/* == pogs::PogsIndirect<T, MatrixDense<T>> (pogs.h:155-158; ProjectorCgls<T, MatrixDense<T>>,
 * src/cpu/projector/projector_cgls.cpp:91-97): dense A with the CGLS projector instead of the
 * cached factor -- for matrices whose min(m,n)^2 factor does not fit, or when the one-time Gram
 * matrix + Cholesky is not worth it.  The reference reaches this only from C++; no one-shot
 * C entry point exists for it there either. */
pogs_b200_handle *pogs_b200_create_dense_indirect_s(enum ORD ord, size_t m, size_t n, const float *A, int a_on_device);
pogs_b200_handle *pogs_b200_create_dense_indirect_d(enum ORD ord, size_t m, size_t n, const double *A, int a_on_device);
pogs_b200_handle *pogs_b200_create_sparse_s(enum ORD ord, size_t m, size_t n, size_t nnz, const float *data,
                                            const int *ptr, const int *ind);
pogs_b200_handle *pogs_b200_create_sparse_d(enum ORD ord, size_t m, size_t n, size_t nnz, const double *data,
                                            const int *ptr, const int *ind);
void pogs_b200_destroy(pogs_b200_handle *h);

/* Setters == SetRho/SetAbsTol/SetRelTol/SetMaxIter/SetVerbose/SetAdaptiveRho/
 * SetGapStop (pogs.h:103-110); values are passed as double for both precisions. */
int pogs_b200_set_params(pogs_b200_handle *h, double rho, double abs_tol, double rel_tol, unsigned int max_iter,
                         unsigned int verbose, int adaptive_rho, int gap_stop);
int pogs_b200_set_rho(pogs_b200_handle *h, double rho);
/* == SetInitX + SetInitLambda (pogs.h:111-118); both are required (the reference
 * aborts on a one-sided warm start, pogs.cpp:159-179; here it is an error). */
int pogs_b200_set_init_s(pogs_b200_handle *h, const float *x, const float *lambda);
int pogs_b200_set_init_d(pogs_b200_handle *h, const double *x, const double *lambda);
/* Per-phase CUDA-event timing of every iteration (one launch per iteration,
 * host sync in between); for benchmarks only. */
int pogs_b200_set_profile(pogs_b200_handle *h, int on);

/* == PogsSeparable::Solve(f, g) (pogs.h:122-131): returns the POGS_STATUS. */
int pogs_b200_solve_s(pogs_b200_handle *h,
                      const float *f_a, const float *f_b, const float *f_c, const float *f_d, const float *f_e,
                      const int *f_h,
                      const float *g_a, const float *g_b, const float *g_c, const float *g_d, const float *g_e,
                      const int *g_h);
int pogs_b200_solve_d(pogs_b200_handle *h,
                      const double *f_a, const double *f_b, const double *f_c, const double *f_d, const double *f_e,
                      const int *f_h,
                      const double *g_a, const double *g_b, const double *g_c, const double *g_d, const double *g_e,
                      const int *g_h);

/* Getters == GetX/GetY/GetLambda/GetMu/GetOptval/GetFinalIter/GetRho (pogs.h:87-101).
 * Any output pointer may be NULL. */
int pogs_b200_get_solution_s(pogs_b200_handle *h, float *x, float *y, float *lambda, float *mu, float *optval,
                             unsigned int *final_iter, float *rho);
int pogs_b200_get_solution_d(pogs_b200_handle *h, double *x, double *y, double *lambda, double *mu, double *optval,
                             unsigned int *final_iter, double *rho);

/* Timing of the last solve, milliseconds / counts:
 *  out[0] h2d of A        out[1] setup (equilibrate+normest+factor)   out[2] ADMM loop (CUDA events)
 *  out[3] whole solve (host clock)      out[4] iterations run         out[5] iterations that took the
 *  exact-residual branch  out[6..10] profile mode only: prox, A^T product, factor apply, A product,
 *  control+exact residuals (summed over out[11] profiled iterations)  out[12] CGLS inner iterations
 *  out[13..15] parts of the setup: equilibration, norm estimate, Gram matrix (factor = rest) */
int pogs_b200_get_timing(pogs_b200_handle *h, double out[16]);

/* ---------------------------------------------------------------------------
 * Part 2b -- row-block multi-GPU (new; the reference has no multi-device path).
 * One process per GPU.  A = [A_1; ...; A_G] by rows: rank g holds A_g (m_local x n,
 * row-major), the matching slices of f, y, lambda, and a replica of g, x, mu.  The
 * only per-iteration exchange is the sum over ranks of the n-vector A_g^T y_g plus a
 * few scalars; it runs over NVLink peer memory inside the A^T kernel (no collective
 * launch).  The communicator owns one cudaMalloc'ed region per rank, exported with
 * CUDA IPC: create it on every rank, all-gather the 64-byte handles with whatever the
 * host program uses for plumbing (torch.distributed in pogs_b200/dist.py), open them.
 * ------------------------------------------------------------------------- */
typedef struct pogs_b200_comm pogs_b200_comm;
/* slot_bytes >= 8 * (n rounded up to 4): size of one exchange slot. */
pogs_b200_comm *pogs_b200_comm_create(int rank, int world, size_t slot_bytes);
int pogs_b200_comm_handle(pogs_b200_comm *c, void *out64);
int pogs_b200_comm_open(pogs_b200_comm *c, const void *handles /* world x 64 bytes, rank-major */);
void pogs_b200_comm_destroy(pogs_b200_comm *c);
/* In-place sum over ranks of a DEVICE buffer (len elements; allocation padded to 16 B). */
int pogs_b200_comm_allreduce_s(pogs_b200_comm *c, float *dev_buf, size_t len);
int pogs_b200_comm_allreduce_d(pogs_b200_comm *c, double *dev_buf, size_t len);
/* Solver on one row block; m_global > n required.  The handle API above applies: f
 * arrays have m_local entries, g arrays n; get_solution returns x (replicated), and the
 * local slices of y and lambda. */
pogs_b200_handle *pogs_b200_create_dense_rowblock_s(size_t m_local, size_t n, size_t m_global, const float *A_local,
                                                    int a_on_device, pogs_b200_comm *comm);
pogs_b200_handle *pogs_b200_create_dense_rowblock_d(size_t m_local, size_t n, size_t m_global, const double *A_local,
                                                    int a_on_device, pogs_b200_comm *comm);

/* More counters of the last solve: out[0] iterations that ran on ONE pass over A (committed
 * speculation of the single-pass kernel), out[1] power-iteration sweeps of the norm estimate,
 * out[2] factor time (ms), out[3] see below. */
int pogs_b200_get_stats(pogs_b200_handle *h, double out[8]);
/* out[3] of get_stats: times the rare path of the one-launch iteration ran (two-pass kernels and/or
 * the standalone factor apply before the pass); out[4]: 1 when the iterations ran on the one-launch
 * iteration kernel (k_admm_pass); out[5]: committed speculations across a rho change (the kernel
 * speculates on the controller repeating its last rho action).
 * With POGS_B200_PASS_TIMING=1 in the environment when the handle is created: mean time (us per
 * iteration, measured with %globaltimer on CTA 0) of the phases of the one-launch iteration kernel:
 * out[0] A: pass over A   out[1] grid barrier   out[2] B: fold + exchange + x half-step (speculative)
 * out[3] grid barrier     out[4] C: controller  out[5] D: factor apply (packed triangle streamed)
 * out[6] grid barrier     out[7] E: fold + exchange + x half-step of the next iteration
 * out[8] grid barrier before the next iteration's pass (same launch)     out[9..15] reserved */
int pogs_b200_get_pass_phases(pogs_b200_handle *h, double out[16]);

/* Device buffers come from a library-owned memory pool that keeps freed blocks for the next
 * solver (the reference's malloc/free of its matrix copy, src/cpu/matrix/matrix_dense.cpp:76-90,
 * has no such cost to hide).  This returns the cached blocks to the driver. */
void pogs_b200_trim_memory(void);

const char *pogs_b200_last_error(void);
/* Number of kernel launches issued by this library since load (all handles). */
unsigned long long pogs_b200_launch_count(void);

/* ---------------------------------------------------------------------------
 * Part 3 -- unit-level hooks for the parity tests (host pointers).
 * ------------------------------------------------------------------------- */
/* Host logic, no device needed: the P x Q tile grid the sparse operator chooses for one compressed copy
 * (rows x cols, nnz entries, elem = 4 or 8 bytes per value) on a GPU with `sms` SMs.  out = {P, Q, rows per
 * tile, columns per tile, 32-row slices per tile, tiles, dynamic shared memory of the product kernel in bytes,
 * shared-memory limit the planner assumes}; returns 0, or 1 when the layout does not apply. */
int pogs_b200_plan_sparse_tiles(size_t rows, size_t cols, size_t nnz, unsigned sms, size_t elem, unsigned long long out[8]);
/* Vector ProxEval / FuncEval on the device (reference src/include/prox_lib.h:504-529). */
int pogs_b200_prox_eval_s(size_t n, const int *h, const float *a, const float *b, const float *c, const float *d,
                          const float *e, float rho, const float *in, float *out);
int pogs_b200_prox_eval_d(size_t n, const int *h, const double *a, const double *b, const double *c,
                          const double *d, const double *e, double rho, const double *in, double *out);
int pogs_b200_func_eval_s(size_t n, const int *h, const float *a, const float *b, const float *c, const float *d,
                          const float *e, const float *in, double *sum);
int pogs_b200_func_eval_d(size_t n, const int *h, const double *a, const double *b, const double *c,
                          const double *d, const double *e, const double *in, double *sum);
/* out = op(A) v on the raw (un-equilibrated) matrix; trans: 0 -> A v, 1 -> A^T v;
 * square != 0 uses A.^2 (== MatrixDense::Mul, reference src/cpu/matrix/matrix_dense.cpp:93-113). */
int pogs_b200_gemv_s(enum ORD ord, size_t m, size_t n, const float *A, int trans, int square, const float *v,
                     float *out);
int pogs_b200_gemv_d(enum ORD ord, size_t m, size_t n, const double *A, int trans, int square, const double *v,
                     double *out);
/* Setup results: d (m), e (n), estimated ||A^||_2 (== MatrixDense::Equil + Norm2Est). */
/* One-time Gram matrix of the direct projector (reference: cblas_ssyrk in ProjectorDirect::Init,
 * src/cpu/projector/projector_direct_dense.cpp:62-81): G (n x n, row-major) = A^T A for a row-major
 * m x n host array.  use_tc = 1: tcgen05 3xTF32 kernel; use_tc = 0: the library's CUDA-core product
 * (dense_factor.cuh).  G is full and symmetric in both cases. */
int pogs_b200_gram_s(size_t m, size_t n, const float *A, float *G, int use_tc);
/* Bring-up aid: also returns the first shared-memory pipeline stage as the tensor core reads it
 * (12288 floats) and the raw 128 x 256 accumulator of the first tile. */
int pogs_b200_gram_debug_s(size_t m, size_t n, const float *A, float *G, float *stage0, float *acc0);

int pogs_b200_get_equil_s(pogs_b200_handle *h, float *d, float *e, float *nrmA);
int pogs_b200_get_equil_d(pogs_b200_handle *h, double *d, double *e, double *nrmA);
/* One projection onto {y = A^ x} in the equilibrated space (== Projector::Project). */
int pogs_b200_project_s(pogs_b200_handle *h, const float *x0, const float *y0, float *x, float *y);
int pogs_b200_project_d(pogs_b200_handle *h, const double *x0, const double *y0, double *x, double *y);

#ifdef __cplusplus
}
#endif
#endif /* POGS_B200_H_ */
