/*
 * oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the reference's algorithm for
 * the hot path.  It exists to CHECK the CUDA path;
 * it is never linked into, imported by, or called from the product
 * (gtsam_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  The restatement is validated against the unmodified
 * reference compiled here into oracle/_ref/libgtsam_ref.so (oracle/Makefile,
 * oracle/ref_harness.cpp) and against the golden vectors of the reference's
 * own tests committed under tests/golden/ (see tests/test_oracle_golden.py).
 *
 * It consumes the same problem description as the product's C-ABI
 * (include/gtsam_b200.h: b200_problem_desc) so that the parity tests feed
 * byte-identical inputs to both sides.
 */
#ifndef GTSAM_B200_ORACLE_H
#define GTSAM_B200_ORACLE_H

#include <stdint.h>
#include "../include/gtsam_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_problem orc_problem;

/* geometry primitives (for unit tests against the reference's KATs) */
void orc_so3_expmap(const double w[3], double R[9]);
void orc_so3_logmap(const double R[9], double w[3]);
void orc_pose3_expmap(const double xi[6], double T[12]);
void orc_pose3_logmap(const double T[12], double xi[6]);
void orc_pose3_compose(const double A[12], const double B[12], double C[12]);
void orc_pose3_inverse(const double A[12], double C[12]);
void orc_pose3_adjoint_map(const double T[12], double Ad[36]); /* row-major 6x6 */
/* choleskyPartial on a column-major n x n matrix (upper triangle used);
 * returns 1 on success, 0 on failure (gtsam/base/cholesky.cpp:107-158) */
int orc_cholesky_partial(double* ABC, int64_t n, int64_t nFrontal);

int orc_problem_create(const b200_problem_desc* desc, orc_problem** out);
/* GaussianFactorGraph level (b200_linear_create): JacobianFactors of any arity; orc_solve, orc_get_delta,
 * orc_hessian_diagonal, orc_get_conditional, orc_get_cliques and the marginals work on the result */
int orc_linear_create(const b200_linear_desc* desc, orc_problem** out);
int orc_linear_update(orc_problem* p, int64_t group, const double* Ab, const double* sigmas);
int orc_linear_update_hessian(orc_problem* p, int64_t hgroup, const double* info);
void orc_problem_destroy(orc_problem* p);
void orc_set_values(orc_problem* p, const double* packed);
void orc_get_values(const orc_problem* p, double* packed);
int64_t orc_values_size(const orc_problem* p);
int64_t orc_delta_size(const orc_problem* p);

double orc_error(orc_problem* p);
void orc_linearize(orc_problem* p);
/* mirror of b200_set_jacobian_precision: round the whitened [A|b] to float after every linearize */
void orc_set_jacobian_fp32(orc_problem* p, int on);
void orc_get_jacobians(const orc_problem* p, int64_t group, double* out);
void orc_hessian_diagonal(const orc_problem* p, double* out);
int orc_solve(orc_problem* p, double lambda, int diagonal_damping, double min_diagonal,
              double max_diagonal, double* lin_err0, double* lin_err_delta, int64_t* fail_var);
void orc_get_delta(const orc_problem* p, double* out);
double orc_try_step(orc_problem* p);
void orc_accept_step(orc_problem* p);

/* LM / GN control logic */
typedef struct orc_lm {
  orc_problem* prob;
  b200_lm_params params;
  b200_lm_state state;
} orc_lm;
void orc_lm_init(orc_lm* lm, orc_problem* p, const b200_lm_params* params);
int orc_lm_iterate(orc_lm* lm);
int orc_lm_optimize(orc_lm* lm);
int orc_gn_iterate(orc_problem* p, double* new_error);
/* DoglegOptimizer::iterate; error_io = state error, delta_io = trust region radius */
int orc_dogleg_iterate(orc_problem* p, double* error_io, double* delta_io);

/* Marginals::marginalCovariance (gtsam/nonlinear/Marginals.cpp:118-154): d x d column-major block of H^-1 */
int orc_marginal_covariance(orc_problem* p, int64_t var, double* out);
/* Marginals::jointMarginalCovariance: vars sorted ascending; out D x D column-major, blocks in that order */
int orc_joint_marginal_covariance(orc_problem* p, const int64_t* vars, int64_t nv, double* out);
/* x = H^-1 g with the factorisation left by the last successful orc_solve */
void orc_solve_rhs(const orc_problem* p, const double* g, double* x);

/* symbolic introspection */
void orc_symbolic_info_get(const orc_problem* p, b200_symbolic_info* info);
void orc_get_cliques(const orc_problem* p, int64_t* frontal_ptr, int64_t* frontal_vars,
                     int64_t* separator_ptr, int64_t* separator_vars, int64_t* parent);
/* conditional [R S d] of clique c: f x (f+s+1) column-major (parity of a14) */
void orc_get_conditional(const orc_problem* p, int64_t clique, double* out);

#ifdef __cplusplus
}
#endif
#endif
