/*
 * oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C, single-threaded restatement of the reference algorithm for the hot
 * path.  Every function cites the reference file:line it follows (paths
 * relative to the borglab/gtsam tree).  Written for clarity, not speed: dense
 * per-clique matrices, no blocking, literal control flow of the reference.
 *
 * Build flags of the reference that this restatement assumes:
 * Rot3 = 3x3 matrix, GTSAM_POSE3_EXPMAP, GTSAM_ROT3_EXPMAP,
 * GTSAM_THROW_CHEIRALITY_EXCEPTION, fast (approximate) BetweenFactor Jacobian.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------ */
/* static layout tables (mirror include/gtsam_b200.h)                        */
/* ------------------------------------------------------------------------ */
static const int VAR_STORAGE[B200_NUM_VAR_TYPES] = {12, 3, 17, 3};
static const int VAR_DIM[B200_NUM_VAR_TYPES] = {6, 3, 9, 3};
static const int F_ARITY[B200_NUM_FACTOR_TYPES] = {2, 1, 1, 2, 2, 1, 2, 1};
static const int F_MEAS[B200_NUM_FACTOR_TYPES] = {12, 12, 3, 2, 2, 17, 3, 3};
static const int F_DIM[B200_NUM_FACTOR_TYPES] = {6, 6, 3, 2, 2, 9, 3, 3};
static const int F_VT[B200_NUM_FACTOR_TYPES][2] = {
    {B200_VAR_POSE3, B200_VAR_POSE3},       {B200_VAR_POSE3, -1},
    {B200_VAR_POINT3, -1},                  {B200_VAR_POSE3, B200_VAR_POINT3},
    {B200_VAR_CAM_BUNDLER, B200_VAR_POINT3}, {B200_VAR_CAM_BUNDLER, -1},
    {B200_VAR_POSE2, B200_VAR_POSE2},       {B200_VAR_POSE2, -1}};

/* ------------------------------------------------------------------------ */
/* small dense helpers (3x3 row-major)                                       */
/* ------------------------------------------------------------------------ */
static void m3_mul(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, T, sizeof T);
}
static void m3_tr(const double* A, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * j + i];
  memcpy(C, T, sizeof T);
}
static void m3_vec(const double* A, const double* v, double* r) {
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
  r[0] = t[0]; r[1] = t[1]; r[2] = t[2];
}
static void skew(double x, double y, double z, double* W) {
  W[0] = 0; W[1] = -z; W[2] = y;
  W[3] = z; W[4] = 0; W[5] = -x;
  W[6] = -y; W[7] = x; W[8] = 0;
}

/* ------------------------------------------------------------------------ */
/* SO(3)                                                                     */
/* ------------------------------------------------------------------------ */
/* so3::ExpmapFunctor, gtsam/geometry/SO3.cpp:49-87 (Rot3::Expmap -> SO3::Expmap :202-211) */
void orc_so3_expmap(const double w[3], double R[9]) {
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double W[9];
  skew(w[0], w[1], w[2], W);
  if (theta2 <= DBL_EPSILON) { /* nearZero: I + W */
    for (int i = 0; i < 9; i++) R[i] = W[i];
    R[0] += 1; R[4] += 1; R[8] += 1;
    return;
  }
  const double theta = sqrt(theta2);
  const double sin_theta = sin(theta);
  const double s2 = sin(theta / 2.0);
  const double one_minus_cos = 2.0 * s2 * s2;
  double K[9], KK[9];
  for (int i = 0; i < 9; i++) K[i] = W[i] / theta;
  m3_mul(K, K, KK);
  for (int i = 0; i < 9; i++) R[i] = sin_theta * K[i] + one_minus_cos * KK[i];
  R[0] += 1; R[4] += 1; R[8] += 1;
}

/* SO3::Logmap, gtsam/geometry/SO3.cpp:247-325 (three branches) */
void orc_so3_logmap(const double R[9], double w[3]) {
  const double R11 = R[0], R12 = R[1], R13 = R[2];
  const double R21 = R[3], R22 = R[4], R23 = R[5];
  const double R31 = R[6], R32 = R[7], R33 = R[8];
  const double tr = R11 + R22 + R33;
  if (tr + 1.0 < 1e-3) {
    if (R33 > R22 && R33 > R11) {
      const double W = R21 - R12, Q1 = 2.0 + 2.0 * R33, Q2 = R31 + R13, Q3 = R23 + R32;
      const double r = sqrt(Q1), one_over_r = 1 / r;
      const double norm = sqrt(Q1 * Q1 + Q2 * Q2 + Q3 * Q3 + W * W);
      const double sgn_w = W < 0 ? -1.0 : 1.0;
      const double mag = M_PI - (2 * sgn_w * W) / norm;
      const double scale = 0.5 * one_over_r * mag;
      w[0] = sgn_w * scale * Q2; w[1] = sgn_w * scale * Q3; w[2] = sgn_w * scale * Q1;
    } else if (R22 > R11) {
      const double W = R13 - R31, Q1 = 2.0 + 2.0 * R22, Q2 = R23 + R32, Q3 = R12 + R21;
      const double r = sqrt(Q1), one_over_r = 1 / r;
      const double norm = sqrt(Q1 * Q1 + Q2 * Q2 + Q3 * Q3 + W * W);
      const double sgn_w = W < 0 ? -1.0 : 1.0;
      const double mag = M_PI - (2 * sgn_w * W) / norm;
      const double scale = 0.5 * one_over_r * mag;
      w[0] = sgn_w * scale * Q3; w[1] = sgn_w * scale * Q1; w[2] = sgn_w * scale * Q2;
    } else {
      const double W = R32 - R23, Q1 = 2.0 + 2.0 * R11, Q2 = R12 + R21, Q3 = R31 + R13;
      const double r = sqrt(Q1), one_over_r = 1 / r;
      const double norm = sqrt(Q1 * Q1 + Q2 * Q2 + Q3 * Q3 + W * W);
      const double sgn_w = W < 0 ? -1.0 : 1.0;
      const double mag = M_PI - (2 * sgn_w * W) / norm;
      const double scale = 0.5 * one_over_r * mag;
      w[0] = sgn_w * scale * Q1; w[1] = sgn_w * scale * Q2; w[2] = sgn_w * scale * Q3;
    }
  } else {
    double magnitude;
    const double tr_3 = tr - 3.0;
    if (tr_3 < -1e-6) {
      const double theta = acos((tr - 1.0) / 2.0);
      magnitude = theta / (2.0 * sin(theta));
    } else {
      magnitude = 0.5 - tr_3 / 12.0 + tr_3 * tr_3 / 60.0;
    }
    w[0] = magnitude * (R32 - R23);
    w[1] = magnitude * (R13 - R31);
    w[2] = magnitude * (R21 - R12);
  }
}

/* ------------------------------------------------------------------------ */
/* SE(3): T = [R(9) row-major, t(3)]                                         */
/* ------------------------------------------------------------------------ */
/* Pose3::operator*, gtsam/geometry/Pose3.h:114-116 */
void orc_pose3_compose(const double A[12], const double B[12], double C[12]) {
  double R[9], t[3];
  m3_mul(A, B, R);
  m3_vec(A, B + 9, t);
  for (int i = 0; i < 3; i++) t[i] = A[9 + i] + t[i];
  memcpy(C, R, sizeof R);
  memcpy(C + 9, t, sizeof t);
}
/* Pose3::inverse, gtsam/geometry/Pose3.cpp:49-52 */
void orc_pose3_inverse(const double A[12], double C[12]) {
  double Rt[9], nt[3] = {-A[9], -A[10], -A[11]}, t[3];
  m3_tr(A, Rt);
  m3_vec(Rt, nt, t);
  memcpy(C, Rt, sizeof Rt);
  memcpy(C + 9, t, sizeof t);
}
/* Pose3::AdjointMap, gtsam/geometry/Pose3.cpp:57-63: [R 0; [t]x R, R] */
void orc_pose3_adjoint_map(const double T[12], double Ad[36]) {
  double S[9], A[9];
  skew(T[9], T[10], T[11], S);
  m3_mul(S, T, A);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Ad[6 * i + j] = T[3 * i + j];
      Ad[6 * i + 3 + j] = 0.0;
      Ad[6 * (3 + i) + j] = A[3 * i + j];
      Ad[6 * (3 + i) + 3 + j] = T[3 * i + j];
    }
}
/* Pose3::Expmap, gtsam/geometry/Pose3.cpp:169-185 */
void orc_pose3_expmap(const double xi[6], double T[12]) {
  const double* omega = xi;
  const double* v = xi + 3;
  orc_so3_expmap(omega, T);
  const double theta2 = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  if (theta2 > DBL_EPSILON) {
    const double wv = omega[0] * v[0] + omega[1] * v[1] + omega[2] * v[2];
    const double tp[3] = {omega[0] * wv, omega[1] * wv, omega[2] * wv};
    const double oxv[3] = {omega[1] * v[2] - omega[2] * v[1], omega[2] * v[0] - omega[0] * v[2],
                           omega[0] * v[1] - omega[1] * v[0]};
    double Roxv[3];
    m3_vec(T, oxv, Roxv);
    for (int i = 0; i < 3; i++) T[9 + i] = (oxv[i] - Roxv[i] + tp[i]) / theta2;
  } else {
    T[9] = v[0]; T[10] = v[1]; T[11] = v[2];
  }
}
/* Pose3::Logmap, gtsam/geometry/Pose3.cpp:188-208 */
void orc_pose3_logmap(const double T[12], double xi[6]) {
  double w[3];
  orc_so3_logmap(T, w);
  const double* Tt = T + 9;
  const double t = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  xi[0] = w[0]; xi[1] = w[1]; xi[2] = w[2];
  if (t < 1e-10) {
    xi[3] = Tt[0]; xi[4] = Tt[1]; xi[5] = Tt[2];
  } else {
    double W[9], WT[3], WWT[3];
    skew(w[0] / t, w[1] / t, w[2] / t, W);
    const double Tan = tan(0.5 * t);
    m3_vec(W, Tt, WT);
    m3_vec(W, WT, WWT);
    for (int i = 0; i < 3; i++)
      xi[3 + i] = Tt[i] - (0.5 * t) * WT[i] + (1 - t / (2. * Tan)) * WWT[i];
  }
}
/* LieGroup::between, gtsam/base/Lie.h:63-69: inverse(a)*b */
static void pose3_between(const double A[12], const double B[12], double C[12]) {
  double Ai[12];
  orc_pose3_inverse(A, Ai);
  orc_pose3_compose(Ai, B, C);
}
/* LieGroup::localCoordinates, gtsam/base/Lie.h:152-160 with Pose3::ChartAtOrigin::Local = Logmap */
static void pose3_local(const double A[12], const double B[12], double xi[6]) {
  double h[12];
  pose3_between(A, B, h);
  orc_pose3_logmap(h, xi);
}
/* LieGroup::retract, gtsam/base/Lie.h:131-133 */
static void pose3_retract(const double A[12], const double xi[6], double C[12]) {
  double E[12];
  orc_pose3_expmap(xi, E);
  orc_pose3_compose(A, E, C);
}

/* ------------------------------------------------------------------------ */
/* pinhole projection                                                        */
/* ------------------------------------------------------------------------ */
/* PinholeBase::project2, gtsam/geometry/CalibratedCamera.cpp:116-133 with
 * Pose3::transformTo (Pose3.cpp:371-388), Project (:88-94), Dpose (:27-34),
 * Dpoint (:37-46).  Returns 0 on cheirality failure (q.z <= 0).
 * Dpose 2x6 row-major, Dpoint 2x3 row-major. */
static int project2(const double pose[12], const double p[3], double pn[2], double Dpose[12],
                    double Dpoint[6]) {
  double Rt[9], d3[3] = {p[0] - pose[9], p[1] - pose[10], p[2] - pose[11]}, q[3];
  m3_tr(pose, Rt);
  m3_vec(Rt, d3, q);
  if (q[2] <= 0) return 0;
  const double d = 1.0 / q[2];
  const double u = q[0] * d, v = q[1] * d;
  pn[0] = u; pn[1] = v;
  if (Dpose) {
    const double uv = u * v, uu = u * u, vv = v * v;
    Dpose[0] = uv; Dpose[1] = -1 - uu; Dpose[2] = v; Dpose[3] = -d; Dpose[4] = 0; Dpose[5] = d * u;
    Dpose[6] = 1 + vv; Dpose[7] = -uv; Dpose[8] = -u; Dpose[9] = 0; Dpose[10] = -d; Dpose[11] = d * v;
  }
  if (Dpoint) {
    for (int j = 0; j < 3; j++) {
      Dpoint[j] = (Rt[j] - u * Rt[6 + j]) * d;
      Dpoint[3 + j] = (Rt[3 + j] - v * Rt[6 + j]) * d;
    }
  }
  return 1;
}

/* 2x2 (row-major) times 2xN (row-major) */
static void m22_mul(const double* D, const double* A, int ncols, double* C) {
  for (int j = 0; j < ncols; j++) {
    const double a0 = A[j], a1 = A[ncols + j];
    C[j] = D[0] * a0 + D[1] * a1;
    C[ncols + j] = D[2] * a0 + D[3] * a1;
  }
}

/* PinholePose<Cal3_S2>::_project (gtsam/geometry/PinholePose.h:89-109) with
 * Cal3_S2::uncalibrate (gtsam/geometry/Cal3_S2.cpp:44-51). K = fx fy s u0 v0 */
static int project_cal3s2(const double pose[12], const double K[5], const double p[3], double pi[2],
                          double Dpose[12], double Dpoint[6]) {
  double pn[2], Dp0[12], Dq0[6];
  if (!project2(pose, p, pn, Dpose ? Dp0 : 0, Dpoint ? Dq0 : 0)) return 0;
  pi[0] = K[0] * pn[0] + K[2] * pn[1] + K[3];
  pi[1] = K[1] * pn[1] + K[4];
  const double Dpi_pn[4] = {K[0], K[2], 0.0, K[1]};
  if (Dpose) m22_mul(Dpi_pn, Dp0, 6, Dpose);
  if (Dpoint) m22_mul(Dpi_pn, Dq0, 3, Dpoint);
  return 1;
}

/* PinholeCamera<Cal3Bundler>::project2 (gtsam/geometry/PinholeCamera.h:230-247)
 * with Cal3Bundler::uncalibrate (gtsam/geometry/Cal3Bundler.cpp:66-92).
 * cam = pose(12) f k1 k2 u0 v0.  Dcam 2x9 row-major = [Dpose | Dcal]. */
static int project_bundler(const double cam[17], const double p[3], double pi[2], double Dcam[18],
                           double Dpoint[6]) {
  double pn[2], Dp0[12], Dq0[6];
  if (!project2(cam, p, pn, Dcam ? Dp0 : 0, Dpoint ? Dq0 : 0)) return 0;
  const double f = cam[12], k1 = cam[13], k2 = cam[14], u0 = cam[15], v0 = cam[16];
  const double x = pn[0], y = pn[1];
  const double r = x * x + y * y;
  const double g = 1. + (k1 + k2 * r) * r;
  const double u = g * x, v = g * y;
  pi[0] = u0 + f * u;
  pi[1] = v0 + f * v;
  if (Dcam || Dpoint) {
    const double a = 2. * (k1 + 2. * k2 * r);
    const double axx = a * x * x, axy = a * x * y, ayy = a * y * y;
    const double Dp[4] = {(g + axx) * f, axy * f, axy * f, (g + ayy) * f};
    if (Dcam) {
      double Dpose[12];
      m22_mul(Dp, Dp0, 6, Dpose);
      const double rx = r * x, ry = r * y;
      for (int j = 0; j < 6; j++) {
        Dcam[j] = Dpose[j];
        Dcam[9 + j] = Dpose[6 + j];
      }
      Dcam[6] = u; Dcam[7] = f * rx; Dcam[8] = f * r * rx;
      Dcam[15] = v; Dcam[16] = f * ry; Dcam[17] = f * r * ry;
    }
    if (Dpoint) m22_mul(Dp, Dq0, 3, Dpoint);
  }
  return 1;
}

/* ------------------------------------------------------------------------ */
/* dense partial Cholesky                                                    */
/* ------------------------------------------------------------------------ */
/* gtsam::choleskyPartial, gtsam/base/cholesky.cpp:107-158.  Column-major
 * n x n, upper triangle.  Eigen LLT<Upper> fails when a pivot is <= 0
 * (Eigen/src/Cholesky/LLT.h llt_inplace::unblocked). */
int orc_cholesky_partial(double* M, int64_t n, int64_t nF) {
#define AT(i, j) M[(i) + (j) * n]
  if (nF == 0) return 1;
  for (int64_t k = 0; k < nF; k++) {
    double x = AT(k, k);
    for (int64_t p = 0; p < k; p++) x -= AT(p, k) * AT(p, k);
    if (x <= 0.0) return 0;
    x = sqrt(x);
    AT(k, k) = x;
    /* row k of R (and of S): R(k,j) = (A(k,j) - sum_p R(p,k) R(p,j)) / x */
    for (int64_t j = k + 1; j < n; j++) {
      double s = AT(k, j);
      for (int64_t p = 0; p < k; p++) s -= AT(p, k) * AT(p, j);
      AT(k, j) = s / x;
    }
  }
  /* C -= S' S on the upper triangle */
  for (int64_t j = nF; j < n; j++)
    for (int64_t i = nF; i <= j; i++) {
      double s = 0;
      for (int64_t p = 0; p < nF; p++) s += AT(p, i) * AT(p, j);
      AT(i, j) -= s;
    }
  /* underconstrained check on the last two pivots (:144-157) */
  if (nF >= 2) {
    int e2, e1;
    (void)frexp(AT(nF - 2, nF - 2), &e2);
    (void)frexp(AT(nF - 1, nF - 1), &e1);
    return (e2 - e1 < 12);
  } else {
    int e1;
    (void)frexp(AT(0, 0), &e1);
    return (e1 > -12);
  }
#undef AT
}

/* ------------------------------------------------------------------------ */
/* problem container                                                         */
/* ------------------------------------------------------------------------ */
typedef struct {
  int64_t* d;
  int64_t n, cap;
} ivec;
static void iv_push(ivec* v, int64_t x) {
  if (v->n == v->cap) {
    v->cap = v->cap ? 2 * v->cap : 4;
    v->d = (int64_t*)realloc(v->d, (size_t)v->cap * sizeof(int64_t));
  }
  v->d[v->n++] = x;
}
static void iv_append(ivec* v, const ivec* o) {
  for (int64_t i = 0; i < o->n; i++) iv_push(v, o->d[i]);
}
static void iv_free(ivec* v) {
  free(v->d);
  v->d = 0;
  v->n = v->cap = 0;
}

typedef struct {
  int32_t type, noise_kind, per_factor;
  int64_t count, graph_index0;
  int64_t* keys;
  double* meas;
  double* noise;
  int noise_size;
  int32_t* cal_index;
  int has_body;
  int robust_kind;
  double robust_param;
  double body[12]; /* body_P_sensor of the group (projection factors) */
  int d, ncols; /* rows, n1+n2+1 */
  int arity;    /* keys per factor (typed groups: F_ARITY; JacobianFactor groups: any, <= B200_JACOBIAN_MAX_ARITY) */
  double* J;    /* count * d * ncols, factor-major col-major [A1 A2 b] */
} ogroup;

struct orc_problem {
  int64_t nvars;
  int linear;        /* created by orc_linear_create: JacobianFactor groups, no Values */
  int jac_f32;       /* the "FP32 linearize + FP64 solve" mode: the whitened [A|b] are rounded to float after linearize */
  int32_t* var_dim;  /* tangent dimension of every variable */
  int32_t* var_type;
  int64_t *val_off, *dof_off;
  double *values, *new_values, *delta;
  int64_t *ordering, *pos;
  int64_t ncal;
  double* cal;
  int64_t ngroups;
  ogroup* groups;
  int64_t nfactors;
  int32_t* fgroup; /* graph position -> group */
  int64_t* fidx;   /* graph position -> index in group */
  /* symbolic */
  int64_t ncliques;
  int64_t *front_ptr, *front_vars, *sep_ptr, *sep_vars, *parent;
  int64_t *cf_ptr, *cf_list; /* factors (graph positions) of each clique, reference order */
  int64_t *ch_ptr, *ch_list; /* children cliques, reference order */
  int64_t* clique_of_var;
  /* numeric */
  int64_t* cond_off;
  double* cond;
};

static int noise_payload(int kind, int d) {
  switch (kind) {
    case B200_NOISE_UNIT: return 0;
    case B200_NOISE_ISOTROPIC: return 1;
    case B200_NOISE_DIAGONAL: return d;
    case B200_NOISE_GAUSSIAN: return d * d;
  }
  return -1;
}

static int64_t cmp_i64(const void* a, const void* b) {
  const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return (x > y) - (x < y);
}
static int cmp_i64_int(const void* a, const void* b) { return (int)cmp_i64(a, b); }

static void factor_keys(const orc_problem* p, int64_t gi, const int64_t** keys, int* arity) {
  const ogroup* g = &p->groups[p->fgroup[gi]];
  *arity = g->arity;
  *keys = g->keys + p->fidx[gi] * (*arity);
}

/* ---- symbolic phase ------------------------------------------------------
 * VariableIndex (gtsam/inference/VariableIndex-inl.h:27-50), EliminationTree
 * (gtsam/inference/EliminationTree-inst.h:77-155), JunctionTree with the
 * child-merge rule (gtsam/inference/JunctionTree-inst.h:63-119) and the
 * cluster merge bookkeeping (gtsam/inference/ClusterTree-inst.h:46-95). */
static void symbolic(orc_problem* p) {
  const int64_t n = p->nvars, m = p->nfactors;
  /* VariableIndex: per variable, factor positions ascending */
  ivec* vi = (ivec*)calloc((size_t)n, sizeof(ivec));
  for (int64_t i = 0; i < m; i++) {
    const int64_t* keys;
    int ar;
    factor_keys(p, i, &keys, &ar);
    for (int a = 0; a < ar; a++) iv_push(&vi[keys[a]], i);
  }
  /* elimination tree */
  const int64_t none = -1;
  int64_t* parents = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  int64_t* prevCol = (int64_t*)malloc((size_t)(m ? m : 1) * sizeof(int64_t));
  ivec* echildren = (ivec*)calloc((size_t)n, sizeof(ivec));
  ivec* nfactors = (ivec*)calloc((size_t)n, sizeof(ivec));
  for (int64_t j = 0; j < n; j++) parents[j] = none;
  for (int64_t i = 0; i < m; i++) prevCol[i] = none;
  for (int64_t j = 0; j < n; j++) {
    const ivec* fs = &vi[p->ordering[j]];
    for (int64_t q = 0; q < fs->n; q++) {
      const int64_t i = fs->d[q];
      if (prevCol[i] != none) {
        int64_t r = prevCol[i];
        while (parents[r] != none) r = parents[r];
        if (r != j) {
          parents[r] = j;
          iv_push(&echildren[j], r);
        }
      } else {
        iv_push(&nfactors[j], i);
      }
      prevCol[i] = j;
    }
  }
  /* symbolic elimination per etree node: sep(j) as ascending positions */
  ivec* sep = (ivec*)calloc((size_t)n, sizeof(ivec));
  int64_t* mark = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  for (int64_t j = 0; j < n; j++) mark[j] = -1;
  /* clusters */
  ivec* cfront = (ivec*)calloc((size_t)n, sizeof(ivec));
  ivec* cfact = (ivec*)calloc((size_t)n, sizeof(ivec));
  ivec* cchild = (ivec*)calloc((size_t)n, sizeof(ivec));
  char* alive = (char*)malloc((size_t)n);
  for (int64_t j = 0; j < n; j++) {
    /* own factors' variables + children separators, minus j */
    ivec* s = &sep[j];
    for (int64_t q = 0; q < nfactors[j].n; q++) {
      const int64_t* keys;
      int ar;
      factor_keys(p, nfactors[j].d[q], &keys, &ar);
      for (int a = 0; a < ar; a++) {
        const int64_t pj = p->pos[keys[a]];
        if (pj != j && mark[pj] != j) { mark[pj] = j; iv_push(s, pj); }
      }
    }
    for (int64_t c = 0; c < echildren[j].n; c++) {
      const ivec* cs = &sep[echildren[j].d[c]];
      for (int64_t q = 0; q < cs->n; q++) {
        const int64_t pj = cs->d[q];
        if (pj != j && mark[pj] != j) { mark[pj] = j; iv_push(s, pj); }
      }
    }
    /* cluster for this node */
    alive[j] = 1;
    iv_push(&cfront[j], j);
    iv_append(&cfact[j], &nfactors[j]);
    /* merge rule: JunctionTree-inst.h:98-118 */
    const int64_t myNrParents = s->n;
    int64_t myNrFrontals = 1;
    const int64_t nch = echildren[j].n;
    char* merge = (char*)calloc((size_t)(nch ? nch : 1), 1);
    for (int64_t c = 0; c < nch; c++) {
      const int64_t ch = echildren[j].d[c];
      if (myNrParents + myNrFrontals == sep[ch].n) {
        myNrFrontals += cfront[ch].n;
        merge[c] = 1;
      }
    }
    /* Cluster::mergeChildren, ClusterTree-inst.h:58-95 */
    for (int64_t c = 0; c < nch; c++) {
      const int64_t ch = echildren[j].d[c];
      if (merge[c]) {
        for (int64_t q = cfront[ch].n - 1; q >= 0; q--) iv_push(&cfront[j], cfront[ch].d[q]);
        iv_append(&cfact[j], &cfact[ch]);
        iv_append(&cchild[j], &cchild[ch]);
        alive[ch] = 0;
      } else {
        iv_push(&cchild[j], ch);
      }
    }
    for (int64_t a = 0, b = cfront[j].n - 1; a < b; a++, b--) {
      const int64_t t = cfront[j].d[a];
      cfront[j].d[a] = cfront[j].d[b];
      cfront[j].d[b] = t;
    }
    free(merge);
  }
  /* number cliques by ascending head position */
  int64_t* cid = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  int64_t nc = 0, nfv = 0, nsv = 0, ncf = 0, nch = 0;
  for (int64_t j = 0; j < n; j++) {
    cid[j] = -1;
    if (alive[j]) {
      cid[j] = nc++;
      nfv += cfront[j].n; nsv += sep[j].n; ncf += cfact[j].n; nch += cchild[j].n;
    }
  }
  p->ncliques = nc;
  p->front_ptr = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
  p->sep_ptr = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
  p->cf_ptr = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
  p->ch_ptr = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
  p->front_vars = (int64_t*)malloc((size_t)(nfv ? nfv : 1) * sizeof(int64_t));
  p->sep_vars = (int64_t*)malloc((size_t)(nsv ? nsv : 1) * sizeof(int64_t));
  p->cf_list = (int64_t*)malloc((size_t)(ncf ? ncf : 1) * sizeof(int64_t));
  p->ch_list = (int64_t*)malloc((size_t)(nch ? nch : 1) * sizeof(int64_t));
  p->parent = (int64_t*)malloc((size_t)(nc ? nc : 1) * sizeof(int64_t));
  p->clique_of_var = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  int64_t c = 0;
  for (int64_t j = 0; j < n; j++) {
    if (!alive[j]) continue;
    for (int64_t q = 0; q < cfront[j].n; q++) {
      const int64_t v = p->ordering[cfront[j].d[q]];
      p->front_vars[p->front_ptr[c] + q] = v;
      p->clique_of_var[v] = c;
    }
    p->front_ptr[c + 1] = p->front_ptr[c] + cfront[j].n;
    /* separator keys sorted by id (== sorted by Key): gtsam/linear/Scatter.cpp:69-72 */
    for (int64_t q = 0; q < sep[j].n; q++) p->sep_vars[p->sep_ptr[c] + q] = p->ordering[sep[j].d[q]];
    qsort(p->sep_vars + p->sep_ptr[c], (size_t)sep[j].n, sizeof(int64_t), cmp_i64_int);
    p->sep_ptr[c + 1] = p->sep_ptr[c] + sep[j].n;
    memcpy(p->cf_list + p->cf_ptr[c], cfact[j].d, (size_t)cfact[j].n * sizeof(int64_t));
    p->cf_ptr[c + 1] = p->cf_ptr[c] + cfact[j].n;
    for (int64_t q = 0; q < cchild[j].n; q++) p->ch_list[p->ch_ptr[c] + q] = cid[cchild[j].d[q]];
    p->ch_ptr[c + 1] = p->ch_ptr[c] + cchild[j].n;
    c++;
  }
  for (c = 0; c < nc; c++) p->parent[c] = -1;
  for (c = 0; c < nc; c++)
    for (int64_t q = p->ch_ptr[c]; q < p->ch_ptr[c + 1]; q++) p->parent[p->ch_list[q]] = c;
  /* conditional storage */
  p->cond_off = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
  for (c = 0; c < nc; c++) {
    int64_t f = 0, s = 0;
    for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++) f += p->var_dim[p->front_vars[q]];
    for (int64_t q = p->sep_ptr[c]; q < p->sep_ptr[c + 1]; q++) s += p->var_dim[p->sep_vars[q]];
    p->cond_off[c + 1] = p->cond_off[c] + f * (f + s + 1);
  }
  p->cond = (double*)calloc((size_t)(p->cond_off[nc] ? p->cond_off[nc] : 1), sizeof(double));
  for (int64_t j = 0; j < n; j++) {
    iv_free(&vi[j]); iv_free(&echildren[j]); iv_free(&nfactors[j]); iv_free(&sep[j]);
    iv_free(&cfront[j]); iv_free(&cfact[j]); iv_free(&cchild[j]);
  }
  free(vi); free(echildren); free(nfactors); free(sep); free(cfront); free(cfact); free(cchild);
  free(parents); free(prevCol); free(mark); free(alive); free(cid);
}

int orc_problem_create(const b200_problem_desc* desc, orc_problem** out) {
  orc_problem* p = (orc_problem*)calloc(1, sizeof(orc_problem));
  const int64_t n = desc->nvars;
  p->nvars = n;
  p->var_type = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  p->var_dim = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  memcpy(p->var_type, desc->var_type, (size_t)n * sizeof(int32_t));
  p->val_off = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  p->dof_off = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  for (int64_t v = 0; v < n; v++) {
    if (p->var_type[v] < 0 || p->var_type[v] >= B200_NUM_VAR_TYPES) return B200_INVALID_ARGUMENT;
    p->val_off[v + 1] = p->val_off[v] + VAR_STORAGE[p->var_type[v]];
    p->var_dim[v] = VAR_DIM[p->var_type[v]];
    p->dof_off[v + 1] = p->dof_off[v] + p->var_dim[v];
  }
  p->values = (double*)malloc((size_t)p->val_off[n] * sizeof(double));
  p->new_values = (double*)malloc((size_t)p->val_off[n] * sizeof(double));
  memcpy(p->values, desc->values, (size_t)p->val_off[n] * sizeof(double));
  p->delta = (double*)calloc((size_t)p->dof_off[n], sizeof(double));
  p->ordering = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  p->pos = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  memcpy(p->ordering, desc->ordering, (size_t)n * sizeof(int64_t));
  for (int64_t v = 0; v < n; v++) p->pos[v] = -1;
  for (int64_t j = 0; j < n; j++) {
    if (p->ordering[j] < 0 || p->ordering[j] >= n || p->pos[p->ordering[j]] != -1) return B200_INVALID_ARGUMENT;
    p->pos[p->ordering[j]] = j;
  }
  p->ncal = desc->ncal;
  p->cal = (double*)malloc((size_t)(desc->ncal ? desc->ncal : 1) * 5 * sizeof(double));
  if (desc->ncal) memcpy(p->cal, desc->cal, (size_t)desc->ncal * 5 * sizeof(double));
  p->ngroups = desc->ngroups;
  p->groups = (ogroup*)calloc((size_t)(desc->ngroups ? desc->ngroups : 1), sizeof(ogroup));
  int64_t next = 0, total = 0;
  for (int64_t g = 0; g < desc->ngroups; g++) total += desc->groups[g].count;
  p->nfactors = total;
  p->fgroup = (int32_t*)malloc((size_t)(total ? total : 1) * sizeof(int32_t));
  p->fidx = (int64_t*)malloc((size_t)(total ? total : 1) * sizeof(int64_t));
  for (int64_t i = 0; i < total; i++) p->fgroup[i] = -1;
  for (int64_t g = 0; g < desc->ngroups; g++) {
    const b200_factor_group* s = &desc->groups[g];
    ogroup* o = &p->groups[g];
    if (s->type < 0 || s->type >= B200_NUM_FACTOR_TYPES) return B200_UNSUPPORTED_FACTOR;
    o->type = s->type; o->noise_kind = s->noise_kind; o->per_factor = s->noise_per_factor;
    o->robust_kind = s->robust_kind; o->robust_param = s->robust_param;
    if (s->robust_kind < 0 || s->robust_kind > B200_ROBUST_FAIR || (s->robust_kind && s->type == B200_FACTOR_SFM_BUNDLER)) return B200_UNSUPPORTED_NOISE;
    o->count = s->count;
    o->graph_index0 = s->graph_index ? -1 : (s->graph_index0 < 0 ? next : s->graph_index0);
    if (!s->graph_index) next = o->graph_index0 + s->count;
    const int ar = F_ARITY[s->type], ms = F_MEAS[s->type], d = F_DIM[s->type];
    o->d = d;
    o->arity = ar;
    o->ncols = VAR_DIM[F_VT[s->type][0]] + (ar == 2 ? VAR_DIM[F_VT[s->type][1]] : 0) + 1;
    o->noise_size = noise_payload(s->noise_kind, d);
    if (o->noise_size < 0) return B200_UNSUPPORTED_NOISE;
    o->keys = (int64_t*)malloc((size_t)(s->count * ar + 1) * sizeof(int64_t));
    memcpy(o->keys, s->keys, (size_t)(s->count * ar) * sizeof(int64_t));
    o->meas = (double*)malloc((size_t)(s->count * ms + 1) * sizeof(double));
    memcpy(o->meas, s->meas, (size_t)(s->count * ms) * sizeof(double));
    const int64_t nn = (int64_t)o->noise_size * (s->noise_per_factor ? s->count : 1);
    o->noise = (double*)malloc((size_t)(nn + 1) * sizeof(double));
    if (nn) memcpy(o->noise, s->noise, (size_t)nn * sizeof(double));
    if (s->type == B200_FACTOR_PROJECTION_CAL3S2) {
      o->cal_index = (int32_t*)calloc((size_t)s->count + 1, sizeof(int32_t));
      if (s->cal_index) memcpy(o->cal_index, s->cal_index, (size_t)s->count * sizeof(int32_t));
      if (s->body_P_sensor) { o->has_body = 1; memcpy(o->body, s->body_P_sensor, sizeof o->body); }
    }
    o->J = (double*)calloc((size_t)(s->count * d * o->ncols + 1), sizeof(double));
    for (int64_t i = 0; i < s->count; i++) {
      const int64_t gi = s->graph_index ? s->graph_index[i] : o->graph_index0 + i;
      if (gi < 0 || gi >= total || p->fgroup[gi] != -1) return B200_INVALID_ARGUMENT;
      p->fgroup[gi] = (int32_t)g;
      p->fidx[gi] = i;
      for (int a = 0; a < ar; a++) {
        const int64_t k = o->keys[i * ar + a];
        if (k < 0 || k >= n || p->var_type[k] != F_VT[s->type][a]) return B200_INVALID_ARGUMENT;
      }
    }
  }
  symbolic(p);
  *out = p;
  return B200_OK;
}

/* ---- GaussianFactorGraph level: a graph of JacobianFactors of any arity and block widths
 * (gtsam/linear/JacobianFactor.h:93-103) solved by GaussianFactorGraph::optimize(ordering,
 * EliminatePreferCholesky) (gtsam/linear/GaussianFactorGraph.cpp:316-319).  The factors are stored
 * whitened, as JacobianFactor::updateHessian uses them (JacobianFactor.cpp:563-598 -> whiten() :743-750,
 * noiseModel::Diagonal::WhitenInPlace: row r times invsigmas[r] = 1/sigmas[r]). */
static void whiten_jacobian_group(ogroup* o, const double* Ab, const double* sigmas) {
  const int64_t per = (int64_t)o->d * o->ncols;
  for (int64_t i = 0; i < o->count; i++)
    for (int64_t e = 0; e < per; e++) {
      double x = Ab[i * per + e];
      if (sigmas) x *= 1.0 / sigmas[i * o->d + e % o->d];
      o->J[i * per + e] = x;
    }
}

int orc_linear_create(const b200_linear_desc* desc, orc_problem** out) {
  orc_problem* p = (orc_problem*)calloc(1, sizeof(orc_problem));
  const int64_t n = desc->nvars;
  p->nvars = n;
  p->linear = 1;
  p->var_type = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  p->var_dim = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  p->val_off = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  p->dof_off = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  for (int64_t v = 0; v < n; v++) {
    if (desc->var_dim[v] < 1) return B200_INVALID_ARGUMENT;
    p->var_dim[v] = desc->var_dim[v];
    p->dof_off[v + 1] = p->dof_off[v] + p->var_dim[v];
  }
  p->values = (double*)malloc(sizeof(double));
  p->new_values = (double*)malloc(sizeof(double));
  p->delta = (double*)calloc((size_t)p->dof_off[n] + 1, sizeof(double));
  p->ordering = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
  p->pos = (int64_t*)malloc((size_t)(n + 1) * sizeof(int64_t));
  memcpy(p->ordering, desc->ordering, (size_t)n * sizeof(int64_t));
  for (int64_t v = 0; v < n; v++) p->pos[v] = -1;
  for (int64_t j = 0; j < n; j++) {
    if (p->ordering[j] < 0 || p->ordering[j] >= n || p->pos[p->ordering[j]] != -1) return B200_INVALID_ARGUMENT;
    p->pos[p->ordering[j]] = j;
  }
  p->cal = (double*)malloc(5 * sizeof(double));
  p->ngroups = desc->ngroups;
  p->groups = (ogroup*)calloc((size_t)(desc->ngroups ? desc->ngroups : 1), sizeof(ogroup));
  p->ngroups = desc->ngroups + desc->nhgroups;
  free(p->groups);
  p->groups = (ogroup*)calloc((size_t)(p->ngroups ? p->ngroups : 1), sizeof(ogroup));
  int64_t next = 0, total = 0;
  for (int64_t g = 0; g < desc->ngroups; g++) total += desc->groups[g].count;
  for (int64_t g = 0; g < desc->nhgroups; g++) total += desc->hgroups[g].count;
  p->nfactors = total;
  p->fgroup = (int32_t*)malloc((size_t)(total ? total : 1) * sizeof(int32_t));
  p->fidx = (int64_t*)malloc((size_t)(total ? total : 1) * sizeof(int64_t));
  for (int64_t i = 0; i < total; i++) p->fgroup[i] = -1;
  for (int64_t hg = 0; hg < desc->nhgroups; hg++) { /* HessianFactor groups are stored after the Jacobian groups */
    const b200_hessian_group* s = &desc->hgroups[hg];
    const int64_t g = desc->ngroups + hg;
    ogroup* o = &p->groups[g];
    if (s->arity < 1 || s->arity > B200_JACOBIAN_MAX_ARITY) return B200_UNSUPPORTED_FACTOR;
    o->type = B200_FACTOR_HESSIAN;
    o->count = s->count;
    o->arity = s->arity;
    o->ncols = 1;
    for (int a = 0; a < s->arity; a++) o->ncols += s->dims[a];
    o->d = o->ncols;
    o->keys = (int64_t*)malloc((size_t)(s->count * s->arity + 1) * sizeof(int64_t));
    memcpy(o->keys, s->keys, (size_t)(s->count * s->arity) * sizeof(int64_t));
    o->J = (double*)calloc((size_t)(s->count * o->d * o->ncols + 1), sizeof(double));
    memcpy(o->J, s->info, (size_t)(s->count * o->d * o->ncols) * sizeof(double));
  }
  for (int64_t g = 0; g < desc->ngroups; g++) {
    const b200_jacobian_group* s = &desc->groups[g];
    ogroup* o = &p->groups[g];
    if (s->arity < 1 || s->arity > B200_JACOBIAN_MAX_ARITY) return B200_UNSUPPORTED_FACTOR;
    o->type = B200_FACTOR_JACOBIAN;
    o->count = s->count;
    o->graph_index0 = s->graph_index ? -1 : (s->graph_index0 < 0 ? next : s->graph_index0);
    if (!s->graph_index) next = o->graph_index0 + s->count;
    o->d = s->rows;
    o->arity = s->arity;
    o->ncols = 1;
    for (int a = 0; a < s->arity; a++) o->ncols += s->dims[a];
    o->keys = (int64_t*)malloc((size_t)(s->count * s->arity + 1) * sizeof(int64_t));
    memcpy(o->keys, s->keys, (size_t)(s->count * s->arity) * sizeof(int64_t));
    o->meas = 0; o->noise = 0; o->cal_index = 0;
    o->J = (double*)calloc((size_t)(s->count * o->d * o->ncols + 1), sizeof(double));
    if (s->sigmas)
      for (int64_t i = 0; i < s->count * s->rows; i++)
        if (!(s->sigmas[i] > 0)) return B200_UNSUPPORTED_NOISE; /* Constrained: needs QR */
    whiten_jacobian_group(o, s->Ab, s->sigmas);
    for (int64_t i = 0; i < s->count; i++) {
      const int64_t gi = s->graph_index ? s->graph_index[i] : o->graph_index0 + i;
      if (gi < 0 || gi >= total || p->fgroup[gi] != -1) return B200_INVALID_ARGUMENT;
      p->fgroup[gi] = (int32_t)g;
      p->fidx[gi] = i;
      for (int a = 0; a < s->arity; a++) {
        const int64_t k = o->keys[i * s->arity + a];
        if (k < 0 || k >= n || p->var_dim[k] != s->dims[a]) return B200_INVALID_ARGUMENT;
      }
    }
  }
  for (int64_t hg = 0; hg < desc->nhgroups; hg++) {
    const b200_hessian_group* s = &desc->hgroups[hg];
    const int64_t g = desc->ngroups + hg;
    ogroup* o = &p->groups[g];
    o->graph_index0 = s->graph_index ? -1 : (s->graph_index0 < 0 ? next : s->graph_index0);
    if (!s->graph_index) next = o->graph_index0 + s->count;
    for (int64_t i = 0; i < s->count; i++) {
      const int64_t gi = s->graph_index ? s->graph_index[i] : o->graph_index0 + i;
      if (gi < 0 || gi >= total || p->fgroup[gi] != -1) return B200_INVALID_ARGUMENT;
      p->fgroup[gi] = (int32_t)g;
      p->fidx[gi] = i;
      for (int a = 0; a < s->arity; a++) {
        const int64_t k = o->keys[i * s->arity + a];
        if (k < 0 || k >= n || p->var_dim[k] != s->dims[a]) return B200_INVALID_ARGUMENT;
      }
    }
  }
  symbolic(p);
  *out = p;
  return B200_OK;
}

int orc_linear_update_hessian(orc_problem* p, int64_t hgroup, const double* info) {
  int64_t k = 0;
  for (int64_t g = 0; g < p->ngroups; g++)
    if (p->groups[g].type == B200_FACTOR_HESSIAN && k++ == hgroup) {
      ogroup* o = &p->groups[g];
      memcpy(o->J, info, (size_t)(o->count * o->d * o->ncols) * sizeof(double));
      return B200_OK;
    }
  return B200_INVALID_ARGUMENT;
}

int orc_linear_update(orc_problem* p, int64_t group, const double* Ab, const double* sigmas) {
  if (!p->linear || group < 0 || group >= p->ngroups) return B200_INVALID_ARGUMENT;
  whiten_jacobian_group(&p->groups[group], Ab, sigmas);
  return B200_OK;
}

void orc_problem_destroy(orc_problem* p) {
  if (!p) return;
  for (int64_t g = 0; g < p->ngroups; g++) {
    ogroup* o = &p->groups[g];
    free(o->keys); free(o->meas); free(o->noise); free(o->cal_index); free(o->J);
  }
  free(p->groups); free(p->var_type); free(p->var_dim); free(p->val_off); free(p->dof_off); free(p->values);
  free(p->new_values); free(p->delta); free(p->ordering); free(p->pos); free(p->cal);
  free(p->fgroup); free(p->fidx); free(p->front_ptr); free(p->front_vars); free(p->sep_ptr);
  free(p->sep_vars); free(p->parent); free(p->cf_ptr); free(p->cf_list); free(p->ch_ptr);
  free(p->ch_list); free(p->clique_of_var); free(p->cond_off); free(p->cond);
  free(p);
}

void orc_set_values(orc_problem* p, const double* v) { memcpy(p->values, v, (size_t)p->val_off[p->nvars] * sizeof(double)); }
void orc_get_values(const orc_problem* p, double* v) { memcpy(v, p->values, (size_t)p->val_off[p->nvars] * sizeof(double)); }
int64_t orc_values_size(const orc_problem* p) { return p->val_off[p->nvars]; }
int64_t orc_delta_size(const orc_problem* p) { return p->dof_off[p->nvars]; }

/* ------------------------------------------------------------------------ */
/* per-factor unwhitened residual r and Jacobians (row-major d x n_i)        */
/* ------------------------------------------------------------------------ */
/* ---- Pose2 = (x, y, theta), tangent (x, y, theta) ---------------------------------------------
 * between(a, b) = a^-1 b: Pose2::inverse (gtsam/geometry/Pose2.cpp:201-203) then operator*
 * (Pose2.h:131-133; Rot2::operator* goes through fromCosSin -> normalize, Rot2.cpp:27-30,56-64);
 * theta() = atan2(s, c) (Rot2.h:186-188). */
static void rot2_normalize(double* c, double* s) {
  double scale = (*c) * (*c) + (*s) * (*s);
  if (fabs(scale - 1.0) > 1e-10) { scale = 1 / sqrt(scale); *c *= scale; *s *= scale; }
}
/* out = (x, y, c, s) of a^-1 b */
static void pose2_between(const double a[3], const double b[3], double out[4]) {
  const double ca = cos(a[2]), sa = sin(a[2]), cb = cos(b[2]), sb = sin(b[2]);
  /* a^-1 = (R_a^T, unrotate(-t_a)) */
  const double ix = ca * (-a[0]) + sa * (-a[1]), iy = -sa * (-a[0]) + ca * (-a[1]);
  double c = ca * cb - (-sa) * sb, s = (-sa) * cb + ca * sb;
  rot2_normalize(&c, &s);
  out[0] = ix + (ca * b[0] + sa * b[1]);   /* t = t_inv + R_a^T t_b */
  out[1] = iy + (-sa * b[0] + ca * b[1]);
  out[2] = c; out[3] = s;
}

/* Returns r (d) and, if H1, the Jacobians.  `active` semantics: all factors
 * active.  Follows NoiseModelFactorN::unwhitenedError -> evaluateError. */
static void eval_factor(const orc_problem* p, const ogroup* g, int64_t i, const double* values,
                        double* r, double* H1, double* H2) {
  const int ar = g->arity;
  const int64_t* keys = g->keys + i * ar;
  const double* x1 = values + p->val_off[keys[0]];
  const double* x2 = ar == 2 ? values + p->val_off[keys[1]] : 0;
  const double* z = g->meas + i * F_MEAS[g->type];
  switch (g->type) {
    case B200_FACTOR_BETWEEN_POSE3: {
      /* gtsam/slam/BetweenFactor.h:111-124 (fast Jacobian variant) */
      double hx[12], hinv[12], zh[12];
      pose3_between(x1, x2, hx);
      pose3_between(z, hx, zh); /* Local(measured, hx) = Logmap(measured^-1 hx) */
      orc_pose3_logmap(zh, r);
      if (H1) {
        orc_pose3_inverse(hx, hinv);
        orc_pose3_adjoint_map(hinv, H1);
        for (int k = 0; k < 36; k++) H1[k] = -H1[k];
        memset(H2, 0, 36 * sizeof(double));
        for (int k = 0; k < 6; k++) H2[7 * k] = 1.0;
      }
      break;
    }
    case B200_FACTOR_PRIOR_POSE3: {
      /* gtsam/nonlinear/PriorFactor.h:98-102: -Local(x, prior), H = I */
      pose3_local(x1, z, r);
      for (int k = 0; k < 6; k++) r[k] = -r[k];
      if (H1) {
        memset(H1, 0, 36 * sizeof(double));
        for (int k = 0; k < 6; k++) H1[7 * k] = 1.0;
      }
      break;
    }
    case B200_FACTOR_BETWEEN_POSE2: {
      /* gtsam/slam/BetweenFactor.h:111-124 (fast variant): hx = between(p1, p2) with H1 = -AdjointMap(hx^-1), H2 = I
         (gtsam/base/Lie.h:63-69, Pose2::AdjointMap gtsam/geometry/Pose2.cpp:127-135); r = Local(measured, hx) =
         (x, y, theta) of measured^-1 hx (Pose2::ChartAtOrigin::Local, Pose2.cpp:111-122) */
      double hx[4], d[4];
      pose2_between(x1, x2, hx);
      const double hxp[3] = {hx[0], hx[1], atan2(hx[3], hx[2])};
      pose2_between(z, hxp, d);
      r[0] = d[0]; r[1] = d[1]; r[2] = atan2(d[3], d[2]);
      if (H1) {
        /* hx^-1 = (R^T, unrotate(-t)): c' = c, s' = -s, (x', y') = R^T (-t) */
        const double c = hx[2], s = hx[3];
        const double xi = c * (-hx[0]) + s * (-hx[1]), yi = -s * (-hx[0]) + c * (-hx[1]);
        const double Ad[9] = {c, s, yi, -s, c, -xi, 0, 0, 1};   /* AdjointMap of (c, -s, xi, yi): [[c', -s', y'],[s', c', -x'],[0,0,1]] */
        for (int k = 0; k < 9; k++) H1[k] = -Ad[k];
        memset(H2, 0, 9 * sizeof(double));
        for (int k = 0; k < 3; k++) H2[4 * k] = 1.0;
      }
      break;
    }
    case B200_FACTOR_PRIOR_POSE2: {
      /* gtsam/nonlinear/PriorFactor.h:98-102: -Local(x, prior), H = I */
      double d[4];
      pose2_between(x1, z, d);
      r[0] = -d[0]; r[1] = -d[1]; r[2] = -atan2(d[3], d[2]);
      if (H1) {
        memset(H1, 0, 9 * sizeof(double));
        for (int k = 0; k < 3; k++) H1[4 * k] = 1.0;
      }
      break;
    }
    case B200_FACTOR_PRIOR_POINT3: {
      for (int k = 0; k < 3; k++) r[k] = -(z[k] - x1[k]);
      if (H1) {
        memset(H1, 0, 9 * sizeof(double));
        for (int k = 0; k < 3; k++) H1[4 * k] = 1.0;
      }
      break;
    }
    case B200_FACTOR_PRIOR_CAM_BUNDLER: {
      /* PinholeCamera::localCoordinates, gtsam/geometry/PinholeCamera.h:208-213;
         Cal3Bundler::localCoordinates = T2.vector() - vector() */
      pose3_local(x1, z, r);
      for (int k = 0; k < 3; k++) r[6 + k] = z[12 + k] - x1[12 + k];
      for (int k = 0; k < 9; k++) r[k] = -r[k];
      if (H1) {
        memset(H1, 0, 81 * sizeof(double));
        for (int k = 0; k < 9; k++) H1[10 * k] = 1.0;
      }
      break;
    }
    case B200_FACTOR_PROJECTION_CAL3S2: {
      /* gtsam/slam/ProjectionFactor.h:138-166 */
      const double* K = p->cal + 5 * (g->cal_index ? g->cal_index[i] : 0);
      double pi[2], cam[12];
      const double* cpose = x1;
      if (g->has_body) { /* camera = pose.compose(body_P_sensor, H0), ProjectionFactor.h:141-151 */
        orc_pose3_compose(x1, g->body, cam);
        cpose = cam;
      }
      if (project_cal3s2(cpose, K, x2, pi, H1, H2)) {
        r[0] = pi[0] - z[0];
        r[1] = pi[1] - z[1];
        if (g->has_body && H1) { /* H1 = H1 * H0, H0 = body_P_sensor.inverse().AdjointMap() (Lie.h compose) */
          double Si[12], Ad[36], T[12];
          orc_pose3_inverse(g->body, Si);
          orc_pose3_adjoint_map(Si, Ad);
          for (int rr = 0; rr < 2; rr++)
            for (int cc = 0; cc < 6; cc++) {
              double sacc = 0;
              for (int k = 0; k < 6; k++) sacc += H1[rr * 6 + k] * Ad[k * 6 + cc];
              T[rr * 6 + cc] = sacc;
            }
          memcpy(H1, T, sizeof T);
        }
      } else { /* cheirality: zero Jacobians, constant residual 2*fx */
        if (H1) { memset(H1, 0, 12 * sizeof(double)); memset(H2, 0, 6 * sizeof(double)); }
        r[0] = r[1] = 2.0 * K[0];
      }
      break;
    }
    case B200_FACTOR_SFM_BUNDLER: {
      /* gtsam/slam/GeneralSFMFactor.h:127-168 */
      double pi[2];
      if (project_bundler(x1, x2, pi, H1, H2)) {
        r[0] = pi[0] - z[0];
        r[1] = pi[1] - z[1];
      } else { /* cheirality: H = 0, b = 0 */
        if (H1) { memset(H1, 0, 18 * sizeof(double)); memset(H2, 0, 6 * sizeof(double)); }
        r[0] = r[1] = 0.0;
      }
      break;
    }
  }
}

/* noise whitening: Unit / Isotropic (NoiseModel.cpp:646-675) / Diagonal
 * (:322-340) / Gaussian (:163-238).  M is row-major d x ncols. */
static void whiten_rows(const ogroup* g, int64_t i, double* M, int ncols) {
  const int d = g->d;
  const double* nz = g->noise + (g->per_factor ? i * g->noise_size : 0);
  switch (g->noise_kind) {
    case B200_NOISE_UNIT: break;
    case B200_NOISE_ISOTROPIC: {
      const double inv = 1.0 / nz[0];
      for (int k = 0; k < d * ncols; k++) M[k] *= inv;
      break;
    }
    case B200_NOISE_DIAGONAL:
      for (int rr = 0; rr < d; rr++) {
        const double inv = 1.0 / nz[rr];
        for (int c = 0; c < ncols; c++) M[rr * ncols + c] *= inv;
      }
      break;
    case B200_NOISE_GAUSSIAN: {
      double T[9 * 10];
      for (int rr = 0; rr < d; rr++)
        for (int c = 0; c < ncols; c++) {
          double s = 0;
          for (int k = 0; k < d; k++) s += nz[rr * d + k] * M[k * ncols + c];
          T[rr * ncols + c] = s;
        }
      memcpy(M, T, (size_t)(d * ncols) * sizeof(double));
      break;
    }
  }
}

/* m-estimators, gtsam/linear/LossFunctions.cpp: weight(distance) and loss(distance) */
static double robust_weight(int kind, double k, double distance) {
  const double a = fabs(distance);
  switch (kind) {
    case B200_ROBUST_HUBER: return (a <= k) ? 1.0 : (k / a);
    case B200_ROBUST_CAUCHY: return (k * k) / (k * k + distance * distance);
    case B200_ROBUST_TUKEY: {
      if (a <= k) { const double t = 1.0 - distance * distance / (k * k); return t * t; }
      return 0.0;
    }
    case B200_ROBUST_FAIR: return 1.0 / (1.0 + a / k);
  }
  return 1.0;
}
static double robust_loss(int kind, double k, double distance) {
  const double a = fabs(distance);
  switch (kind) {
    case B200_ROBUST_HUBER: return (a <= k) ? distance * distance / 2 : k * (a - (k / 2));
    case B200_ROBUST_CAUCHY: return k * k * log1p(distance * distance / (k * k)) * 0.5;
    case B200_ROBUST_TUKEY: {
      if (a <= k) { const double u = 1.0 - distance * distance / (k * k); return k * k * (1 - u * u * u) / 6.0; }
      return k * k / 6.0;
    }
    case B200_ROBUST_FAIR: { const double ne = a / k; return k * k * (ne - log1p(ne)); }
  }
  return 0.5 * distance * distance;
}

/* NonlinearFactorGraph::error, gtsam/nonlinear/NonlinearFactorGraph.cpp:170-179;
 * NoiseModelFactor::error, gtsam/nonlinear/NonlinearFactor.cpp:133-146 */
static double graph_error(const orc_problem* p, const double* values) {
  double total = 0.0;
  for (int64_t gi = 0; gi < p->nfactors; gi++) {
    const ogroup* g = &p->groups[p->fgroup[gi]];
    double r[9];
    eval_factor(p, g, p->fidx[gi], values, r, 0, 0);
    whiten_rows(g, p->fidx[gi], r, 1);
    double s = 0;
    for (int k = 0; k < g->d; k++) s += r[k] * r[k];
    /* Gaussian: 0.5 d^2; Robust::loss(d^2) = rho(sqrt(d^2)) (NoiseModel.h Robust::loss) */
    total += g->robust_kind ? robust_loss(g->robust_kind, g->robust_param, sqrt(s)) : 0.5 * s;
  }
  return total;
}
double orc_error(orc_problem* p) { return graph_error(p, p->values); }

/* NoiseModelFactor::linearize, gtsam/nonlinear/NonlinearFactor.cpp:150-182;
 * GeneralSFMFactor::linearize, gtsam/slam/GeneralSFMFactor.h:141-177 */
void orc_linearize(orc_problem* p) {
  if (p->linear) return; /* a linear problem is its own linearization */
  for (int64_t gidx = 0; gidx < p->ngroups; gidx++) {
    ogroup* g = &p->groups[gidx];
    const int d = g->d, ar = g->arity;
    const int n1 = VAR_DIM[F_VT[g->type][0]], n2 = ar == 2 ? VAR_DIM[F_VT[g->type][1]] : 0;
    for (int64_t i = 0; i < g->count; i++) {
      double r[9], H1[81], H2[36];
      eval_factor(p, g, i, p->values, r, H1, H2);
      for (int k = 0; k < d; k++) r[k] = -r[k]; /* b = -error */
      whiten_rows(g, i, H1, n1);
      if (n2) whiten_rows(g, i, H2, n2);
      whiten_rows(g, i, r, 1);
      if (g->robust_kind) { /* Robust::WhitenSystem: Base::reweight, Block mode (LossFunctions.cpp:79-106) */
        double nrm = 0;
        for (int k = 0; k < d; k++) nrm += r[k] * r[k];
        const double w = sqrt(robust_weight(g->robust_kind, g->robust_param, sqrt(nrm)));
        for (int k = 0; k < d * n1; k++) H1[k] *= w;
        for (int k = 0; k < d * n2; k++) H2[k] *= w;
        for (int k = 0; k < d; k++) r[k] *= w;
      }
      double* J = g->J + i * d * g->ncols;
      for (int rr = 0; rr < d; rr++) {
        for (int c = 0; c < n1; c++) J[rr + c * d] = H1[rr * n1 + c];
        for (int c = 0; c < n2; c++) J[rr + (n1 + c) * d] = H2[rr * n2 + c];
        J[rr + (n1 + n2) * d] = r[rr];
      }
      if (p->jac_f32) /* storage precision of the device's FP32 mode: evaluated in FP64, kept as floats */
        for (int e = 0; e < d * g->ncols; e++) J[e] = (double)(float)J[e];
    }
  }
}

void orc_set_jacobian_fp32(orc_problem* p, int on) { p->jac_f32 = on != 0; }

void orc_get_jacobians(const orc_problem* p, int64_t group, double* out) {
  const ogroup* g = &p->groups[group];
  memcpy(out, g->J, (size_t)(g->count * g->d * g->ncols) * sizeof(double));
}

/* GaussianFactorGraph::hessianDiagonal, gtsam/linear/GaussianFactorGraph.cpp:279-287;
 * JacobianFactor::hessianDiagonalAdd, gtsam/linear/JacobianFactor.cpp:516-541 */
void orc_hessian_diagonal(const orc_problem* p, double* out) {
  memset(out, 0, (size_t)p->dof_off[p->nvars] * sizeof(double));
  for (int64_t gi = 0; gi < p->nfactors; gi++) {
    const ogroup* g = &p->groups[p->fgroup[gi]];
    const int64_t i = p->fidx[gi];
    const int d = g->d, ar = g->arity;
    const double* J = g->J + i * d * g->ncols;
    int col = 0;
    for (int a = 0; a < ar; a++) {
      const int64_t v = g->keys[i * ar + a];
      const int nv = p->var_dim[v];
      for (int c = 0; c < nv; c++, col++) {
        double s = 0;
        if (g->type == B200_FACTOR_HESSIAN) s = J[col + col * d]; /* HessianFactor::hessianDiagonalAdd, HessianFactor.cpp:292-304 */
        else for (int rr = 0; rr < d; rr++) s += J[rr + col * d] * J[rr + col * d];
        out[p->dof_off[v] + c] += s;
      }
    }
  }
}

/* GaussianFactorGraph::error(x), gtsam/linear/GaussianFactorGraph.cpp:71-78;
 * JacobianFactor::error, gtsam/linear/JacobianFactor.cpp:486-491 */
static double linear_error(const orc_problem* p, const double* delta) {
  double total = 0;
  for (int64_t gi = 0; gi < p->nfactors; gi++) {
    const ogroup* g = &p->groups[p->fgroup[gi]];
    const int64_t i = p->fidx[gi];
    const int d = g->d, ar = g->arity;
    const double* J = g->J + i * d * g->ncols;
    if (g->type == B200_FACTOR_HESSIAN) { /* HessianFactor::error, HessianFactor.cpp:331-346: 0.5 (f - 2 x'g + x'G x) */
      const int N = d - 1;
      double x[N > 0 ? N : 1];
      int col = 0;
      for (int a = 0; a < ar; a++) {
        const int64_t v = g->keys[i * ar + a];
        for (int c = 0; c < p->var_dim[v]; c++, col++) x[col] = delta ? delta[p->dof_off[v] + c] : 0.0;
      }
      double xg = 0, xGx = 0;
      for (int r = 0; r < N; r++) {
        xg += x[r] * J[r + N * d];
        for (int c = 0; c < N; c++) xGx += x[r] * x[c] * (r <= c ? J[r + c * d] : J[c + r * d]);
      }
      total += 0.5 * (J[N + N * d] - 2.0 * xg + xGx);
      continue;
    }
    double e[d > 0 ? d : 1];
    for (int rr = 0; rr < d; rr++) e[rr] = -J[rr + (g->ncols - 1) * d];
    int col = 0;
    for (int a = 0; a < ar; a++) {
      const int64_t v = g->keys[i * ar + a];
      const int nv = p->var_dim[v];
      for (int c = 0; c < nv; c++, col++) {
        const double x = delta ? delta[p->dof_off[v] + c] : 0.0;
        for (int rr = 0; rr < d; rr++) e[rr] += J[rr + col * d] * x;
      }
    }
    double s = 0;
    for (int rr = 0; rr < d; rr++) s += e[rr] * e[rr];
    total += 0.5 * s;
  }
  return total;
}

/* ------------------------------------------------------------------------ */
/* multifrontal elimination + back-substitution                              */
/* ------------------------------------------------------------------------ */
/* EliminateCholesky (gtsam/linear/HessianFactor.cpp:516-536): Scatter, sum of
 * updateHessian (JacobianFactor.cpp:563-598, HessianFactor.cpp:348-374),
 * choleskyPartial, split.  Driver: ClusterTree-inst.h:218-265 (post-order). */
int orc_solve(orc_problem* p, double lambda, int diagonal_damping, double min_diagonal,
              double max_diagonal, double* lin_err0, double* lin_err_delta, int64_t* fail_var) {
  const int64_t nc = p->ncliques;
  double* hdiag = 0;
  if (lambda > 0 && diagonal_damping) {
    hdiag = (double*)malloc((size_t)p->dof_off[p->nvars] * sizeof(double));
    orc_hessian_diagonal(p, hdiag);
  }
  double** schur = (double**)calloc((size_t)nc, sizeof(double*)); /* (s+1)^2 col-major, upper */
  int64_t* slot = (int64_t*)malloc((size_t)p->nvars * sizeof(int64_t));
  int status = B200_OK;
  if (fail_var) *fail_var = -1;
  for (int64_t c = 0; c < nc && status == B200_OK; c++) {
    int64_t f = 0, s = 0;
    for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++) {
      slot[p->front_vars[q]] = f;
      f += p->var_dim[p->front_vars[q]];
    }
    for (int64_t q = p->sep_ptr[c]; q < p->sep_ptr[c + 1]; q++) {
      slot[p->sep_vars[q]] = f + s;
      s += p->var_dim[p->sep_vars[q]];
    }
    const int64_t n = f + s + 1;
    double* M = (double*)calloc((size_t)(n * n), sizeof(double));
#define MM(i, j) M[(i) + (j) * n]
    /* own factors */
    for (int64_t q = p->cf_ptr[c]; q < p->cf_ptr[c + 1]; q++) {
      const int64_t gi = p->cf_list[q];
      const ogroup* g = &p->groups[p->fgroup[gi]];
      const int64_t i = p->fidx[gi];
      const int d = g->d, ar = g->arity;
      const double* J = g->J + i * d * g->ncols;
      int64_t off[B200_JACOBIAN_MAX_ARITY + 1];
      int dim[B200_JACOBIAN_MAX_ARITY + 1], col0[B200_JACOBIAN_MAX_ARITY + 1];
      int col = 0;
      for (int a = 0; a < ar; a++) {
        const int64_t v = g->keys[i * ar + a];
        off[a] = slot[v];
        dim[a] = p->var_dim[v];
        col0[a] = col;
        col += dim[a];
      }
      off[ar] = n - 1; dim[ar] = 1; col0[ar] = col; /* rhs column b */
      for (int a = 0; a <= ar; a++)
        for (int b = a; b <= ar; b++)
          for (int ca = 0; ca < dim[a]; ca++)
            for (int cb = 0; cb < dim[b]; cb++) {
              double sum = 0;
              if (g->type == B200_FACTOR_HESSIAN) /* HessianFactor::updateHessian, HessianFactor.cpp:348-374: upper blocks of info */
                sum = J[(col0[a] + ca) + (col0[b] + cb) * d];
              else
                for (int rr = 0; rr < d; rr++) sum += J[rr + (col0[a] + ca) * d] * J[rr + (col0[b] + cb) * d];
              int64_t I = off[a] + ca, Jx = off[b] + cb;
              if (a == b && ca > cb) continue; /* diagonal block: upper only */
              if (I > Jx) { const int64_t t = I; I = Jx; Jx = t; }
              MM(I, Jx) += sum;
            }
    }
    /* damping priors (LevenbergMarquardtState.h:125-156): lambda*I or lambda*clip(diag H) */
    if (lambda > 0) {
      for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++) {
        const int64_t v = p->front_vars[q];
        const int nv = p->var_dim[v];
        for (int k = 0; k < nv; k++) {
          double a2 = 1.0;
          if (diagonal_damping) {
            double h = hdiag[p->dof_off[v] + k];
            h = h > min_diagonal ? h : min_diagonal;
            h = h < max_diagonal ? h : max_diagonal;
            const double sq = sqrt(h);
            a2 = sq * sq;
          }
          /* whitened A = sqrt(lambda) * a: contribution lambda * a^2 */
          const double sl = 1.0 / (1.0 / sqrt(lambda)); /* 1/sigma with sigma = 1/sqrt(lambda) */
          MM(slot[v] + k, slot[v] + k) += (sl * sl) * a2;
        }
      }
    }
    /* children Schur complements (extend-add) */
    for (int64_t q = p->ch_ptr[c]; q < p->ch_ptr[c + 1]; q++) {
      const int64_t ch = p->ch_list[q];
      int64_t cs = 0;
      for (int64_t qq = p->sep_ptr[ch]; qq < p->sep_ptr[ch + 1]; qq++) cs += p->var_dim[p->sep_vars[qq]];
      const int64_t cn = cs + 1;
      int64_t* map = (int64_t*)malloc((size_t)cn * sizeof(int64_t));
      int64_t k = 0;
      for (int64_t qq = p->sep_ptr[ch]; qq < p->sep_ptr[ch + 1]; qq++) {
        const int64_t v = p->sep_vars[qq];
        for (int t = 0; t < p->var_dim[v]; t++) map[k++] = slot[v] + t;
      }
      map[k] = n - 1;
      const double* S = schur[ch];
      for (int64_t jj = 0; jj < cn; jj++)
        for (int64_t ii = 0; ii <= jj; ii++) {
          int64_t I = map[ii], Jx = map[jj];
          if (I > Jx) { const int64_t t = I; I = Jx; Jx = t; }
          MM(I, Jx) += S[ii + jj * cn];
        }
      free(map);
      free(schur[ch]);
      schur[ch] = 0;
    }
    if (!orc_cholesky_partial(M, n, f)) {
      status = B200_INDETERMINATE;
      if (fail_var) *fail_var = p->front_vars[p->front_ptr[c]];
    } else {
      double* C = p->cond + p->cond_off[c];
      for (int64_t j = 0; j < n; j++)
        for (int64_t i = 0; i < f; i++) C[i + j * f] = (i <= j) ? MM(i, j) : 0.0;
      const int64_t cn = s + 1;
      schur[c] = (double*)malloc((size_t)(cn * cn) * sizeof(double));
      for (int64_t j = 0; j < cn; j++)
        for (int64_t i = 0; i < cn; i++) schur[c][i + j * cn] = (i <= j) ? MM(f + i, f + j) : 0.0;
    }
#undef MM
    free(M);
  }
  for (int64_t c = 0; c < nc; c++) free(schur[c]);
  free(schur);
  free(hdiag);
  if (status != B200_OK) { free(slot); return status; }
  /* back-substitution, pre-order: gtsam/linear/linearAlgorithms-inst.h:50-117 */
  for (int64_t c = nc - 1; c >= 0 && status == B200_OK; c--) {
    int64_t f = 0, s = 0;
    for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++) f += p->var_dim[p->front_vars[q]];
    for (int64_t q = p->sep_ptr[c]; q < p->sep_ptr[c + 1]; q++) s += p->var_dim[p->sep_vars[q]];
    const int64_t n = f + s + 1;
    const double* C = p->cond + p->cond_off[c];
    double* x = (double*)malloc((size_t)f * sizeof(double));
    for (int64_t i = 0; i < f; i++) x[i] = C[i + (n - 1) * f];
    int64_t col = f;
    for (int64_t q = p->sep_ptr[c]; q < p->sep_ptr[c + 1]; q++) {
      const int64_t v = p->sep_vars[q];
      for (int t = 0; t < p->var_dim[v]; t++, col++) {
        const double xs = p->delta[p->dof_off[v] + t];
        for (int64_t i = 0; i < f; i++) x[i] -= C[i + col * f] * xs;
      }
    }
    for (int64_t i = f - 1; i >= 0; i--) {
      double sum = x[i];
      for (int64_t j = i + 1; j < f; j++) sum -= C[i + j * f] * x[j];
      x[i] = sum / C[i + i * f];
    }
    int64_t k = 0;
    for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++) {
      const int64_t v = p->front_vars[q];
      for (int t = 0; t < p->var_dim[v]; t++, k++) {
        if (isnan(x[k])) {
          status = B200_INDETERMINATE;
          if (fail_var) *fail_var = p->front_vars[p->front_ptr[c]];
        }
        p->delta[p->dof_off[v] + t] = x[k];
      }
    }
    free(x);
  }
  free(slot);
  if (status == B200_OK) {
    if (lin_err0) *lin_err0 = linear_error(p, 0);
    if (lin_err_delta) *lin_err_delta = linear_error(p, p->delta);
  }
  return status;
}

/* ------------------------------------------------------------------------ */
/* Marginals (groundwork for the device path).  The reference computes
 * Marginals::marginalCovariance(j) = marginalInformation(j)^-1 with marginalInformation taken from
 * bayesTree_.marginalFactor(j, EliminatePreferCholesky) (gtsam/nonlinear/Marginals.cpp:118-154,
 * gtsam/inference/BayesTree-inst.h:287-318).  The marginal information of x_j under the Gaussian
 * N(H^-1 g, H^-1) of the linearised graph is the inverse of the (j, j) block of H^-1, so the
 * covariance is that block itself: the columns H^-1 e_k, k in dofs(j), obtained from the multifrontal
 * factor H = U^T U whose rows [R S] are the stored conditionals (forward solve U^T y = e_k in
 * elimination order, back-substitution U x = y as linearAlgorithms-inst.h:50-117). */
static void clique_fs(const orc_problem* p, int64_t c, int64_t* f, int64_t* s) {
  *f = 0; *s = 0;
  for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++) *f += p->var_dim[p->front_vars[q]];
  for (int64_t q = p->sep_ptr[c]; q < p->sep_ptr[c + 1]; q++) *s += p->var_dim[p->sep_vars[q]];
}

/* x = H^-1 g with the factorisation left by the last successful orc_solve (its damping included) */
void orc_solve_rhs(const orc_problem* p, const double* g, double* x) {
  const int64_t ntot = p->dof_off[p->nvars], nc = p->ncliques;
  double* y = (double*)malloc((size_t)ntot * sizeof(double));
  memcpy(y, g, (size_t)ntot * sizeof(double));
  int64_t* idx = (int64_t*)malloc((size_t)(ntot + 1) * sizeof(int64_t));
  for (int64_t c = 0; c < nc; c++) {   /* U^T y = g */
    int64_t f, s;
    clique_fs(p, c, &f, &s);
    const double* C = p->cond + p->cond_off[c];
    int64_t k = 0;
    for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++)
      for (int t = 0; t < p->var_dim[p->front_vars[q]]; t++) idx[k++] = p->dof_off[p->front_vars[q]] + t;
    for (int64_t q = p->sep_ptr[c]; q < p->sep_ptr[c + 1]; q++)
      for (int t = 0; t < p->var_dim[p->sep_vars[q]]; t++) idx[k++] = p->dof_off[p->sep_vars[q]] + t;
    for (int64_t i = 0; i < f; i++) {
      double sum = y[idx[i]];
      for (int64_t kk = 0; kk < i; kk++) sum -= C[kk + i * f] * y[idx[kk]];
      y[idx[i]] = sum / C[i + i * f];
    }
    for (int64_t col = f; col < f + s; col++) {
      double sum = 0;
      for (int64_t i = 0; i < f; i++) sum += C[i + col * f] * y[idx[i]];
      y[idx[col]] -= sum;
    }
  }
  for (int64_t c = nc - 1; c >= 0; c--) {   /* U x = y */
    int64_t f, s;
    clique_fs(p, c, &f, &s);
    const double* C = p->cond + p->cond_off[c];
    int64_t k = 0;
    for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++)
      for (int t = 0; t < p->var_dim[p->front_vars[q]]; t++) idx[k++] = p->dof_off[p->front_vars[q]] + t;
    for (int64_t q = p->sep_ptr[c]; q < p->sep_ptr[c + 1]; q++)
      for (int t = 0; t < p->var_dim[p->sep_vars[q]]; t++) idx[k++] = p->dof_off[p->sep_vars[q]] + t;
    for (int64_t i = f - 1; i >= 0; i--) {
      double sum = y[idx[i]];
      for (int64_t col = f; col < f + s; col++) sum -= C[i + col * f] * x[idx[col]];
      for (int64_t j = i + 1; j < f; j++) sum -= C[i + j * f] * x[idx[j]];
      x[idx[i]] = sum / C[i + i * f];
    }
  }
  free(idx);
  free(y);
}

/* out: d x d column-major covariance of variable `var` at the current values */
int orc_marginal_covariance(orc_problem* p, int64_t var, double* out) {
  const int64_t ntot = p->dof_off[p->nvars];
  const int d = p->var_dim[var];
  orc_linearize(p);
  int64_t fv;
  const int st = orc_solve(p, 0.0, 0, 0, 0, 0, 0, &fv);
  if (st != B200_OK) return st;
  double* g = (double*)calloc((size_t)ntot, sizeof(double));
  double* x = (double*)calloc((size_t)ntot, sizeof(double));
  for (int k = 0; k < d; k++) {
    g[p->dof_off[var] + k] = 1.0;
    orc_solve_rhs(p, g, x);
    g[p->dof_off[var] + k] = 0.0;
    for (int i = 0; i < d; i++) out[i + k * d] = x[p->dof_off[var] + i];
  }
  free(g);
  free(x);
  return B200_OK;
}

/* Marginals::jointMarginalCovariance (gtsam/nonlinear/Marginals.cpp:128-190): the covariance of several
 * variables, blocks in sorted-key order = the rows/columns of H^-1 picked at those variables' dofs.
 * vars must be sorted ascending and distinct; out: D x D column-major, D = sum of their dims. */
int orc_joint_marginal_covariance(orc_problem* p, const int64_t* vars, int64_t nv, double* out) {
  const int64_t ntot = p->dof_off[p->nvars];
  int64_t D = 0;
  for (int64_t a = 0; a < nv; a++) D += p->var_dim[vars[a]];
  orc_linearize(p);
  int64_t fv;
  const int st = orc_solve(p, 0.0, 0, 0, 0, 0, 0, &fv);
  if (st != B200_OK) return st;
  double* g = (double*)calloc((size_t)ntot, sizeof(double));
  double* x = (double*)calloc((size_t)ntot, sizeof(double));
  int64_t col = 0;
  for (int64_t a = 0; a < nv; a++)
    for (int k = 0; k < p->var_dim[vars[a]]; k++, col++) {
      g[p->dof_off[vars[a]] + k] = 1.0;
      orc_solve_rhs(p, g, x);
      g[p->dof_off[vars[a]] + k] = 0.0;
      int64_t row = 0;
      for (int64_t b = 0; b < nv; b++)
        for (int i = 0; i < p->var_dim[vars[b]]; i++, row++) out[row + col * D] = x[p->dof_off[vars[b]] + i];
    }
  free(g);
  free(x);
  return B200_OK;
}

void orc_get_delta(const orc_problem* p, double* out) { memcpy(out, p->delta, (size_t)p->dof_off[p->nvars] * sizeof(double)); }

void orc_get_conditional(const orc_problem* p, int64_t c, double* out) {
  memcpy(out, p->cond + p->cond_off[c], (size_t)(p->cond_off[c + 1] - p->cond_off[c]) * sizeof(double));
}

/* Values::retract, gtsam/nonlinear/Values.cpp:52-63 */
static void retract_all(const orc_problem* p, const double* values, const double* delta, double* out) {
  for (int64_t v = 0; v < p->nvars; v++) {
    const double* x = values + p->val_off[v];
    const double* d = delta + p->dof_off[v];
    double* y = out + p->val_off[v];
    switch (p->var_type[v]) {
      case B200_VAR_POSE3: pose3_retract(x, d, y); break;
      case B200_VAR_POINT3: for (int k = 0; k < 3; k++) y[k] = x[k] + d[k]; break;
      case B200_VAR_POSE2: {
        /* x * ChartAtOrigin::Retract(d) = x * Pose2(d0, d1, d2) (gtsam/geometry/Pose2.cpp:99-109, Pose2.h:131-133) */
        const double c = cos(x[2]), s = sin(x[2]), cd = cos(d[2]), sd = sin(d[2]);
        double cn = c * cd - s * sd, sn = s * cd + c * sd;
        rot2_normalize(&cn, &sn);
        y[0] = x[0] + (c * d[0] - s * d[1]);
        y[1] = x[1] + (s * d[0] + c * d[1]);
        y[2] = atan2(sn, cn);
        break;
      }
      case B200_VAR_CAM_BUNDLER:
        /* PinholeCamera::retract, gtsam/geometry/PinholeCamera.h:199-205;
           Cal3Bundler::retract: (f,k1,k2) + d, u0 v0 kept */
        pose3_retract(x, d, y);
        for (int k = 0; k < 3; k++) y[12 + k] = x[12 + k] + d[6 + k];
        y[15] = x[15]; y[16] = x[16];
        break;
    }
  }
}
double orc_try_step(orc_problem* p) {
  retract_all(p, p->values, p->delta, p->new_values);
  return graph_error(p, p->new_values);
}
void orc_accept_step(orc_problem* p) { memcpy(p->values, p->new_values, (size_t)p->val_off[p->nvars] * sizeof(double)); }

/* ------------------------------------------------------------------------ */
/* LM / GN control                                                           */
/* ------------------------------------------------------------------------ */
void orc_lm_init(orc_lm* lm, orc_problem* p, const b200_lm_params* params) {
  lm->prob = p;
  lm->params = *params;
  lm->state.error = orc_error(p);
  lm->state.lambda = params->lambda_initial;
  lm->state.current_factor = params->lambda_factor;
  lm->state.iterations = 0;
  lm->state.total_inner_iterations = 0;
}

/* LevenbergMarquardtOptimizer::tryLambda, .cpp:121-270. Returns 1 when the
 * lambda search for this outer iteration is finished. */
static int try_lambda(orc_lm* lm) {
  orc_problem* p = lm->prob;
  const b200_lm_params* P = &lm->params;
  b200_lm_state* S = &lm->state;
  double modelFidelity = 0.0, newError = INFINITY, costChange = 0.0;
  int step_is_successful = 0, stopSearchingLambda = 0;
  double e0 = 0, e1 = 0;
  int64_t fv;
  const int ok = orc_solve(p, S->lambda, P->diagonal_damping, P->min_diagonal, P->max_diagonal, &e0, &e1, &fv) == B200_OK;
  if (ok) {
    const double linearizedCostChange = e0 - e1;
    if (linearizedCostChange >= 0) {
      newError = orc_try_step(p);
      costChange = S->error - newError;
      if (linearizedCostChange > DBL_EPSILON * e0) {
        modelFidelity = costChange / linearizedCostChange;
        step_is_successful = modelFidelity > P->min_model_fidelity;
      }
      const double minAbsoluteTolerance = P->relative_error_tol * S->error;
      if (fabs(costChange) < minAbsoluteTolerance) stopSearchingLambda = 1;
    }
  }
  if (step_is_successful) {
    /* decreaseLambda, internal/LevenbergMarquardtState.h:81-94 */
    double newLambda = S->lambda, newFactor = S->current_factor;
    if (P->use_fixed_lambda_factor) {
      newLambda /= S->current_factor;
    } else {
      const double t = 1.0 - pow(2.0 * modelFidelity - 1.0, 3);
      newLambda *= (1.0 / 3.0 > t ? 1.0 / 3.0 : t);
      newFactor = 2.0 * S->current_factor;
    }
    newLambda = P->lambda_lower_bound > newLambda ? P->lambda_lower_bound : newLambda;
    orc_accept_step(p);
    S->error = newError;
    S->lambda = newLambda;
    S->current_factor = newFactor;
    S->iterations += 1;
    S->total_inner_iterations += 1;
    return 1;
  } else if (!stopSearchingLambda) {
    /* increaseLambda, :70-76 */
    S->lambda *= S->current_factor;
    S->total_inner_iterations += 1;
    if (!P->use_fixed_lambda_factor) S->current_factor *= 2.0;
    return S->lambda >= P->lambda_upper_bound ? 1 : 0;
  } else {
    return 1;
  }
}

/* LevenbergMarquardtOptimizer::iterate, .cpp:273-308 */
int orc_lm_iterate(orc_lm* lm) {
  orc_linearize(lm->prob);
  while (!try_lambda(lm)) {
  }
  return B200_OK;
}

/* checkConvergence, gtsam/nonlinear/NonlinearOptimizer.cpp:182-231 */
static int check_convergence(double rel, double absT, double errT, double cur, double nw) {
  if (nw <= errT) return 1;
  const double absoluteDecrease = cur - nw;
  const double relativeDecrease = absoluteDecrease / cur;
  return (rel && (relativeDecrease <= rel)) || (absoluteDecrease <= absT);
}

/* NonlinearOptimizer::defaultOptimize, gtsam/nonlinear/NonlinearOptimizer.cpp:62-117 */
int orc_lm_optimize(orc_lm* lm) {
  const b200_lm_params* P = &lm->params;
  double currentError = lm->state.error;
  if (currentError <= P->error_tol) return B200_OK;
  if (lm->state.iterations >= P->max_iterations) return B200_OK;
  double newError = currentError;
  do {
    currentError = newError;
    orc_lm_iterate(lm);
    newError = lm->state.error;
  } while (lm->state.iterations < P->max_iterations &&
           !check_convergence(P->relative_error_tol, P->absolute_error_tol, P->error_tol, currentError, newError) &&
           isfinite(currentError));
  return B200_OK;
}

/* GaussNewtonOptimizer::iterate, gtsam/nonlinear/GaussNewtonOptimizer.cpp:44-67 */
int orc_gn_iterate(orc_problem* p, double* new_error) {
  orc_linearize(p);
  int64_t fv;
  const int st = orc_solve(p, 0.0, 0, 0, 0, 0, 0, &fv);
  if (st != B200_OK) return st;
  const double e = orc_try_step(p);
  orc_accept_step(p);
  if (new_error) *new_error = e;
  return B200_OK;
}

/* ------------------------------------------------------------------------ */
/* Dogleg: DoglegOptimizer::iterate, gtsam/nonlinear/DoglegOptimizer.cpp:84-121 with
 * DoglegOptimizerImpl::Iterate (ONE_STEP_PER_ITERATION), gtsam/nonlinear/DoglegOptimizerImpl.h:139-258,
 * ComputeDoglegPoint / ComputeBlend, DoglegOptimizerImpl.cpp:25-98, and
 * GaussianFactorGraph::optimizeGradientSearch, gtsam/linear/GaussianFactorGraph.cpp:381-407.
 * The Bayes tree and the linear graph encode the same quadratic up to a constant, so gradient,
 * |R g|^2 = |A g|^2 and error differences are evaluated on the Jacobian factors. */
static double dot_n(const double* a, const double* b, int64_t n) { double s = 0; for (int64_t i = 0; i < n; i++) s += a[i] * b[i]; return s; }

int orc_dogleg_iterate(orc_problem* p, double* error_io, double* delta_io) {
  const int64_t n = p->dof_off[p->nvars];
  orc_linearize(p);
  int64_t fv;
  double e0 = 0, e1 = 0;
  const int st = orc_solve(p, 0.0, 0, 0, 0, &e0, &e1, &fv);
  if (st != B200_OK) return st;
  double* dxn = (double*)malloc((size_t)n * sizeof(double));
  double* dxu = (double*)calloc((size_t)n, sizeof(double));
  double* dxd = (double*)malloc((size_t)n * sizeof(double));
  memcpy(dxn, p->delta, (size_t)n * sizeof(double));
  /* gradientAtZero = -A^T b */
  for (int64_t gi = 0; gi < p->nfactors; gi++) {
    const ogroup* g = &p->groups[p->fgroup[gi]];
    const int64_t i = p->fidx[gi];
    const int d = g->d, ar = g->arity;
    const double* J = g->J + i * d * g->ncols;
    int col = 0;
    for (int a = 0; a < ar; a++) {
      const int64_t v = g->keys[i * ar + a];
      for (int c = 0; c < p->var_dim[v]; c++, col++) {
        double s = 0;
        for (int rr = 0; rr < d; rr++) s += J[rr + col * d] * J[rr + (g->ncols - 1) * d];
        dxu[p->dof_off[v] + c] -= s;
      }
    }
  }
  const double gg = dot_n(dxu, dxu, n);
  /* |A g|^2 */
  double Ag2 = 0;
  for (int64_t gi = 0; gi < p->nfactors; gi++) {
    const ogroup* g = &p->groups[p->fgroup[gi]];
    const int64_t i = p->fidx[gi];
    const int d = g->d, ar = g->arity;
    const double* J = g->J + i * d * g->ncols;
    double e[9] = {0};
    int col = 0;
    for (int a = 0; a < ar; a++) {
      const int64_t v = g->keys[i * ar + a];
      for (int c = 0; c < p->var_dim[v]; c++, col++)
        for (int rr = 0; rr < d; rr++) e[rr] += J[rr + col * d] * dxu[p->dof_off[v] + c];
    }
    for (int rr = 0; rr < d; rr++) Ag2 += e[rr] * e[rr];
  }
  const double step = -gg / Ag2;
  for (int64_t k = 0; k < n; k++) dxu[k] *= step;
  const double uu = dot_n(dxu, dxu, n), un = dot_n(dxu, dxn, n), nn = dot_n(dxn, dxn, n);
  const double f_error = *error_io, M_error = e0;
  double delta = *delta_io, new_f = f_error;
  int stay = 1, zero_step = 0;
  while (stay) {
    /* ComputeDoglegPoint */
    double ca, cb;
    const double deltaSq = delta * delta;
    if (deltaSq < uu) { ca = sqrt(deltaSq / uu); cb = 0; }
    else if (deltaSq < nn) {
      const double a = uu - 2. * un + nn, b = 2. * (un - uu), c = uu - delta * delta;
      const double sq = sqrt(b * b - 4 * a * c);
      const double tau1 = (-b + sq) / (2. * a), tau2 = (-b - sq) / (2. * a);
      const double tau = (-DBL_EPSILON <= tau1 && tau1 <= 1.0 + DBL_EPSILON) ? tau1 : tau2;
      ca = 1. - tau; cb = tau;
    } else { ca = 0; cb = 1; }
    for (int64_t k = 0; k < n; k++) dxd[k] = (cb == 0 ? ca * dxu[k] : (ca == 0 ? dxn[k] : ca * dxu[k] + cb * dxn[k]));
    memcpy(p->delta, dxd, (size_t)n * sizeof(double));
    new_f = orc_try_step(p);
    const double new_M = linear_error(p, dxd);
    const double rho = (fabs(f_error - new_f) < 1e-15 || fabs(M_error - new_M) < 1e-15) ? 0.5 : (f_error - new_f) / (M_error - new_M);
    if (rho >= 0.75) {
      const double nd = sqrt(dot_n(dxd, dxd, n));
      delta = delta > 3.0 * nd ? delta : 3.0 * nd;
      stay = 0;
    } else if (rho >= 0.25) {
      stay = 0;
    } else if (rho >= 0.0) {
      if (delta > 1e-5) delta = 0.5 * delta;
      stay = 0;   /* ONE_STEP_PER_ITERATION */
    } else {
      if (delta > 1e-5) { delta *= 0.5; stay = 1; }
      else { zero_step = 1; new_f = f_error; stay = 0; }
    }
  }
  if (!zero_step) orc_accept_step(p);
  *error_io = new_f;
  *delta_io = delta;
  free(dxn); free(dxu); free(dxd);
  return B200_OK;
}

/* ------------------------------------------------------------------------ */
void orc_symbolic_info_get(const orc_problem* p, b200_symbolic_info* info) {
  memset(info, 0, sizeof *info);
  info->ncliques = p->ncliques;
  info->total_dim = p->dof_off[p->nvars];
  info->frontal_list_len = p->front_ptr[p->ncliques];
  info->separator_list_len = p->sep_ptr[p->ncliques];
  int64_t* level = (int64_t*)calloc((size_t)(p->ncliques ? p->ncliques : 1), sizeof(int64_t));
  for (int64_t c = 0; c < p->ncliques; c++) {
    int64_t f = 0, s = 0;
    for (int64_t q = p->front_ptr[c]; q < p->front_ptr[c + 1]; q++) f += p->var_dim[p->front_vars[q]];
    for (int64_t q = p->sep_ptr[c]; q < p->sep_ptr[c + 1]; q++) s += p->var_dim[p->sep_vars[q]];
    if (f > info->max_frontal_dim) info->max_frontal_dim = f;
    if (s > info->max_separator_dim) info->max_separator_dim = s;
    info->factor_flops += (double)f * f * f / 3.0 + (double)f * f * s + (double)f * s * s;
    info->front_bytes += (f + s + 1) * (f + s + 1) * 8;
    if (level[c] + 1 > info->nlevels) info->nlevels = level[c] + 1;
    if (p->parent[c] >= 0 && level[p->parent[c]] < level[c] + 1) level[p->parent[c]] = level[c] + 1;
  }
  free(level);
}

void orc_get_cliques(const orc_problem* p, int64_t* frontal_ptr, int64_t* frontal_vars,
                     int64_t* separator_ptr, int64_t* separator_vars, int64_t* parent) {
  memcpy(frontal_ptr, p->front_ptr, (size_t)(p->ncliques + 1) * sizeof(int64_t));
  memcpy(separator_ptr, p->sep_ptr, (size_t)(p->ncliques + 1) * sizeof(int64_t));
  memcpy(frontal_vars, p->front_vars, (size_t)p->front_ptr[p->ncliques] * sizeof(int64_t));
  memcpy(separator_vars, p->sep_vars, (size_t)p->sep_ptr[p->ncliques] * sizeof(int64_t));
  memcpy(parent, p->parent, (size_t)p->ncliques * sizeof(int64_t));
}
