// geometry.cuh — SO(3)/SE(3) maps and pinhole projection with analytic
// Jacobians, as device inline functions (FP64).  Conventions follow the
// reference (file:line relative to borglab/gtsam):
//   * Pose3 = R (row-major 9) + t (3); tangent (omega, v), rotation first;
//   * Rot3 is a plain 3x3 matrix; Expmap/Logmap are the full maps
//     (GTSAM_POSE3_EXPMAP / GTSAM_ROT3_EXPMAP builds).
// Everything is written for one thread working in registers: no arrays with
// dynamic indexing on hot paths beyond small fully-unrolled loops.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace b200 {

struct Mat3 {
  double m[9];  // row-major
};
struct Pose {
  Mat3 R;
  double t[3];
};

__device__ __forceinline__ Mat3 mul(const Mat3& A, const Mat3& B) {
  Mat3 C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
__device__ __forceinline__ Mat3 transpose(const Mat3& A) {
  Mat3 C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * j + i];
  return C;
}
__device__ __forceinline__ void mulv(const Mat3& A, const double* v, double* r) {
  const double a = v[0], b = v[1], c = v[2];
#pragma unroll
  for (int i = 0; i < 3; i++) r[i] = A.m[3 * i] * a + A.m[3 * i + 1] * b + A.m[3 * i + 2] * c;
}
// A^T v
__device__ __forceinline__ void mulTv(const Mat3& A, const double* v, double* r) {
  const double a = v[0], b = v[1], c = v[2];
#pragma unroll
  for (int i = 0; i < 3; i++) r[i] = A.m[i] * a + A.m[3 + i] * b + A.m[6 + i] * c;
}
__device__ __forceinline__ Mat3 hat(double x, double y, double z) {
  Mat3 W;
  W.m[0] = 0; W.m[1] = -z; W.m[2] = y;
  W.m[3] = z; W.m[4] = 0; W.m[5] = -x;
  W.m[6] = -y; W.m[7] = x; W.m[8] = 0;
  return W;
}

// so3::ExpmapFunctor::expmap — gtsam/geometry/SO3.cpp:49-87
__device__ __forceinline__ Mat3 so3_expmap(const double* w) {
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  Mat3 W = hat(w[0], w[1], w[2]);
  Mat3 R;
  if (theta2 <= 2.220446049250313e-16) {
#pragma unroll
    for (int i = 0; i < 9; i++) R.m[i] = W.m[i];
  } else {
    const double theta = sqrt(theta2);
    const double sin_theta = sin(theta);
    const double s2 = sin(theta / 2.0);
    const double one_minus_cos = 2.0 * s2 * s2;
    Mat3 K;
#pragma unroll
    for (int i = 0; i < 9; i++) K.m[i] = W.m[i] / theta;
    Mat3 KK = mul(K, K);
#pragma unroll
    for (int i = 0; i < 9; i++) R.m[i] = sin_theta * K.m[i] + one_minus_cos * KK.m[i];
  }
  R.m[0] += 1.0; R.m[4] += 1.0; R.m[8] += 1.0;
  return R;
}

// SO3::Logmap — gtsam/geometry/SO3.cpp:247-325.  The three near-pi branches are
// one formula under a cyclic relabelling (a,b,c) of the axes; written once here.
__device__ __forceinline__ void so3_logmap(const Mat3& Rm, double* w) {
  const double* R = Rm.m;
  const double tr = R[0] + R[4] + R[8];
  if (tr + 1.0 < 1e-3) {
    // pick the largest diagonal entry as axis a; b, c follow cyclically
    int a;
    if (R[8] > R[4] && R[8] > R[0]) a = 2;
    else if (R[4] > R[0]) a = 1;
    else a = 0;
    const int b = (a + 1) % 3, c = (a + 2) % 3;
    const double W = R[3 * c + b] - R[3 * b + c];
    const double Q1 = 2.0 + 2.0 * R[3 * a + a];
    const double Q2 = R[3 * a + b] + R[3 * b + a];
    const double Q3 = R[3 * c + a] + R[3 * a + c];
    const double r = sqrt(Q1);
    const double one_over_r = 1 / r;
    const double norm = sqrt(Q1 * Q1 + Q2 * Q2 + Q3 * Q3 + W * W);
    const double sgn_w = W < 0 ? -1.0 : 1.0;
    const double mag = 3.14159265358979323846 - (2 * sgn_w * W) / norm;
    const double scale = 0.5 * one_over_r * mag;
    w[a] = sgn_w * scale * Q1;
    w[b] = sgn_w * scale * Q2;
    w[c] = sgn_w * scale * Q3;
  } else {
    double magnitude;
    const double tr_3 = tr - 3.0;
    if (tr_3 < -1e-6) {
      const double theta = acos((tr - 1.0) / 2.0);
      magnitude = theta / (2.0 * sin(theta));
    } else {
      magnitude = 0.5 - tr_3 / 12.0 + tr_3 * tr_3 / 60.0;
    }
    w[0] = magnitude * (R[7] - R[5]);
    w[1] = magnitude * (R[2] - R[6]);
    w[2] = magnitude * (R[3] - R[1]);
  }
}

// Pose3::operator* — gtsam/geometry/Pose3.h:114-116
__device__ __forceinline__ Pose compose(const Pose& A, const Pose& B) {
  Pose C;
  C.R = mul(A.R, B.R);
  double t[3];
  mulv(A.R, B.t, t);
#pragma unroll
  for (int i = 0; i < 3; i++) C.t[i] = A.t[i] + t[i];
  return C;
}
// Pose3::inverse — gtsam/geometry/Pose3.cpp:49-52
__device__ __forceinline__ Pose inverse(const Pose& A) {
  Pose C;
  C.R = transpose(A.R);
  const double nt[3] = {-A.t[0], -A.t[1], -A.t[2]};
  mulv(C.R, nt, C.t);
  return C;
}
// LieGroup::between — gtsam/base/Lie.h:63-69
__device__ __forceinline__ Pose between(const Pose& A, const Pose& B) {
  Pose C;
  Mat3 At = transpose(A.R);
  C.R = mul(At, B.R);
  // inverse(A) * B: t = Rt*(-tA) + Rt*tB, evaluated in the reference's order
  double ti[3], tb[3];
  const double nt[3] = {-A.t[0], -A.t[1], -A.t[2]};
  mulv(At, nt, ti);
  mulv(At, B.t, tb);
#pragma unroll
  for (int i = 0; i < 3; i++) C.t[i] = ti[i] + tb[i];
  return C;
}

// Pose3::Expmap — gtsam/geometry/Pose3.cpp:169-185
__device__ __forceinline__ Pose pose_expmap(const double* xi) {
  Pose T;
  T.R = so3_expmap(xi);
  const double* w = xi;
  const double* v = xi + 3;
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (theta2 > 2.220446049250313e-16) {
    const double wv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
    const double oxv[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
    double Roxv[3];
    mulv(T.R, oxv, Roxv);
#pragma unroll
    for (int i = 0; i < 3; i++) T.t[i] = (oxv[i] - Roxv[i] + w[i] * wv) / theta2;
  } else {
    T.t[0] = v[0]; T.t[1] = v[1]; T.t[2] = v[2];
  }
  return T;
}

// Pose3::Logmap — gtsam/geometry/Pose3.cpp:188-208
__device__ __forceinline__ void pose_logmap(const Pose& T, double* xi) {
  double w[3];
  so3_logmap(T.R, w);
  const double t = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  xi[0] = w[0]; xi[1] = w[1]; xi[2] = w[2];
  if (t < 1e-10) {
    xi[3] = T.t[0]; xi[4] = T.t[1]; xi[5] = T.t[2];
  } else {
    const Mat3 W = hat(w[0] / t, w[1] / t, w[2] / t);
    const double Tan = tan(0.5 * t);
    double WT[3], WWT[3];
    mulv(W, T.t, WT);
    mulv(W, WT, WWT);
#pragma unroll
    for (int i = 0; i < 3; i++) xi[3 + i] = T.t[i] - (0.5 * t) * WT[i] + (1 - t / (2. * Tan)) * WWT[i];
  }
}

// LieGroup::retract — gtsam/base/Lie.h:131-133 (x * Expmap(xi))
__device__ __forceinline__ Pose pose_retract(const Pose& x, const double* xi) { return compose(x, pose_expmap(xi)); }
// LieGroup::localCoordinates — gtsam/base/Lie.h:152-160 (Logmap(x^-1 g))
__device__ __forceinline__ void pose_local(const Pose& x, const Pose& g, double* xi) { pose_logmap(between(x, g), xi); }

__device__ __forceinline__ Pose load_pose(const double* __restrict__ p) {
  Pose T;
#pragma unroll
  for (int i = 0; i < 9; i++) T.R.m[i] = p[i];
  T.t[0] = p[9]; T.t[1] = p[10]; T.t[2] = p[11];
  return T;
}
__device__ __forceinline__ void store_pose(const Pose& T, double* p) {
#pragma unroll
  for (int i = 0; i < 9; i++) p[i] = T.R.m[i];
  p[9] = T.t[0]; p[10] = T.t[1]; p[11] = T.t[2];
}

// PinholeBase::project2 — gtsam/geometry/CalibratedCamera.cpp:116-133 with
// transformTo (Pose3.cpp:371-388), Project (:88-94), Dpose (:27-34), Dpoint
// (:37-46).  Returns false on cheirality failure.  Dpose row-major 2x6, Dpoint 2x3.
__device__ __forceinline__ bool project_normalized(const Pose& T, const double* p, double& u, double& v,
                                                   double* Dpose, double* Dpoint) {
  const double d3[3] = {p[0] - T.t[0], p[1] - T.t[1], p[2] - T.t[2]};
  double q[3];
  mulTv(T.R, d3, q);
  if (q[2] <= 0) return false;
  const double d = 1.0 / q[2];
  u = q[0] * d;
  v = q[1] * d;
  const double uv = u * v, uu = u * u, vv = v * v;
  Dpose[0] = uv; Dpose[1] = -1 - uu; Dpose[2] = v; Dpose[3] = -d; Dpose[4] = 0; Dpose[5] = d * u;
  Dpose[6] = 1 + vv; Dpose[7] = -uv; Dpose[8] = -u; Dpose[9] = 0; Dpose[10] = -d; Dpose[11] = d * v;
  // Rt(i,j) = R(j,i)
#pragma unroll
  for (int j = 0; j < 3; j++) {
    Dpoint[j] = (T.R.m[3 * j] - u * T.R.m[3 * j + 2]) * d;
    Dpoint[3 + j] = (T.R.m[3 * j + 1] - v * T.R.m[3 * j + 2]) * d;
  }
  return true;
}

}  // namespace b200
