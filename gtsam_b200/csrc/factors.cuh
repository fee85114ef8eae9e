// factors.cuh — per-factor residual + Jacobian evaluators and noise whitening
// (device, FP64, one thread per factor, everything in registers).
//
// Each evaluator fills a row-major D x NC register block M = [A1 | A2 | b]
// holding the UNWHITENED Jacobians and b = -r(x), exactly what
// NoiseModelFactor::linearize builds before WhitenSystem
// (gtsam/nonlinear/NonlinearFactor.cpp:150-182).  With WITH_J = false only the
// last column is produced (used by the error kernels, which need r only).
#pragma once
#include "geometry.cuh"
#include "../../include/gtsam_b200.h"

namespace b200 {

template <int TYPE> struct FactorTraits;
template <> struct FactorTraits<B200_FACTOR_BETWEEN_POSE3> { enum { D = 6, N1 = 6, N2 = 6, ARITY = 2, MEAS = 12 }; };
template <> struct FactorTraits<B200_FACTOR_PRIOR_POSE3> { enum { D = 6, N1 = 6, N2 = 0, ARITY = 1, MEAS = 12 }; };
template <> struct FactorTraits<B200_FACTOR_PRIOR_POINT3> { enum { D = 3, N1 = 3, N2 = 0, ARITY = 1, MEAS = 3 }; };
template <> struct FactorTraits<B200_FACTOR_PROJECTION_CAL3S2> { enum { D = 2, N1 = 6, N2 = 3, ARITY = 2, MEAS = 2 }; };
template <> struct FactorTraits<B200_FACTOR_SFM_BUNDLER> { enum { D = 2, N1 = 9, N2 = 3, ARITY = 2, MEAS = 2 }; };
template <> struct FactorTraits<B200_FACTOR_PRIOR_CAM_BUNDLER> { enum { D = 9, N1 = 9, N2 = 0, ARITY = 1, MEAS = 17 }; };
template <> struct FactorTraits<B200_FACTOR_BETWEEN_POSE2> { enum { D = 3, N1 = 3, N2 = 3, ARITY = 2, MEAS = 3 }; };
template <> struct FactorTraits<B200_FACTOR_PRIOR_POSE2> { enum { D = 3, N1 = 3, N2 = 0, ARITY = 1, MEAS = 3 }; };

// Everything a factor evaluator may read.
struct EvalCtx {
  const double* __restrict__ values;   // packed Values
  const int* __restrict__ val_off;     // per variable offset into values
  const double* __restrict__ cal;      // Cal3_S2 table
};

template <int TYPE, bool WITH_J> struct Eval;

// BetweenFactor<Pose3>::evaluateError — gtsam/slam/BetweenFactor.h:111-124
// (default, GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR undefined):
//   hx = p1^-1 p2;  r = Logmap(measured^-1 hx);  H1 = -Ad(hx^-1);  H2 = I.
template <bool WITH_J> struct Eval<B200_FACTOR_BETWEEN_POSE3, WITH_J> {
  static __device__ __forceinline__ void run(const EvalCtx& c, int k0, int k1, const double* __restrict__ z, int,
                                             const double*, double* M) {
    enum { NC = 13 };
    const Pose x1 = load_pose(c.values + c.val_off[k0]);
    const Pose x2 = load_pose(c.values + c.val_off[k1]);
    const Pose zm = load_pose(z);
    const Pose hx = between(x1, x2);
    double r[6];
    pose_logmap(between(zm, hx), r);
#pragma unroll
    for (int i = 0; i < 6; i++) M[i * NC + 12] = -r[i];
    if (WITH_J) {
      // AdjointMap of hx^-1: [R 0; [t]x R, R]  (gtsam/geometry/Pose3.cpp:57-63)
      const Pose hi = inverse(hx);
      const Mat3 A = mul(hat(hi.t[0], hi.t[1], hi.t[2]), hi.R);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          M[i * NC + j] = -hi.R.m[3 * i + j];
          M[i * NC + 3 + j] = -0.0;
          M[(3 + i) * NC + j] = -A.m[3 * i + j];
          M[(3 + i) * NC + 3 + j] = -hi.R.m[3 * i + j];
        }
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) M[i * NC + 6 + j] = (i == j) ? 1.0 : 0.0;
    }
  }
};

// PriorFactor<Pose3>::evaluateError — gtsam/nonlinear/PriorFactor.h:98-102: r = -Local(x, prior), H = I
template <bool WITH_J> struct Eval<B200_FACTOR_PRIOR_POSE3, WITH_J> {
  static __device__ __forceinline__ void run(const EvalCtx& c, int k0, int, const double* __restrict__ z, int, const double*, double* M) {
    enum { NC = 7 };
    const Pose x = load_pose(c.values + c.val_off[k0]);
    const Pose pz = load_pose(z);
    double l[6];
    pose_local(x, pz, l);
#pragma unroll
    for (int i = 0; i < 6; i++) M[i * NC + 6] = l[i];  // b = -r = Local(x, prior)
    if (WITH_J) {
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) M[i * NC + j] = (i == j) ? 1.0 : 0.0;
    }
  }
};

// ---- Pose2 = (x, y, theta), tangent (x, y, theta): BASELINE configs[0]'s factor family ---------------------------
// between(a, b) = a^-1 b as Pose2::inverse (gtsam/geometry/Pose2.cpp:201-203) followed by operator* (Pose2.h:131-133;
// Rot2::operator* normalizes through fromCosSin, Rot2.cpp:27-30,56-64); theta() = atan2(s, c) (Rot2.h:186-188).
struct P2 { double x, y, c, s; };
__device__ __forceinline__ P2 load_pose2(const double* __restrict__ v) {
  P2 p;
  p.x = v[0]; p.y = v[1];
  sincos(v[2], &p.s, &p.c);
  return p;
}
__device__ __forceinline__ void rot2_normalize(double& c, double& s) {
  double scale = c * c + s * s;
  if (fabs(scale - 1.0) > 1e-10) { scale = 1.0 / sqrt(scale); c *= scale; s *= scale; }
}
__device__ __forceinline__ P2 between2(const P2& a, const P2& b) {
  P2 g;
  const double ix = a.c * (-a.x) + a.s * (-a.y), iy = -a.s * (-a.x) + a.c * (-a.y);   // unrotate(-t_a)
  g.c = a.c * b.c + a.s * b.s;
  g.s = -a.s * b.c + a.c * b.s;
  rot2_normalize(g.c, g.s);
  g.x = ix + (a.c * b.x + a.s * b.y);
  g.y = iy + (-a.s * b.x + a.c * b.y);
  return g;
}

// BetweenFactor<Pose2>::evaluateError — gtsam/slam/BetweenFactor.h:111-124 (fast variant):
//   hx = between(p1, p2), H1 = -AdjointMap(hx^-1), H2 = I (gtsam/base/Lie.h:63-69, Pose2::AdjointMap Pose2.cpp:127-135);
//   r = Local(measured, hx) = (x, y, theta) of measured^-1 hx (Pose2::ChartAtOrigin::Local, Pose2.cpp:111-122)
template <bool WITH_J> struct Eval<B200_FACTOR_BETWEEN_POSE2, WITH_J> {
  static __device__ __forceinline__ void run(const EvalCtx& c, int k0, int k1, const double* __restrict__ z, int,
                                             const double*, double* M) {
    enum { NC = 7 };
    const P2 hx = between2(load_pose2(c.values + c.val_off[k0]), load_pose2(c.values + c.val_off[k1]));
    const P2 d = between2(load_pose2(z), hx);
    M[0 * NC + 6] = -d.x; M[1 * NC + 6] = -d.y; M[2 * NC + 6] = -atan2(d.s, d.c);   // b = -r
    if (WITH_J) {
      // hx^-1 = (c, -s, unrotate(-t)); AdjointMap(p) = [[c, -s, y], [s, c, -x], [0, 0, 1]]
      const double xi = hx.c * (-hx.x) + hx.s * (-hx.y), yi = -hx.s * (-hx.x) + hx.c * (-hx.y);
      M[0 * NC + 0] = -hx.c; M[0 * NC + 1] = -hx.s; M[0 * NC + 2] = -yi;
      M[1 * NC + 0] = hx.s;  M[1 * NC + 1] = -hx.c; M[1 * NC + 2] = xi;
      M[2 * NC + 0] = -0.0;  M[2 * NC + 1] = -0.0;  M[2 * NC + 2] = -1.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) M[i * NC + 3 + j] = (i == j) ? 1.0 : 0.0;
    }
  }
};

// PriorFactor<Pose2>::evaluateError — gtsam/nonlinear/PriorFactor.h:98-102: r = -Local(x, prior), H = I
template <bool WITH_J> struct Eval<B200_FACTOR_PRIOR_POSE2, WITH_J> {
  static __device__ __forceinline__ void run(const EvalCtx& c, int k0, int, const double* __restrict__ z, int, const double*, double* M) {
    enum { NC = 4 };
    const P2 d = between2(load_pose2(c.values + c.val_off[k0]), load_pose2(z));
    M[0 * NC + 3] = d.x; M[1 * NC + 3] = d.y; M[2 * NC + 3] = atan2(d.s, d.c);   // b = -r = Local(x, prior)
    if (WITH_J) {
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) M[i * NC + j] = (i == j) ? 1.0 : 0.0;
    }
  }
};

template <bool WITH_J> struct Eval<B200_FACTOR_PRIOR_POINT3, WITH_J> {
  static __device__ __forceinline__ void run(const EvalCtx& c, int k0, int, const double* __restrict__ z, int, const double*, double* M) {
    enum { NC = 4 };
    const double* x = c.values + c.val_off[k0];
#pragma unroll
    for (int i = 0; i < 3; i++) M[i * NC + 3] = z[i] - x[i];
    if (WITH_J) {
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) M[i * NC + j] = (i == j) ? 1.0 : 0.0;
    }
  }
};

// PriorFactor<PinholeCamera<Cal3Bundler>>: Local = [pose local ; (f,k1,k2) difference]
// (gtsam/geometry/PinholeCamera.h:208-213, gtsam/geometry/Cal3Bundler.h:145-152)
template <bool WITH_J> struct Eval<B200_FACTOR_PRIOR_CAM_BUNDLER, WITH_J> {
  static __device__ __forceinline__ void run(const EvalCtx& c, int k0, int, const double* __restrict__ z, int, const double*, double* M) {
    enum { NC = 10 };
    const double* xv = c.values + c.val_off[k0];
    const Pose x = load_pose(xv);
    const Pose pz = load_pose(z);
    double l[6];
    pose_local(x, pz, l);
#pragma unroll
    for (int i = 0; i < 6; i++) M[i * NC + 9] = l[i];
#pragma unroll
    for (int i = 0; i < 3; i++) M[(6 + i) * NC + 9] = z[12 + i] - xv[12 + i];
    if (WITH_J) {
#pragma unroll
      for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) M[i * NC + j] = (i == j) ? 1.0 : 0.0;
    }
  }
};

// GenericProjectionFactor<Pose3,Point3,Cal3_S2>::evaluateError —
// gtsam/slam/ProjectionFactor.h:138-166; PinholePose::_project
// (gtsam/geometry/PinholePose.h:89-109); Cal3_S2::uncalibrate
// (gtsam/geometry/Cal3_S2.cpp:44-51).  Cheirality: H = 0, r = (2fx, 2fx).
template <bool WITH_J> struct Eval<B200_FACTOR_PROJECTION_CAL3S2, WITH_J> {
  static __device__ __forceinline__ void run(const EvalCtx& c, int k0, int k1, const double* __restrict__ z, int cal,
                                             const double* __restrict__ body, double* M) {
    enum { NC = 10 };
    Pose T = load_pose(c.values + c.val_off[k0]);
    Pose S;
    if (body) {   // camera = pose.compose(body_P_sensor, H0): gtsam/slam/ProjectionFactor.h:141-151
      S = load_pose(body);
      T = compose(T, S);
    }
    const double* pp = c.values + c.val_off[k1];
    const double p[3] = {pp[0], pp[1], pp[2]};
    const double* K = c.cal + 5 * cal;
    const double fx = K[0], fy = K[1], s = K[2], u0 = K[3], v0 = K[4];
    double u, v, Dp[12], Dq[6];
    if (project_normalized(T, p, u, v, Dp, Dq)) {
      const double px = fx * u + s * v + u0, py = fy * v + v0;
      M[9] = -(px - z[0]);
      M[NC + 9] = -(py - z[1]);
      if (WITH_J) {
#pragma unroll
        for (int j = 0; j < 6; j++) {
          M[j] = fx * Dp[j] + s * Dp[6 + j];
          M[NC + j] = 0.0 * Dp[j] + fy * Dp[6 + j];
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
          M[6 + j] = fx * Dq[j] + s * Dq[3 + j];
          M[NC + 6 + j] = 0.0 * Dq[j] + fy * Dq[3 + j];
        }
        if (body) {   // H1 <- H1 * H0 with H0 = Ad(body_P_sensor^-1) = [R 0; [t]x R, R] (gtsam/base/Lie.h compose)
          const Pose Si = inverse(S);
          const Mat3 A = mul(hat(Si.t[0], Si.t[1], Si.t[2]), Si.R);
#pragma unroll
          for (int r = 0; r < 2; r++) {
            double h[6];
#pragma unroll
            for (int j = 0; j < 6; j++) h[j] = M[r * NC + j];
#pragma unroll
            for (int j = 0; j < 3; j++) {
              M[r * NC + j] = h[0] * Si.R.m[j] + h[1] * Si.R.m[3 + j] + h[2] * Si.R.m[6 + j] +
                              h[3] * A.m[j] + h[4] * A.m[3 + j] + h[5] * A.m[6 + j];
              M[r * NC + 3 + j] = h[3] * Si.R.m[j] + h[4] * Si.R.m[3 + j] + h[5] * Si.R.m[6 + j];
            }
          }
        }
      }
    } else {
      M[9] = -2.0 * fx;
      M[NC + 9] = -2.0 * fx;
      if (WITH_J) {
#pragma unroll
        for (int j = 0; j < 9; j++) { M[j] = 0.0; M[NC + j] = 0.0; }
      }
    }
  }
};

// GeneralSFMFactor<PinholeCamera<Cal3Bundler>,Point3>::linearize —
// gtsam/slam/GeneralSFMFactor.h:141-177; PinholeCamera::project2
// (gtsam/geometry/PinholeCamera.h:230-247); Cal3Bundler::uncalibrate
// (gtsam/geometry/Cal3Bundler.cpp:66-92).  Cheirality: H = 0, b = 0.
template <bool WITH_J> struct Eval<B200_FACTOR_SFM_BUNDLER, WITH_J> {
  static __device__ __forceinline__ void run(const EvalCtx& c, int k0, int k1, const double* __restrict__ z, int, const double*, double* M) {
    enum { NC = 13 };
    const double* cv = c.values + c.val_off[k0];
    const Pose T = load_pose(cv);
    const double f = cv[12], k1c = cv[13], k2c = cv[14], u0 = cv[15], v0 = cv[16];
    const double* pp = c.values + c.val_off[k1];
    const double p[3] = {pp[0], pp[1], pp[2]};
    double x, y, Dp[12], Dq[6];
    if (project_normalized(T, p, x, y, Dp, Dq)) {
      const double r = x * x + y * y;
      const double g = 1. + (k1c + k2c * r) * r;
      const double u = g * x, v = g * y;
      M[12] = -((u0 + f * u) - z[0]);
      M[NC + 12] = -((v0 + f * v) - z[1]);
      if (WITH_J) {
        const double a = 2. * (k1c + 2. * k2c * r);
        const double axx = a * x * x, axy = a * x * y, ayy = a * y * y;
        const double D00 = (g + axx) * f, D01 = axy * f, D10 = axy * f, D11 = (g + ayy) * f;
#pragma unroll
        for (int j = 0; j < 6; j++) {
          M[j] = D00 * Dp[j] + D01 * Dp[6 + j];
          M[NC + j] = D10 * Dp[j] + D11 * Dp[6 + j];
        }
        const double rx = r * x, ry = r * y;
        M[6] = u; M[7] = f * rx; M[8] = f * r * rx;
        M[NC + 6] = v; M[NC + 7] = f * ry; M[NC + 8] = f * r * ry;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          M[9 + j] = D00 * Dq[j] + D01 * Dq[3 + j];
          M[NC + 9 + j] = D10 * Dq[j] + D11 * Dq[3 + j];
        }
      }
    } else {
      M[12] = 0.0;
      M[NC + 12] = 0.0;
      if (WITH_J) {
#pragma unroll
        for (int j = 0; j < 12; j++) { M[j] = 0.0; M[NC + j] = 0.0; }
      }
    }
  }
};

// Noise whitening of columns [C0, NC) of a row-major D x NC block:
// Unit / Isotropic (gtsam/linear/NoiseModel.cpp:646-675) / Diagonal (:322-340)
// / Gaussian with sqrt information R (:163-238).
template <int D, int NC, int C0>
__device__ __forceinline__ void whiten(double* M, int kind, const double* __restrict__ nz) {
  if (kind == B200_NOISE_UNIT) return;
  if (kind == B200_NOISE_ISOTROPIC) {
    const double inv = 1.0 / nz[0];
#pragma unroll
    for (int r = 0; r < D; r++)
#pragma unroll
      for (int c = C0; c < NC; c++) M[r * NC + c] *= inv;
  } else if (kind == B200_NOISE_DIAGONAL) {
#pragma unroll
    for (int r = 0; r < D; r++) {
      const double inv = 1.0 / nz[r];
#pragma unroll
      for (int c = C0; c < NC; c++) M[r * NC + c] *= inv;
    }
  } else {
#pragma unroll
    for (int c = C0; c < NC; c++) {
      double col[D];
#pragma unroll
      for (int r = 0; r < D; r++) col[r] = M[r * NC + c];
#pragma unroll
      for (int r = 0; r < D; r++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < D; k++) s += nz[r * D + k] * col[k];
        M[r * NC + c] = s;
      }
    }
  }
}

// m-estimators of noiseModel::Robust (gtsam/linear/LossFunctions.cpp:146-267)
__device__ __forceinline__ double robust_weight(int kind, double k, double distance) {
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
__device__ __forceinline__ double robust_loss(int kind, double k, double distance) {
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

}  // namespace b200
