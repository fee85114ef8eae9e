"""Dataset readers that feed the flat :class:`Problem` tables directly.

The reference goes text -> ``SfmData`` / ``NonlinearFactorGraph`` + ``Values`` (heap objects)
-> and only then could be packed; here the text is parsed straight into the SoA the device
consumes.  Conventions follow the reference's own loaders line by line:

* BAL ("Bundle Adjustment in the Large") files — ``SfmData::FromBalFile``
  (gtsam/sfm/SfmData.cpp:189-246): every number after the ids is parsed **through float**
  (``float u, v; is >> u >> v``), measurements are ``(u, -v)``, the Rodrigues vector gives the
  OpenGL rotation and ``openGL2gtsam`` (gtsam/sfm/SfmData.cpp:79-86) turns it into a GTSAM
  camera pose ``wRc = R^-1 * diag(1,-1,-1)``, ``wtc = R^T (-t)``; calibration
  ``Cal3Bundler(f, k1, k2)``.
* g2o 3D pose graphs — ``load3D`` (gtsam/slam/dataset.cpp:922-944) with
  ``VERTEX_SE3:QUAT`` (:756-772, quaternion normalised as in :738-744) and ``EDGE_SE3:QUAT``
  (:811-866: the 6x6 information is stored in (t, R) order and permuted to GTSAM's (R, t)).

Variable ids are assigned in ascending Key order (cameras before points: Symbol 'c' < 'p').
Orderings stay an input: the readers return the Schur ordering for BAL and the natural
ordering for pose graphs unless one is supplied.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import problem as P


# ---- small helpers that mirror the reference's arithmetic -----------------------------
def _rodrigues(w):
    """Rot3::Rodrigues == SO3::Expmap (gtsam/geometry/SO3.cpp:49-87), row-major 3x3."""
    wx, wy, wz = (float(x) for x in w)
    theta2 = wx * wx + wy * wy + wz * wz
    W = np.array([[0.0, -wz, wy], [wz, 0.0, -wx], [-wy, wx, 0.0]])
    if theta2 <= np.finfo(np.float64).eps:
        return np.eye(3) + W
    theta = math.sqrt(theta2)
    s2 = math.sin(theta / 2.0)
    K = W / theta
    return np.eye(3) + math.sin(theta) * K + (2.0 * s2 * s2) * (K @ K)


def _quat_to_rot(x, y, z, w):
    """istream >> Quaternion (dataset.cpp:738-744) then Eigen's toRotationMatrix."""
    f = 1.0 / math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = f * w, f * x, f * y, f * z
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1.0 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1.0 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1.0 - (txx + tyy)]])


def _rot_log(R):
    """Rotation vector of R (writer side only)."""
    c = max(-1.0, min(1.0, (np.trace(R) - 1.0) / 2.0))
    th = math.acos(c)
    if th < 1e-12:
        return np.zeros(3)
    return th / (2.0 * math.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])


def _rot_to_quat(R):
    w = math.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
    if w > 1e-6:
        return (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w
    x = math.sqrt(max(0.0, 1.0 + R[0, 0] - R[1, 1] - R[2, 2])) / 2.0
    return x, (R[0, 1] + R[1, 0]) / (4 * x), (R[0, 2] + R[2, 0]) / (4 * x), (R[2, 1] - R[1, 2]) / (4 * x)


_R90 = np.diag([1.0, -1.0, -1.0])


# ---- BAL ---------------------------------------------------------------------------------
def read_bal(path: str, noise_sigma: Optional[float] = None, priors: bool = False) -> P.Problem:
    """BAL file -> GeneralSFMFactor<PinholeCamera<Cal3Bundler>, Point3> problem.

    ``noise_sigma=None``: unit noise (tests/testGeneralSFMFactorB.cpp:44-63);
    ``priors=True`` adds the priors of examples/SFMExample_bal.cpp:66-67 (camera 0 and point 0,
    isotropic 0.1) after the projection factors."""
    tok = open(path).read().split()
    it = iter(tok)
    ncam, npts, nobs = int(next(it)), int(next(it)), int(next(it))
    f32 = lambda: float(np.float32(next(it)))   # noqa: E731  the reference parses through float
    cam_idx = np.zeros(nobs, dtype=np.int64)
    pt_idx = np.zeros(nobs, dtype=np.int64)
    uv = np.zeros((nobs, 2))
    for k in range(nobs):
        cam_idx[k], pt_idx[k] = int(next(it)), int(next(it))
        u, v = f32(), f32()
        uv[k] = (u, -v)
    cams = np.zeros((ncam, 17))
    for i in range(ncam):
        w = (f32(), f32(), f32())
        t = np.array([f32(), f32(), f32()])
        R = _rodrigues(w)
        wRc = R.T @ _R90
        wtc = R.T @ (-t)
        cams[i, :9] = wRc.ravel()
        cams[i, 9:12] = wtc
        cams[i, 12:15] = (f32(), f32(), f32())
    pts = np.zeros((npts, 3))
    for j in range(npts):
        pts[j] = (f32(), f32(), f32())
    # tracks hold their measurements in file order; the graph is built track by track
    order_k = np.lexsort((np.arange(nobs), pt_idx))
    keys = np.stack([cam_idx[order_k], ncam + pt_idx[order_k]], -1)
    kind = P.NOISE_UNIT if noise_sigma is None else P.NOISE_ISOTROPIC
    groups = [P.FactorGroup(P.FACTOR_SFM_BUNDLER, keys, uv[order_k], kind, None if noise_sigma is None else np.array([noise_sigma]))]
    if priors:
        groups.append(P.FactorGroup(P.FACTOR_PRIOR_CAM_BUNDLER, np.array([[0]]), cams[:1], P.NOISE_ISOTROPIC, np.array([0.1])))
        groups.append(P.FactorGroup(P.FACTOR_PRIOR_POINT3, np.array([[ncam]]), pts[:1], P.NOISE_ISOTROPIC, np.array([0.1])))
    var_type = np.concatenate([np.full(ncam, P.VAR_CAM_BUNDLER), np.full(npts, P.VAR_POINT3)])
    ordering = np.concatenate([ncam + np.arange(npts), np.arange(ncam)])      # Schur: points, then cameras
    pr = P.Problem(var_type, np.concatenate([cams.ravel(), pts.ravel()]), ordering, groups, name=f"bal:{path}")
    pr.meta = dict(kind="bal", ncams=ncam, npoints=npts, ordering="schur")
    return pr


def write_bal(path: str, cams: np.ndarray, pts: np.ndarray, cam_idx, pt_idx, uv) -> None:
    """Inverse of read_bal for synthetic data: cams (n,17) GTSAM convention, uv = GTSAM (u, v)."""
    with open(path, "w") as f:
        f.write(f"{len(cams)} {len(pts)} {len(cam_idx)}\n")
        for c, p, m in zip(cam_idx, pt_idx, uv):
            f.write(f"{int(c)} {int(p)} {m[0]:.9g} {-m[1]:.9g}\n")
        for c in cams:
            wRc, wtc = c[:9].reshape(3, 3), c[9:12]
            R = (wRc @ _R90).T                       # openGL rotation
            t = -(R @ wtc)
            for x in list(_rot_log(R)) + list(t) + list(c[12:15]):
                f.write(f"{x:.9g}\n")
        for p in pts:
            for x in p:
                f.write(f"{x:.9g}\n")


# ---- g2o 3D ------------------------------------------------------------------------------
def read_g2o_3d(path: str, add_prior: bool = True, ordering: Optional[np.ndarray] = None) -> P.Problem:
    """VERTEX_SE3:QUAT / EDGE_SE3:QUAT file -> BetweenFactor<Pose3> problem, plus (like
    examples/Pose3SLAMExample_g2o.cpp:42-49) a PriorFactor<Pose3>(firstKey, identity,
    Diagonal::Variances(1e-6 x3, 1e-4 x3)) when ``add_prior``."""
    vid, vpose, edges, meas, infos = [], [], [], [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "VERTEX_SE3:QUAT":
            x = [float(v) for v in t[2:9]]
            R = _quat_to_rot(x[3], x[4], x[5], x[6])
            vid.append(int(t[1]))
            vpose.append(np.concatenate([R.ravel(), x[:3]]))
        elif t[0] == "EDGE_SE3:QUAT":
            x = [float(v) for v in t[3:10]]
            R = _quat_to_rot(x[3], x[4], x[5], x[6])
            edges.append((int(t[1]), int(t[2])))
            meas.append(np.concatenate([R.ravel(), x[:3]]))
            m = np.zeros((6, 6))
            vals = iter(float(v) for v in t[10:31])
            for i in range(6):
                for j in range(i, 6):
                    m[i, j] = m[j, i] = next(vals)
            g = np.zeros((6, 6))                       # g2o (t, R) order -> GTSAM (R, t)
            g[:3, :3], g[3:, 3:], g[3:, :3], g[:3, 3:] = m[3:, 3:], m[:3, :3], m[:3, 3:], m[3:, :3]
            infos.append(g)
    vid = np.asarray(vid, dtype=np.int64)
    srt = np.argsort(vid)                               # ids in ascending Key order
    rank = {int(k): i for i, k in enumerate(vid[srt])}
    values = np.asarray(vpose)[srt]
    keys = np.array([[rank[a], rank[b]] for a, b in edges], dtype=np.int64)
    # noiseModel::Gaussian::Information(M): R = upper Cholesky factor, info = R^T R
    Rs = np.stack([np.linalg.cholesky(g).T for g in infos]).reshape(len(infos), 36)
    groups = [P.FactorGroup(P.FACTOR_BETWEEN_POSE3, keys, np.asarray(meas), P.NOISE_GAUSSIAN, Rs)]
    if add_prior:
        ident = np.array([[1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]])
        groups.append(P.FactorGroup(P.FACTOR_PRIOR_POSE3, np.array([[0]]), ident, P.NOISE_DIAGONAL,
                                    np.sqrt(np.array([1e-6] * 3 + [1e-4] * 3))))
    n = len(vid)
    pr = P.Problem(np.full(n, P.VAR_POSE3), values.ravel(), np.arange(n) if ordering is None else ordering, groups,
                   name=f"g2o:{path}")
    pr.meta = dict(kind="g2o", keys=vid[srt], ordering="natural" if ordering is None else "given")
    return pr


def write_g2o_3d(path: str, poses: np.ndarray, edges: np.ndarray, meas: np.ndarray, infos_gtsam: np.ndarray) -> None:
    """poses (n,12), meas (m,12) GTSAM convention; infos_gtsam (m,6,6) in (R, t) order."""
    with open(path, "w") as f:
        for i, p in enumerate(poses):
            q = _rot_to_quat(p[:9].reshape(3, 3))
            f.write("VERTEX_SE3:QUAT %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n" % (i, p[9], p[10], p[11], *q))
        for (a, b), z, g in zip(edges, meas, infos_gtsam):
            q = _rot_to_quat(z[:9].reshape(3, 3))
            m = np.zeros((6, 6))                       # back to g2o (t, R) order
            m[3:, 3:], m[:3, :3], m[:3, 3:], m[3:, :3] = g[:3, :3], g[3:, 3:], g[3:, :3], g[:3, 3:]
            up = " ".join("%.17g" % m[i, j] for i in range(6) for j in range(i, 6))
            f.write("EDGE_SE3:QUAT %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %s\n" % (a, b, z[9], z[10], z[11], *q, up))
