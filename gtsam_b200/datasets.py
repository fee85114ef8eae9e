"""Synthetic workloads of BASELINE.json / SURVEY.md §8(d), generated with numpy.

The generators build the flat :class:`Problem` directly (no GTSAM objects) so the
same bytes feed the CUDA path, the C oracle and — through ``Problem.save`` and
oracle/ref_harness.cpp — the unmodified reference.

* ``sphere``  — config 2: Pose3 "sphere2500-style" graph, BetweenFactor<Pose3>
  with the sphere2500 information (examples/Data/sphere2500.txt:1), prior on
  pose 0 (examples/Pose3SLAMExample_g2o.cpp:42-49).
* ``bal``     — configs 3-5: cameras on a ring looking at a point cloud,
  GenericProjectionFactor<Pose3,Point3,Cal3_S2> (or the BAL-native
  GeneralSFMFactor<PinholeCamera<Cal3Bundler>,Point3> variant), Schur ordering
  (timing/timeSFMBAL.h:74-83: all points, then all cameras).
"""
from __future__ import annotations

import numpy as np

from . import problem as P


# ---- small SE(3) helpers for data generation (batch, numpy) ---------------------
def _hat(w):
    z = np.zeros(w.shape[:-1])
    return np.stack([np.stack([z, -w[..., 2], w[..., 1]], -1),
                     np.stack([w[..., 2], z, -w[..., 0]], -1),
                     np.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def so3_exp(w):
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    W = _hat(w)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    A = np.where(small, 1.0, np.sin(ths) / ths)
    B = np.where(small, 0.5, (1 - np.cos(ths)) / ths ** 2)
    return np.eye(3) + A * W + B * (W @ W)


def se3_exp(xi):
    """Pose3::Expmap (gtsam/geometry/Pose3.cpp:169-185); xi = (omega, v)."""
    xi = np.asarray(xi, dtype=np.float64)
    w, v = xi[..., :3], xi[..., 3:]
    R = so3_exp(w)
    th2 = np.sum(w * w, -1)[..., None]
    wxv = np.cross(w, v)
    tpar = w * np.sum(w * v, -1)[..., None]
    small = th2 < 1e-16
    t = np.where(small, v, (wxv - np.einsum("...ij,...j->...i", R, wxv) + tpar) / np.where(small, 1.0, th2))
    return R, t


def pose_compose(Ra, ta, Rb, tb):
    return Ra @ Rb, ta + np.einsum("...ij,...j->...i", Ra, tb)


def pose_between(Ra, ta, Rb, tb):
    Rat = np.swapaxes(Ra, -1, -2)
    return Rat @ Rb, np.einsum("...ij,...j->...i", Rat, tb - ta)


def pack_pose(R, t):
    return np.concatenate([R.reshape(R.shape[:-2] + (9,)), t], -1)


def rot_ypr(y, p, r):
    """Rot3::Ypr = Rz(y) Ry(p) Rx(r) (gtsam/geometry/Rot3.h)."""
    cy, sy, cp, sp, cr, sr = np.cos(y), np.sin(y), np.cos(p), np.sin(p), np.cos(r), np.sin(r)
    o, z = np.ones_like(y), np.zeros_like(y)
    Rz = np.stack([np.stack([cy, -sy, z], -1), np.stack([sy, cy, z], -1), np.stack([z, z, o], -1)], -2)
    Ry = np.stack([np.stack([cp, z, sp], -1), np.stack([z, o, z], -1), np.stack([-sp, z, cp], -1)], -2)
    Rx = np.stack([np.stack([o, z, z], -1), np.stack([z, cr, -sr], -1), np.stack([z, sr, cr], -1)], -2)
    return Rz @ Ry @ Rx


# ---- config 2: sphere ------------------------------------------------------------
def sphere(layers: int = 50, per_ring: int = 50, radius: float = 50.0, seed: int = 7,
           ordering: str = "natural", noise: str = "diagonal", robust=None) -> P.Problem:
    rng = np.random.default_rng(seed)
    n = layers * per_ring
    idx = np.arange(n)
    ring, k = idx // per_ring, idx % per_ring
    phi = np.pi * (ring + 1) / (layers + 1)        # polar angle, poles excluded
    theta = 2 * np.pi * k / per_ring
    t = radius * np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)], -1)
    R = rot_ypr(theta + np.pi / 2, np.zeros(n), phi - np.pi / 2)
    # edges i->i+1, i->i+per_ring, i->i+per_ring+1, i->i+2*per_ring
    e = []
    for off in (1, per_ring, per_ring + 1, 2 * per_ring):
        i = idx[idx + off < n]
        e.append(np.stack([i, i + off], -1))
    edges = np.concatenate(e, 0)
    edges = edges[np.lexsort((edges[:, 1], edges[:, 0]))]
    prec = np.array([10.0, 10, 10, 100, 100, 25])   # sphere2500 information diagonal (t first in
    # g2o; here already permuted to GTSAM's (omega, v) order: rot 10,10,10... ) kept as stated in SURVEY
    sig = 1.0 / np.sqrt(prec)
    Ri, ti = R[edges[:, 0]], t[edges[:, 0]]
    Rj, tj = R[edges[:, 1]], t[edges[:, 1]]
    Rz, tz = pose_between(Ri, ti, Rj, tj)
    Rn, tn = se3_exp(rng.normal(size=(edges.shape[0], 6)) * sig)
    Rz, tz = pose_compose(Rz, tz, Rn, tn)
    # initial estimate: gt perturbed by N(0, 0.02 rad / 0.2 m)
    Rp, tp = se3_exp(rng.normal(size=(n, 6)) * np.array([0.02] * 3 + [0.2] * 3))
    R0, t0 = pose_compose(R, t, Rp, tp)
    values = pack_pose(R0, t0).ravel()
    prior = P.FactorGroup(P.FACTOR_PRIOR_POSE3, np.array([[0]]), pack_pose(R[0], t[0])[None],
                          P.NOISE_DIAGONAL, np.sqrt(np.array([1e-6] * 3 + [1e-4] * 3)))
    if noise == "diagonal":
        between = P.FactorGroup(P.FACTOR_BETWEEN_POSE3, edges, pack_pose(Rz, tz), P.NOISE_DIAGONAL, sig)
    elif noise == "gaussian":
        # per-factor full information matrices (what g2o EDGE_SE3:QUAT lines carry,
        # gtsam/slam/dataset.cpp:838-859): R = upper Cholesky factor of a random SPD information
        ne = edges.shape[0]
        A = rng.normal(size=(ne, 6, 6)) * 0.3
        info = np.einsum("nij,nkj->nik", A, A) + np.diag(prec)[None]
        Rup = np.transpose(np.linalg.cholesky(info), (0, 2, 1))          # info = R^T R, R upper
        between = P.FactorGroup(P.FACTOR_BETWEEN_POSE3, edges, pack_pose(Rz, tz), P.NOISE_GAUSSIAN, Rup.reshape(ne, 36))
    else:
        raise ValueError(noise)
    # graph order as Pose3SLAMExample_g2o builds it: between factors, then the prior
    if ordering == "natural":
        order = np.arange(n)
    elif ordering == "reverse":
        order = np.arange(n)[::-1].copy()
    elif ordering in ("colamd", "metis"):
        # orderings are INPUTS at the boundary: these were produced by the reference's own
        # Ordering::Colamd / Ordering::Metis (oracle/ref_harness order) for exactly this graph
        # (layers=50, per_ring=50; the structure does not depend on the seed) and are shipped as data
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"data_sphere{n}_{ordering}.npy")
        if not os.path.exists(path):
            raise ValueError(f"no stored {ordering} ordering for a {n}-pose sphere ({path})")
        order = np.load(path).astype(np.int64)
    else:
        raise ValueError(ordering)
    if robust:
        # a few gross outliers + a Huber/Cauchy/... kernel, as Pose2SLAMExample_g2o exposes for 2D
        # (examples/Pose2SLAMExample_g2o.cpp:53-61)
        bad = rng.choice(edges.shape[0], size=max(1, edges.shape[0] // 15), replace=False)
        Rb, tb = se3_exp(rng.normal(size=(bad.size, 6)) * np.array([0.5] * 3 + [5.0] * 3))
        zb = between.meas[bad]
        Rzb, tzb = pose_compose(zb[:, :9].reshape(-1, 3, 3), zb[:, 9:], Rb, tb)
        between.meas[bad] = pack_pose(Rzb, tzb)
        between.robust_kind, between.robust_param = robust
    pr = P.Problem(np.full(n, P.VAR_POSE3), values, order, [between, prior], name=f"sphere{n}")
    pr.meta = dict(kind="sphere", layers=layers, per_ring=per_ring, seed=seed, gt=pack_pose(R, t), ordering=ordering)
    return pr


# ---- configs 3-5: BAL --------------------------------------------------------------
def lookat_pose(eye, target, up):
    """PinholeBase::LookatPose, gtsam/geometry/CalibratedCamera.cpp:58-66."""
    zc = target - eye
    zc = zc / np.linalg.norm(zc, axis=-1, keepdims=True)
    xc = np.cross(-up, zc)
    xc = xc / np.linalg.norm(xc, axis=-1, keepdims=True)
    yc = np.cross(zc, xc)
    return np.stack([xc, yc, zc], -1), eye   # columns = camera axes


def ordering_file(ncams, npoints, obs_per_point, visibility, seed, kind):
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        f"data_bal_{ncams}c_{npoints}p_{obs_per_point}o_{visibility}_s{seed}_{kind}.npz")


def bal(ncams: int = 100, npoints: int = 50000, obs_per_point: int = 6, visibility: str = "scattered",
        camera_model: str = "cal3_s2", seed: int = 42, pixel_sigma: float = 1.0, body_sensor: bool = False,
        ordering: str = "schur") -> P.Problem:
    rng = np.random.default_rng(seed)
    th = 2 * np.pi * np.arange(ncams) / ncams
    eye = np.stack([20 * np.cos(th), 20 * np.sin(th), 2 * np.sin(3 * th)], -1)
    Rc, tc = lookat_pose(eye, np.zeros(3), np.array([0.0, 0, 1.0]))
    pts = rng.uniform(-3, 3, size=(npoints, 3))
    start = rng.integers(0, ncams, size=npoints)
    k = np.arange(obs_per_point)
    if visibility == "banded":
        cam = (start[:, None] + 3 * k[None]) % ncams
    elif visibility == "scattered":
        cam = (start[:, None] + 13 * k[None] * (1 + 7 * k[None])) % ncams
    else:
        raise ValueError(visibility)
    pid = np.repeat(np.arange(npoints), obs_per_point)
    cid = cam.ravel()
    # project ground truth
    q = np.einsum("nji,nj->ni", Rc[cid], pts[pid] - tc[cid])     # R^T (p - t)
    pn = q[:, :2] / q[:, 2:3]
    assert np.all(q[:, 2] > 0)
    noise = rng.normal(size=pn.shape) * pixel_sigma
    Rp, tp = se3_exp(rng.normal(size=(ncams, 6)) * 0.01)
    R0, t0 = pose_compose(Rc, tc, Rp, tp)
    pts0 = pts + rng.normal(size=pts.shape) * 0.05
    # variable ids follow Key order: C(i) = 'c'<<56|i  <  P(j) = 'p'<<56|j
    keys = np.stack([cid, ncams + pid], -1)
    if ordering == "schur":
        order = np.concatenate([ncams + np.arange(npoints), np.arange(ncams)])   # points, then cameras
    elif ordering in ("metis", "colamd"):
        # orderings are INPUTS at the boundary: produced by the reference's own Ordering::Metis / Ordering::Colamd
        # for exactly this graph (tests/golden/make_orderings.py through oracle/ref_harness order) and shipped as data;
        # the structure depends on every generator argument, hence the long file name
        import os
        path = ordering_file(ncams, npoints, obs_per_point, visibility, seed, ordering)
        if not os.path.exists(path):
            raise ValueError(f"no stored {ordering} ordering for this BAL graph ({path}); tests/golden/make_orderings.py writes it "
                             "in the build container")
        order = np.load(path)["ordering"].astype(np.int64)
        assert order.size == ncams + npoints
    else:
        raise ValueError(ordering)
    if camera_model == "cal3_s2":
        K = np.array([[500.0, 500.0, 0.0, 320.0, 240.0]])
        z = np.stack([K[0, 0] * pn[:, 0] + K[0, 2] * pn[:, 1] + K[0, 3], K[0, 1] * pn[:, 1] + K[0, 4]], -1) + noise
        body = None
        Rpri, tpri = Rc[:2], tc[:2]
        if body_sensor:
            # the variables are BODY poses; the camera sits at body * body_P_sensor
            # (GenericProjectionFactor's optional argument, gtsam/slam/ProjectionFactor.h:141-151)
            Rs, ts = se3_exp(np.array([0.05, -0.1, 0.2, 0.3, -0.2, 0.1]))
            Rsi, tsi = Rs.T, -Rs.T @ ts
            R0, t0 = pose_compose(R0, t0, Rsi, tsi)
            Rpri, tpri = pose_compose(Rpri, tpri, Rsi, tsi)
            body = pack_pose(Rs, ts)
        cams = pack_pose(R0, t0)
        var_type = np.concatenate([np.full(ncams, P.VAR_POSE3), np.full(npoints, P.VAR_POINT3)])
        proj = P.FactorGroup(P.FACTOR_PROJECTION_CAL3S2, keys, z, P.NOISE_ISOTROPIC, np.array([pixel_sigma]),
                             body_P_sensor=body)
        prior = P.FactorGroup(P.FACTOR_PRIOR_POSE3, np.array([[0], [1]]), pack_pose(Rpri, tpri),
                              P.NOISE_ISOTROPIC, np.array([0.1]))
        groups, cal = [proj, prior], K
    elif camera_model == "bundler":
        f, k1, k2 = 500.0, -0.02, 0.002
        r2 = np.sum(pn * pn, -1, keepdims=True)
        g = 1 + (k1 + k2 * r2) * r2
        z = f * g * pn + noise
        intr = np.tile(np.array([f, k1, k2, 0.0, 0.0]), (ncams, 1))
        cams = np.concatenate([pack_pose(R0, t0), intr], -1)
        gtc = np.concatenate([pack_pose(Rc, tc), intr], -1)
        var_type = np.concatenate([np.full(ncams, P.VAR_CAM_BUNDLER), np.full(npoints, P.VAR_POINT3)])
        proj = P.FactorGroup(P.FACTOR_SFM_BUNDLER, keys, z, P.NOISE_ISOTROPIC, np.array([pixel_sigma]))
        prior = P.FactorGroup(P.FACTOR_PRIOR_CAM_BUNDLER, np.array([[0], [1]]), gtc[:2],
                              P.NOISE_ISOTROPIC, np.array([0.1]))
        groups, cal = [proj, prior], np.zeros((0, 5))
    else:
        raise ValueError(camera_model)
    values = np.concatenate([cams.ravel(), pts0.ravel()])
    pr = P.Problem(var_type, values, order, groups, cal,
                   name=f"bal_{ncams}c_{npoints}p_{visibility}_{camera_model}")
    pr.meta = dict(kind="bal", ncams=ncams, npoints=npoints, visibility=visibility,
                   camera_model=camera_model, seed=seed, ordering=ordering)
    return pr


# ---- config 1 family: planar pose graphs ----------------------------------------------------------
def pose2_ring(n: int = 40, seed: int = 21, ordering: str = "natural") -> P.Problem:
    """A Pose2 pose graph of the kind Pose2SLAMExample_g2o reads (BASELINE configs[0]): n poses on a ring expressed in
    the frame of pose 0, BetweenFactor<Pose2> odometry + two families of loop closures with the information
    diag(400, 400, 10000), and the example's prior on pose 0 (examples/Pose2SLAMExample_g2o.cpp:62-64)."""
    rng = np.random.default_rng(seed)
    th = 2 * np.pi * np.arange(n) / n
    ring = np.stack([10 * np.cos(th), 10 * np.sin(th), th + np.pi / 2], -1)

    def between(a, b):
        c, s = np.cos(a[2]), np.sin(a[2])
        d = b[:2] - a[:2]
        return np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], (b[2] - a[2] + np.pi) % (2 * np.pi) - np.pi])

    gt = np.array([between(ring[0], q) for q in ring])
    init = gt + rng.normal(size=gt.shape) * np.array([0.3, 0.3, 0.05])
    init[0] = 0
    edges = [(i, (i + 1) % n) for i in range(n)] + [(i, (i + 7) % n) for i in range(0, n, 3)] + [(i, (i + 19) % n) for i in range(1, n, 5)]
    z = np.array([between(gt[i], gt[j]) + rng.normal(size=3) * np.array([0.05, 0.05, 0.01]) for i, j in edges])
    btw = P.FactorGroup(P.FACTOR_BETWEEN_POSE2, np.array(edges), z, P.NOISE_DIAGONAL, 1.0 / np.sqrt(np.array([400.0, 400.0, 10000.0])))
    pri = P.FactorGroup(P.FACTOR_PRIOR_POSE2, np.array([[0]]), np.zeros((1, 3)), P.NOISE_DIAGONAL, np.sqrt(np.array([1e-6, 1e-6, 1e-8])))
    order = np.arange(n) if ordering == "natural" else np.arange(n)[::-1].copy()
    pr = P.Problem(np.full(n, P.VAR_POSE2), init.ravel(), order, [btw, pri], name=f"pose2_ring{n}")
    pr.meta = dict(kind="pose2", n=n, seed=seed, ordering=ordering)
    return pr


WORKLOADS = {
    # name: (builder, kwargs) — BASELINE.json configs
    "sphere2500": (sphere, dict(layers=50, per_ring=50, ordering="colamd")),   # configs[1]: 2.5k Pose3 / 9.8k Between, COLAMD
    "sphere2500_metis": (sphere, dict(layers=50, per_ring=50, ordering="metis")),
    "sphere2500_natural": (sphere, dict(layers=50, per_ring=50)),
    "bal_c3": (bal, dict(ncams=100, npoints=50000)),                 # configs[2]: 300k factors
    "bal_1m": (bal, dict(ncams=300, npoints=166667)),                # north_star 1M-factor target
    "bal_c4": (bal, dict(ncams=1000, npoints=500000)),               # configs[3]: 3M factors
    "bal_c5": (bal, dict(ncams=5000, npoints=2000000, obs_per_point=5)),            # configs[4]: 10M factors (FP64 here)
    # the same graphs eliminated in the reference's METIS nested-dissection order (what configs[3] names): all points stay
    # leaves, fewer flops (bal_c4 6.6e9 vs 1.45e10, bal_c5 1.9e10 vs 7.5e10), 24-25 levels instead of 19 / 134, and
    # balanced subtrees for the multi-GPU plan
    "bal_1m_metis": (bal, dict(ncams=300, npoints=166667, ordering="metis")),
    "bal_c4_metis": (bal, dict(ncams=1000, npoints=500000, ordering="metis")),
    "bal_c5_metis": (bal, dict(ncams=5000, npoints=2000000, obs_per_point=5, ordering="metis")),
    "bal_tiny": (bal, dict(ncams=10, npoints=60, visibility="banded")),
    "sphere_tiny": (sphere, dict(layers=5, per_ring=8)),
    "pose2_ring": (pose2_ring, dict(n=40)),                          # configs[0]'s factor family (planar pose graph)
}


def make(name: str, **over) -> P.Problem:
    fn, kw = WORKLOADS[name]
    kw = dict(kw)
    kw.update(over)
    return fn(**kw)
