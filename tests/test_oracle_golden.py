"""CPU: the C oracle (oracle/oracle.c) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/, see make_golden.py) and against the
known-answer tests of the reference's own test-suite.  This is what makes the
oracle a trustworthy checker for the CUDA path."""
import numpy as np
import pytest

import util
from gtsam_b200 import problem as P
from oracle import oracle_py as O


@pytest.mark.parametrize("case", util.CASES + util.EXTRA_CASES)
@pytest.mark.parametrize("kind,lam,diag", [("dump0", 0.0, 0), ("dump1", 1e-2, 1)])
def test_oracle_matches_reference_dump(case, kind, lam, diag):
    prob = util.load_case(case)
    util.check_against_dump(O.OracleProblem(prob), prob, util.golden(case, kind), lam, diag)


@pytest.mark.parametrize("case", util.CASES + util.EXTRA_CASES)
def test_oracle_lm_trace(case):
    """Same accept/reject sequence, lambdas and errors as LevenbergMarquardtOptimizer."""
    prob = util.load_case(case)
    ref = util.golden(case, "lm")
    op = O.OracleProblem(prob)
    lm = op.lm(util.lm_params(case, max_iterations=100 if case.startswith("dub") else 30))
    errs, lams, inner = [lm.state.error], [lm.state.lambda_], [0]
    cur = lm.state.error
    from gtsam_b200.optimizer import checkConvergence
    while True:
        op.lm_iterate(lm)
        new = lm.state.error
        errs.append(new); lams.append(lm.state.lambda_); inner.append(lm.state.total_inner_iterations)
        if not (lm.state.iterations < lm.params.max_iterations and not checkConvergence(
                lm.params.relative_error_tol, lm.params.absolute_error_tol, lm.params.error_tol, cur, new)):
            break
        cur = new
    assert len(errs) == len(ref["lm_errors"])
    # dubrovnik-3-7: cond(H) ~ 1e15 at the first lambdas -> trajectories agree to ~1e-6 only
    assert np.allclose(errs, ref["lm_errors"], rtol=1e-5 if case.startswith("dub") else 1e-7, atol=1e-10)
    assert np.allclose(lams, ref["lm_lambdas"], rtol=1e-12)
    assert inner == list(ref["lm_inner"])
    assert util.relmax(op.get_values(), ref["final_values"]) <= (1e-3 if case.startswith("dub") else 1e-6)


def test_oracle_reference_end_to_end_golden():
    """tests/testGeneralSFMFactorB.cpp:44-63: dubrovnik-3-7-pre, LM -> 0.0199833 +- 1e-5."""
    prob = util.load_case("dubrovnik_3_7_unit")
    op = O.OracleProblem(prob)
    lm = op.lm(util.lm_params())
    op.lm_optimize(lm)
    assert abs(lm.state.error - 0.0199833) < 1e-5


@pytest.mark.parametrize("case", ["sphere_tiny", "sphere_small_colamd", "pose2_ring", "pose2_ring_colamd"])
def test_oracle_gn_trace(case):
    prob = util.load_case(case)
    ref = util.golden(case, "gn")
    op = O.OracleProblem(prob)
    errs = [op.error()]
    for _ in range(len(ref["gn_errors"]) - 1):
        st, e = op.gn_iterate()
        assert st == 0
        errs.append(e)
    assert np.allclose(errs, ref["gn_errors"], rtol=1e-8)


DL_CASES = ["pose2_ring", "bal_tiny_s2", "sphere_tiny", "sphere_small_colamd", "sphere_tiny_gaussian",
            "sphere_tiny_huber", "bal_tiny_bundler", "bal_tiny_body_sensor", "bal_tiny_colamd"]


@pytest.mark.parametrize("case", DL_CASES)
def test_oracle_dogleg_trace(case):
    """DoglegOptimizer::iterate trace of the reference (errors and trust-region radii, deltaInitial = 1)."""
    prob = util.load_case(case)
    ref = util.golden(case, "dl")
    op = O.OracleProblem(prob)
    e, d = op.error(), 1.0
    errs, deltas = [e], [d]
    for _ in range(len(ref["dl_errors"]) - 1):
        st, e, d = op.dogleg_iterate(e, d)
        assert st == 0
        errs.append(e)
        deltas.append(d)
    assert np.allclose(errs, ref["dl_errors"], rtol=1e-8)
    assert np.allclose(deltas, ref["dl_deltas"], rtol=1e-7)
    assert util.relmax(op.get_values(), ref["final_values"]) <= 1e-6


@pytest.mark.parametrize("case", ["bal_tiny_s2", "sphere_tiny", "sphere_tiny_gaussian", "bal_tiny_bundler", "pose2_ring"])
def test_oracle_marginal_covariances(case):
    """Marginals::marginalCovariance of every variable (gtsam/nonlinear/Marginals.cpp:118-154) against the
    unmodified reference (ref_harness marginals): groundwork for the device path of SURVEY 8f rank 3."""
    prob = util.load_case(case)
    ref = util.golden(case, "marg")["marg_cov"]
    op = O.OracleProblem(prob)
    off = 0
    for v in range(prob.nvars):
        d = P.VAR_DIM[int(prob.var_type[v])]
        st, S = op.marginal_covariance(v)
        assert st == 0
        R = ref[off:off + d * d].reshape(d, d).T
        off += d * d
        assert np.abs(S - R).max() <= 1e-8 * np.abs(R).max()
        assert np.allclose(S, S.T, rtol=1e-9, atol=1e-12 * np.abs(S).max())
    assert off == ref.size


@pytest.mark.parametrize("case", sorted(util.JOINT_SETS))
def test_oracle_joint_marginal_covariances(case):
    """Marginals::jointMarginalCovariance (two variables: BayesTree::joint; three: marginalMultifrontalBayesTree;
    gtsam/nonlinear/Marginals.cpp:128-190), full matrix in sorted-key order, against the unmodified reference."""
    prob = util.load_case(case)
    op = O.OracleProblem(prob)
    for i, vs in enumerate(util.JOINT_SETS[case]):
        ref = util.golden(case, f"joint{i}")["joint_cov"]
        st, S = op.joint_marginal_covariance(vs)
        assert st == 0
        R = ref.reshape(S.shape).T
        assert np.abs(S - R).max() <= 1e-8 * np.abs(R).max()


def test_oracle_solve_rhs_reproduces_delta():
    """H^-1 (A^T b) through the stored conditionals == the solve's own back-substituted delta."""
    prob = util.load_case("sphere_small_colamd")
    op = O.OracleProblem(prob)
    op.linearize()
    st, _, _, _ = op.solve(0.0)
    assert st == 0
    delta = op.get_delta()
    g = np.zeros(op.ndelta)
    dof = prob.dof_offsets()
    for gi, grp in enumerate(prob.groups):
        J = op.get_jacobians(gi)                       # (count, d, ncols), last column b
        for f in range(grp.count):
            col = 0
            for v in grp.keys[f]:
                if v < 0:
                    continue
                dv = P.VAR_DIM[int(prob.var_type[v])]
                g[dof[v]:dof[v] + dv] += J[f, :, col:col + dv].T @ J[f, :, -1]
                col += dv
    x = op.solve_rhs(g)
    assert np.linalg.norm(x - delta) <= 1e-9 * np.linalg.norm(delta)


def test_geometry_known_answers():
    """Pose3/Rot3 Expmap, Logmap, AdjointMap, inverse, compose against the reference,
    including the near-zero and near-pi branches (gtsam/geometry/SO3.cpp:264-319)."""
    k = util.golden("geometry", "kat") if False else None
    import os
    from oracle import refio
    k = refio.read_out(os.path.join(util.GOLDEN, "geometry_kat.bin"))
    xi = k["xi"].reshape(-1, 6)
    n = xi.shape[0]
    T = k["expmap"].reshape(n, 12)
    prev = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=float)
    for i in range(n):
        Ti = O.unary("orc_pose3_expmap", xi[i], 12)
        assert np.abs(Ti - T[i]).max() < 1e-13 * max(1, np.abs(T[i]).max())
        assert np.abs(O.unary("orc_pose3_logmap", T[i], 6) - k["logmap"].reshape(n, 6)[i]).max() < 1e-9
        assert np.abs(O.unary("orc_pose3_adjoint_map", T[i], 36) - k["adjoint"].reshape(n, 36)[i]).max() < 1e-12
        assert np.abs(O.unary("orc_pose3_inverse", T[i], 12) - k["inverse"].reshape(n, 12)[i]).max() < 1e-12
        assert np.abs(O.pose3_compose(prev, T[i]) - k["compose_prev"].reshape(n, 12)[i]).max() < 1e-12
        prev = T[i]
        assert np.abs(O.unary("orc_so3_expmap", k["so3_w"].reshape(n, 3)[i], 9) - k["so3_R"].reshape(n, 9)[i]).max() < 1e-14
        assert np.abs(O.unary("orc_so3_logmap", k["so3_R"].reshape(n, 9)[i], 3) - k["so3_log"].reshape(n, 3)[i]).max() < 1e-9


def test_cholesky_partial_known_answer():
    """gtsam/base/tests/testCholesky.cpp:26-68: 7x7 matrix, 3 frontal: R'R must
    reconstruct the leading rows and the trailing block is the Schur complement."""
    ABC = np.array([
        [4.0375, 3.4584, 3.5735, 2.4815, 2.1471, 2.7400, 2.2063],
        [0., 4.7267, 3.8423, 2.3624, 2.8091, 2.9579, 2.5914],
        [0., 0., 5.1600, 2.0797, 3.4690, 3.2419, 2.9992],
        [0., 0., 0., 1.8786, 1.0535, 1.4250, 1.3347],
        [0., 0., 0., 0., 3.0788, 2.6283, 2.3791],
        [0., 0., 0., 0., 0., 2.9227, 2.4056],
        [0., 0., 0., 0., 0., 0., 2.5776]])
    full = np.triu(ABC) + np.triu(ABC, 1).T
    ok, out = O.cholesky_partial(ABC, 3)
    assert ok
    R = np.triu(out[:3, :3])
    S = out[:3, 3:]
    assert np.abs(R.T @ R - full[:3, :3]).max() < 1e-9
    assert np.abs(R.T @ S - full[:3, 3:]).max() < 1e-9
    L = np.triu(out[3:, 3:])
    assert np.abs(L - np.triu(full[3:, 3:] - S.T @ S)).max() < 1e-9


def test_cholesky_partial_failure_modes():
    """gtsam/base/tests/testCholesky.cpp:70-140: negative pivot and underconstrained cases."""
    ok, _ = O.cholesky_partial(np.array([[1.0, 2.0], [2.0, 1.0]]), 2)      # indefinite
    assert not ok
    ok, _ = O.cholesky_partial(np.diag([1.0, 1e-30, 1.0]), 2)              # last two pivots 2^-50 apart
    assert not ok
    ok, _ = O.cholesky_partial(np.diag([1e-10, 1.0]), 1)                   # single tiny pivot
    assert not ok
    ok, _ = O.cholesky_partial(np.diag([2.0, 3.0, 4.0]), 2)
    assert ok


def test_projection_factor_known_answer():
    """gtsam/slam/tests/testProjectionFactor.cpp:96-163: Cal3_S2(fov=60,640,480), pose
    (0,0,-6) identity rotation, point origin, z=(323,240): error (-3,0) and the
    Jacobians H1, H2 given there."""
    fx = 320.0 / np.tan(np.radians(60.0) / 2)   # Cal3_S2(fov, w, h): fx = fy = w/2 / tan(fov/2)
    K = np.array([[fx, fx, 0.0, 320.0, 240.0]])
    pose = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, -6.0])
    prob = P.Problem(np.array([P.VAR_POSE3, P.VAR_POINT3]), np.concatenate([pose, np.zeros(3)]), np.array([1, 0]),
                     [P.FactorGroup(P.FACTOR_PROJECTION_CAL3S2, np.array([[0, 1]]), np.array([[323.0, 240.0]]),
                                    P.NOISE_UNIT)], K)
    op = O.OracleProblem(prob)
    op.linearize()
    J = op.get_jacobians(0)[0]
    H1 = np.array([[0., -554.256, 0., -92.376, 0., 0.], [554.256, 0., 0., 0., -92.376, 0.]])
    H2 = np.array([[92.376, 0., 0.], [0., 92.376, 0.]])
    assert np.abs(J[:, :6] - H1).max() < 1e-3
    assert np.abs(J[:, 6:9] - H2).max() < 1e-3
    assert np.abs(-J[:, 9] - np.array([-3.0, 0.0])).max() < 1e-9


def test_cheirality_semantics():
    """Point behind the camera: GenericProjectionFactor zeroes H and returns r = 2*fx
    (gtsam/slam/ProjectionFactor.h:156-165); GeneralSFMFactor zeroes H and b
    (gtsam/slam/GeneralSFMFactor.h:153-158)."""
    pose = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])
    K = np.array([[500.0, 500.0, 0.0, 320.0, 240.0]])
    prob = P.Problem(np.array([P.VAR_POSE3, P.VAR_POINT3]), np.concatenate([pose, [0.1, 0.2, -3.0]]), np.array([1, 0]),
                     [P.FactorGroup(P.FACTOR_PROJECTION_CAL3S2, np.array([[0, 1]]), np.array([[1.0, 2.0]]), P.NOISE_UNIT)], K)
    op = O.OracleProblem(prob)
    op.linearize()
    J = op.get_jacobians(0)[0]
    assert np.all(J[:, :9] == 0) and np.allclose(J[:, 9], -1000.0)
    assert abs(op.error() - 0.5 * 2 * 1000.0 ** 2) < 1e-9
    cam = np.concatenate([pose, [500.0, 0.0, 0.0, 0.0, 0.0]])
    prob = P.Problem(np.array([P.VAR_CAM_BUNDLER, P.VAR_POINT3]), np.concatenate([cam, [0.1, 0.2, -3.0]]), np.array([1, 0]),
                     [P.FactorGroup(P.FACTOR_SFM_BUNDLER, np.array([[0, 1]]), np.array([[1.0, 2.0]]), P.NOISE_UNIT)])
    op = O.OracleProblem(prob)
    op.linearize()
    assert np.all(op.get_jacobians(0) == 0) and op.error() == 0.0


def test_config1_reference_plumbing_record():
    """BASELINE.json configs[0]: Pose2SLAMExample_g2o on noisyToyGraph.txt with the reference built
    by oracle/Makefile (CPU, no GPU): the recorded run (tests/golden/config1_pose2slam_g2o.json,
    `ref_harness pose2`) reproduces BASELINE.md §2: 0.391637509949 -> 0.0685034664998 in 3 GN iterations."""
    import json
    import os
    rec = json.load(open(os.path.join(util.GOLDEN, "config1_pose2slam_g2o.json")))
    assert abs(rec["initial_error"] - 0.391637509949) < 1e-11
    assert abs(rec["gn_final_error"] - 0.0685034664998) < 1e-11 and rec["gn_iterations"] == 3
    assert abs(rec["lm_final_error"] - rec["gn_final_error"]) < 1e-4   # LM stops on its own tolerance
