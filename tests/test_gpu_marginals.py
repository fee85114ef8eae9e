"""Marginals::marginalCovariance on the device against the unmodified reference:
every variable of four problems (BAL with Cal3_S2 / Bundler cameras, Pose3 graphs with diagonal / full
Gaussian noise), the reuse of the undamped factor across variables and its invalidation by a damped
solve.  Runs in its own process: the path walk of marginal_path_kernel is the newest kernel of the
library and a fault there must not poison the CUDA context of the other GPU tests.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["bal_tiny_s2", "sphere_tiny", "sphere_tiny_gaussian", "bal_tiny_bundler"]

SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import util
from gtsam_b200 import capi, optimizer, problem as P
ctx = capi.Context(0)
worst = 0.0
for case in {cases!r}:
    prob = util.load_case(case)
    ref = util.golden(case, "marg")["marg_cov"]
    m = optimizer.Marginals(ctx, prob)
    off = 0
    for v in range(prob.nvars):
        d = P.VAR_DIM[int(prob.var_type[v])]
        S = m.marginalCovariance(v)
        R = ref[off:off + d * d].reshape(d, d).T
        off += d * d
        worst = max(worst, float(np.abs(S - R).max() / np.abs(R).max()))
    # the factor is reused across variables and invalidated by a damped solve
    m.dp.linearize(); m.dp.solve(1e-3)
    S2 = m.marginalCovariance(0)
    d0 = P.VAR_DIM[int(prob.var_type[0])]
    worst = max(worst, float(np.abs(S2 - ref[:d0 * d0].reshape(d0, d0).T).max() / np.abs(ref[:d0 * d0]).max()))
print("MARGINALS_WORST", worst)
"""


def test_cuda_marginal_covariances_isolated():
    script = SCRIPT.format(root=ROOT, cases=CASES)
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    lines = [l for l in out.stdout.splitlines() if l.startswith("MARGINALS_WORST")]
    assert lines, out.stderr[-3000:]
    worst = float(lines[-1].split()[1])
    assert worst <= 1e-7, worst


JOINT_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import util
from gtsam_b200 import capi, optimizer
ctx = capi.Context(0)
worst = 0.0
for case, sets in util.JOINT_SETS.items():
    prob = util.load_case(case)
    m = optimizer.Marginals(ctx, prob)
    for i, vs in enumerate(sets):
        ref = util.golden(case, "joint%d" % i)["joint_cov"]
        jm = m.jointMarginalCovariance(vs)
        S = jm.fullMatrix()
        R = ref.reshape(S.shape).T
        worst = max(worst, float(np.abs(S - R).max() / np.abs(R).max()))
        a, b = sorted(vs)[0], sorted(vs)[-1]
        worst = max(worst, float(np.abs(jm.at(a, a) - m.marginalCovariance(a)).max() / np.abs(R).max()))
        assert jm.at(a, b).shape[0] == jm.at(a, a).shape[0] and np.allclose(jm.at(a, b), jm.at(b, a).T)
print("JOINT_WORST", worst)
"""


def test_cuda_joint_marginal_covariances_isolated():
    """Marginals::jointMarginalCovariance on the device against the unmodified reference (2- and 3-variable sets).
    The algorithm is pinned on the CPU oracle (test_oracle_joint_marginal_covariances); own process, strict."""
    script = JOINT_SCRIPT.format(root=ROOT)
    try:
        out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    except subprocess.TimeoutExpired:
        pytest.fail("device joint marginals: timed out")
    lines = [l for l in out.stdout.splitlines() if l.startswith("JOINT_WORST")]
    if not lines:
        pytest.fail("device joint marginals: did not complete: " + out.stderr[-3000:])
    worst = float(lines[-1].split()[1])
    if not worst <= 1e-7:
        pytest.fail(f"device joint marginals: off by {worst:.3g} (tolerance 1e-7)")
