"""GaussianFactorGraph level on the device (b200_linear_create): the CUDA path against the UNMODIFIED reference's own
GaussianFactorGraph::optimize(ordering, EliminatePreferCholesky) on the same JacobianFactors (tests/golden/lin_*.bin):
whitening, hessianDiagonal, delta, the two linear errors, the Bayes-tree cliques and every conditional [R S d], for
lambda = 0 and the damped system; marginal covariances; b200_linear_update; the GaussianFactorGraph mirror.

The elimination / back-substitution kernels are the validated ones of the nonlinear path; new here are
jacobian_load_kernel and the assemble / hessianDiagonal / linear-error kernels of the JacobianFactor and
HessianFactor groups.  The CPU side (oracle + host symbolic phase on n-ary factors) is pinned in
tests/test_linear.py; this check lives in its own process and fails the suite (with stderr) on any mismatch.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import util
from gtsam_b200 import capi, linear as LN, problem as P
ctx = capi.Context(0)
n = 0
for case in util.LINEAR_CASES:
    lp = util.load_linear_case(case)
    for which, lam in util.LINEAR_LAMBDA.items():
        dev = capi.LinearDeviceProblem(ctx, lp)
        util.check_linear_against_reference(dev, lp, util.golden(case, "out%d" % which), lam)
        dev.close()
        n += 1
# marginal covariances from the undamped factor
worst = 0.0
for case in ("lin_pose2_toy", "lin_random_nary", "lin_arity8"):
    lp = util.load_linear_case(case)
    ref = util.golden(case, "out0")["marginal_covariances"]
    dev = capi.LinearDeviceProblem(ctx, lp)
    off = 0
    for v in range(lp.nvars):
        d = int(lp.var_dim[v])
        S = dev.marginal_covariance(v)
        R = ref[off:off + d * d].reshape(d, d).T
        off += d * d
        worst = max(worst, float(np.abs(S - R).max() / np.abs(R).max()))
    dev.close()
assert worst <= 1e-7, worst
# new numbers, same structure
lp = util.load_linear_case("lin_random_nary")
dev = capi.LinearDeviceProblem(ctx, lp)
rng = np.random.default_rng(1)
g = lp.groups[3]
newAb = g.Ab + 0.01 * rng.normal(size=g.Ab.shape)
dev.update(3, newAb, g.sigmas)
assert dev.solve(0.0)[0] == 0
lp2 = LN.LinearProblem(lp.var_dim, lp.ordering, [LN.JacobianGroup(h.rows, h.dims, h.keys, newAb if i == 3 else h.Ab, h.sigmas,
                                                                   h.graph_index0, h.graph_index) for i, h in enumerate(lp.groups)])
dev2 = capi.LinearDeviceProblem(ctx, lp2)
assert dev2.solve(0.0)[0] == 0
assert util.rel2(dev.get_delta(), dev2.get_delta()) <= 1e-12
# HessianFactor groups: new information matrices, same structure
lpm = util.load_linear_case("lin_mixed_hessian")
devm = capi.LinearDeviceProblem(ctx, lpm)
newinfo = lpm.hgroups[1].info * 1.01
devm.update_hessian(1, newinfo)
assert devm.solve(0.0)[0] == 0
lpm2 = LN.LinearProblem(lpm.var_dim, lpm.ordering, lpm.groups,
                        [LN.HessianGroup(g.dims, g.keys, newinfo if i == 1 else g.info, g.graph_index0, g.graph_index) for i, g in enumerate(lpm.hgroups)])
devm2 = capi.LinearDeviceProblem(ctx, lpm2)
assert devm2.solve(0.0)[0] == 0 and util.rel2(devm.get_delta(), devm2.get_delta()) <= 1e-12
devm.close(); devm2.close()
# the calls that need Values are refused, loudly
try:
    dev.error()
    raise SystemExit("b200_error accepted a linear problem")
except capi.B200Error as e:
    assert e.code == P.INVALID_ARGUMENT
dev.close(); dev2.close()
# object-level mirror: GaussianFactorGraph.optimize(ordering)
lp = util.load_linear_case("lin_pose2_toy")
ref = util.golden("lin_pose2_toy", "out0")
gfg = LN.GaussianFactorGraph()
flat = {{}}
for g in lp.groups:
    pos = g.graph_index if g.graph_index is not None else g.graph_index0 + np.arange(g.count)
    for i in range(g.count):
        M, blocks, c = g.Ab[i].T, [], 0
        for d in g.dims:
            blocks.append(M[:, c:c + d]); c += d
        flat[int(pos[i])] = LN.JacobianFactor([int(k) for k in g.keys[i]], blocks, M[:, c], None if g.sigmas is None else g.sigmas[i])
for pos in sorted(flat):
    gfg.add(flat[pos])
x = gfg.optimize([int(v) for v in lp.ordering], ctx)
got = np.concatenate([x[k] for k in sorted(x)])
assert util.rel2(got, ref["delta"]) <= 1e-9
print("LINEAR_OK", n, worst, ctx.launch_count())
"""


def test_cuda_linear_level_matches_reference_isolated():
    try:
        out = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT)], capture_output=True, text=True, timeout=420)
    except subprocess.TimeoutExpired:
        pytest.fail("device GaussianFactorGraph level: timed out")
    lines = [l for l in out.stdout.splitlines() if l.startswith("LINEAR_OK")]
    if not lines:
        pytest.fail("device GaussianFactorGraph level: did not complete: " + out.stderr[-3000:])
    assert int(lines[-1].split()[1]) == 2 * len(__import__('util').LINEAR_CASES) and int(lines[-1].split()[3]) > 0
