"""The "FP32 linearize + FP64 solve" mode on the device (b200_set_jacobian_precision, BASELINE.json configs[4]):
the float instantiations of the linearize / assemble / hessianDiagonal / linear-error / leaf kernels against the
oracle's restatement of the mode and against the FP64 reference at the FP32 protocol of SURVEY 8(c) ([A|b] rel <= 1e-5,
final error rel <= 1e-5); switching back restores the FP64 Jacobians bit for bit; a mid-size BAL problem exercises the
point-leaf kernels (cp.async staging of float operands) and the CUDA graph of the LM try.

The default (FP64) instantiations are unchanged by the templating (identical SASS before / after, checked when it was
written).  Own process; strict: any mismatch, crash or timeout fails the suite with stderr (CPU side:
tests/test_precision.py)."""
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
from gtsam_b200 import capi, datasets, optimizer
from oracle import oracle_py as O
ctx = capi.Context(0)
wj = wd = we = 0.0
for case in util.CASES:
    prob = util.load_case(case)
    ref = util.golden(case, "dump0")
    dev, orc = capi.DeviceProblem(ctx, prob), O.OracleProblem(prob)
    dev.linearize()
    st64 = dev.solve(0.0)[0]
    d64, j64 = dev.get_delta(), dev.get_jacobians(0)
    dev.set_jacobian_precision(True); orc.set_jacobian_precision(True)
    dev.linearize(); orc.linearize()
    for gi in range(len(prob.groups)):
        J = dev.get_jacobians(gi)
        assert np.array_equal(J, J.astype(np.float32).astype(np.float64))
        wj = max(wj, util.relmax(J, util.ref_jacobians(prob, ref, gi)), util.relmax(J, orc.get_jacobians(gi)))
    assert util.relmax(dev.hessian_diagonal(), orc.hessian_diagonal()) <= 1e-6
    st, e0, e1, _ = dev.solve(1e-2, True)
    so, f0, f1, _ = orc.solve(1e-2, True)
    assert st == so, (case, st, so)
    if st == 0:
        wd = max(wd, util.rel2(dev.get_delta(), orc.get_delta()))
        assert abs(e0 - f0) <= 1e-6 * f0 and abs(e1 - f1) <= 1e-5 * f0, (case, e0, f0, e1, f1)
    # back to FP64: bit-identical to the run before the switch
    dev.set_jacobian_precision(False)
    dev.linearize()
    assert dev.solve(0.0)[0] == st64 and np.array_equal(dev.get_jacobians(0), j64), case
    # (an indeterminate undamped system leaves no delta; assembly adds with FP64 atomics, so delta is reproducible to
    # rounding only — amplified by cond(H): undamped dubrovnik-3-7 has cond ~ 7e15, util.ILL_CONDITIONED.  This bound
    # at 1e-10 for every case is what failed on the round-1 hardware run: a test tolerance, not a kernel fault.)
    # Round 2: 1e-10 failed again on hardware (bal_tiny_colamd: 1.9e-10) once the per-run Schur complements moved to the
    # tensor-path kernel — same atomics, another order; the undamped BAL systems held by two priors have cond ~ 1e7, so
    # rounding-level differences of H show up at ~1e-9 in delta.  1e-8 = the conditioning-limited bound of SURVEY 8(c).)
    assert st64 != 0 or util.rel2(dev.get_delta(), d64) <= 1e-8 * util.ILL_CONDITIONED.get((case, 0.0), 1.0) ** 2, (case, util.rel2(dev.get_delta(), d64))
    dev.close()
    # LM to convergence with float Jacobians vs the FP64 reference's optimum
    prm = optimizer.LevenbergMarquardtParams.CeresDefaults() if case in util.CERES_CASES else optimizer.LevenbergMarquardtParams()
    dev2 = capi.DeviceProblem(ctx, prob)
    dev2.set_jacobian_precision(True)
    lm = optimizer.LevenbergMarquardtOptimizer(ctx, prob, prm, device_problem=dev2)
    lm.optimize()
    r = util.golden(case, "lm")["lm_errors"][-1]
    if case != "sphere_tiny_huber":      # see tests/test_precision.py: SLOW_CONVERGENCE
        we = max(we, abs(lm.error() - r) / r)
    del lm
    dev2.close()
assert wj <= 1e-6 and wd <= 1e-4 and we <= 1e-5, (wj, wd, we)
# mid-size BAL (point-leaf kernels with many runs, CUDA graph of the try): device vs oracle, both with float Jacobians
for model in ("cal3_s2", "bundler"):
    prob = datasets.make("bal_tiny", ncams=23, npoints=3000, visibility="scattered", camera_model=model)
    dev, orc = capi.DeviceProblem(ctx, prob), O.OracleProblem(prob)
    dev.set_jacobian_precision(True); orc.set_jacobian_precision(True)
    dev.linearize(); orc.linearize()
    st, e0, e1, _ = dev.solve(1e-3)
    so, f0, f1, _ = orc.solve(1e-3)
    assert st == so == 0
    assert util.rel2(dev.get_delta(), orc.get_delta()) <= 1e-5
    lm = optimizer.LevenbergMarquardtOptimizer(ctx, prob, device_problem=dev)
    olm = orc.lm(lm.params()._c)
    for _ in range(3):
        lm.iterate(); orc.lm_iterate(olm)
        assert abs(lm.error() - olm.state.error) <= 1e-6 * olm.state.error
    del lm
    dev.close()
print("F32_OK", wj, wd, we, ctx.launch_count())
"""


def test_cuda_fp32_storage_mode_isolated():
    try:
        out = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT)], capture_output=True, text=True, timeout=420)
    except subprocess.TimeoutExpired:
        pytest.fail("FP32-storage mode: timed out")
    lines = [l for l in out.stdout.splitlines() if l.startswith("F32_OK")]
    if not lines:
        pytest.fail("FP32-storage mode: did not complete: " + out.stderr[-3000:])
    assert int(lines[-1].split()[4]) > 0
