"""Inputs added after the last hardware run of the main -m gpu suite (util.EXTRA_CASES): planar pose graphs
(BetweenFactor<Pose2> / PriorFactor<Pose2>, BASELINE.json configs[0]'s factor family, device-resident since the end of
round 1) and BAL graphs eliminated in the reference's METIS nested-dissection order (BASELINE.json configs[3] names METIS;
gtsam_b200/data_bal_*_metis.npz hold the reference's own Ordering::Metis for the large workloads): the small golden case
against the reference's dump / LM trace, and the 1M-factor workload through the size-independent property of
tests/test_gpu_parity.py (delta satisfies the damped normal equations, the reported linear errors are what they say).

The kernels involved are the validated ones (BAL with a COLAMD ordering and Pose3 graphs with METIS orderings are in the
main suite).  Own process; any mismatch, crash or timeout fails the suite with the subprocess's stderr."""
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
import test_gpu_parity as TP
from gtsam_b200 import capi, datasets, optimizer
ctx = capi.Context(0)
for case in util.EXTRA_CASES:
    prob = util.load_case(case)
    for kind, lam, diag in (("dump0", 0.0, 0), ("dump1", 1e-2, 1)):
        dev = capi.DeviceProblem(ctx, prob)
        util.check_against_dump(dev, prob, util.golden(case, kind), lam, diag)
        dev.close()
    ref = util.golden(case, "lm")
    lm = optimizer.LevenbergMarquardtOptimizer(ctx, prob)
    lm.optimize()
    assert abs(lm.error() - ref["lm_errors"][-1]) <= 1e-7 * ref["lm_errors"][-1]
    del lm
prob = datasets.make("bal_1m_metis")
dev = capi.DeviceProblem(ctx, prob)
dev.linearize()
st, e0, e1, _ = dev.solve(1e-5)
assert st == 0
res, A, b, dl = TP._normal_equation_residual(prob, dev, 1e-5)
assert res <= 1e-9, res
r = A @ dl - b
assert abs(e0 - 0.5 * b @ b) <= 1e-11 * e0 and abs(e1 - 0.5 * r @ r) <= 1e-9 * e0
assert dev.try_step() < dev.error()
info = dev.symbolic_info()
print("ORDERINGS_OK", res, info.ncliques, info.nlevels, ctx.launch_count())
"""


def test_cuda_metis_ordered_bal_isolated():
    try:
        out = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT)], capture_output=True, text=True, timeout=420)
    except subprocess.TimeoutExpired:
        pytest.fail("METIS-ordered BAL: timed out")
    lines = [l for l in out.stdout.splitlines() if l.startswith("ORDERINGS_OK")]
    if not lines:
        pytest.fail("METIS-ordered BAL: did not complete: " + out.stderr[-3000:])
    assert int(lines[-1].split()[4]) > 0


def test_shim_pose2_graph_matches_stock_optimizer():
    """The C++ drop-in (B200LevenbergMarquardtOptimizer, fully device-resident) on a Pose2 pose graph against the stock
    gtsam::LevenbergMarquardtOptimizer on the same GTSAM objects (oracle/_ref/shim_parity)."""
    import json
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    binp = os.path.join(ROOT, "oracle", "_ref", "shim_parity")
    if not os.path.exists(binp):
        pytest.skip("shim_parity not built")
    try:
        out = subprocess.run([binp, os.path.join(util.GOLDEN, "pose2_ring_colamd.prob.bin"), "30", "0"], capture_output=True, text=True, timeout=420)
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except (subprocess.SubprocessError, OSError, ValueError, IndexError) as e:
        pytest.fail(f"shim_parity on a Pose2 graph: did not complete: {e}")
    ok = (len(r["dev_errors"]) == len(r["ref_errors"]) and np.allclose(r["dev_errors"], r["ref_errors"], rtol=1e-7, atol=1e-10)
          and r["dev_inner"] == r["ref_inner"] and r["linearize_max_rel_diff"] <= 1e-12 and r["max_value_diff"] <= 1e-6 and r["launches"] > 0)
    if not ok:
        pytest.fail(f"shim_parity on a Pose2 graph: off: {r}")
