"""GPU: the GaussianFactorGraph level of the C++ drop-in on real GTSAM objects (tests/shim_linear.cpp ->
oracle/_ref/shim_linear): gtsam_b200::optimizeOnDevice / B200LinearSolver against the reference's own
GaussianFactorGraph::optimize, and B200SolveLevenbergMarquardtOptimizer / B200SolveGaussNewtonOptimizer (the
NonlinearOptimizer::solve() seam on the device, linearize on the host) against the stock optimizers on a Pose2 graph —
the factor family of BASELINE.json configs[0], which is NOT one of the device-resident factor kinds.

Strict (the CPU side of the same level is pinned in tests/test_linear.py)."""
import json
import os
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "shim_linear")


def run(*args):
    out = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=420)
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.exists(BIN), reason="shim_linear not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", util.LINEAR_CASES)
def test_shim_optimize_on_device_matches_reference(case):
    try:
        r = run("graph", os.path.join(util.GOLDEN, f"{case}.lin.bin"))
    except (subprocess.SubprocessError, OSError, ValueError, IndexError) as e:
        pytest.fail(f"shim_linear graph: did not complete: {e}")
    if r["ref_status"] != r["dev_status"]:
        pytest.fail(f"shim_linear graph: status differs: {r}")
    if not 0 <= r.get("gradient_diff", -1) <= 1e-12:
        pytest.fail(f"shim_linear graph: gradientAtZero off: {r}")
    if r["ref_status"] == 0 and not (0 <= r["delta_rel_diff"] <= 1e-9 and 0 <= r["reuse_delta_rel_diff"] <= 1e-9
                                     and 0 <= r["bayes_tree_diff"] <= 1e-9 and 0 <= r["marginals_diff"] <= 1e-7
                                     and r["structure_builds"] == 1 and r["solves"] == 2 and r["launches"] > 0):
        pytest.fail(f"shim_linear graph: off: {r}")


@pytest.mark.skipif(not os.path.exists(BIN), reason="shim_linear not built (needs /root/reference at build time)")
def test_shim_solve_seam_on_pose2_graph():
    ref = json.load(open(os.path.join(util.GOLDEN, "pose2_synth_reference.json")))
    try:
        r = run("pose2", os.path.join(util.GOLDEN, "data", "synthetic_pose2.g2o"), 30)
    except (subprocess.SubprocessError, OSError, ValueError, IndexError) as e:
        pytest.fail(f"shim_linear pose2: did not complete: {e}")
    ok = (len(r["lm_dev_errors"]) == len(r["lm_ref_errors"]) and len(r["gn_dev_errors"]) == len(r["gn_ref_errors"])
          and np.allclose(r["lm_dev_errors"], r["lm_ref_errors"], rtol=1e-8) and np.allclose(r["gn_dev_errors"], r["gn_ref_errors"], rtol=1e-8)
          and r["lm_value_diff"] <= 1e-7 and r["gn_value_diff"] <= 1e-7 and r["lm_dev_inner"] == r["lm_ref_inner"]
          and abs(r["lm_ref_errors"][-1] - ref["lm_final_error"]) <= 1e-8 * ref["lm_final_error"]
          and r["launches"] > 0 and r["solves"] >= len(r["lm_dev_errors"]) - 1)
    if not ok:
        pytest.fail(f"shim_linear pose2: off: {r}")


FBIN = os.path.join(ROOT, "oracle", "_ref", "shim_families")


@pytest.mark.skipif(not os.path.exists(FBIN), reason="shim_families not built (needs /root/reference at build time)")
def test_shim_solve_seam_on_rank4_factor_families():
    """GeneralSFMFactor2, smart projection factors (HESSIAN mode) and expression factors: the stock
    LevenbergMarquardtOptimizer against B200SolveLevenbergMarquardtOptimizer (GTSAM linearizes, the device solves)."""
    try:
        out = subprocess.run([FBIN, "gpu"], capture_output=True, text=True, timeout=420)
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except (subprocess.SubprocessError, OSError, ValueError, IndexError) as e:
        pytest.fail(f"shim_families gpu: did not complete: {e}")
    bad = {k: v for k, v in r.items() if not (v["worst_error_rel_diff"] <= 1e-7 and v["value_diff"] <= 1e-6 and v["launches"] > 0)}
    if bad:
        pytest.fail(f"shim_families gpu: off: {bad}")
