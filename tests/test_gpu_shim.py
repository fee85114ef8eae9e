"""GPU: the GTSAM-side drop-in (gtsam_b200/shim: B200LevenbergMarquardtOptimizer, a subclass
of the reference's own optimizer) against the stock gtsam::LevenbergMarquardtOptimizer on
the same NonlinearFactorGraph / Values / Ordering — real GTSAM objects on both sides.
The driver binary (tests/shim_parity.cpp -> oracle/_ref/shim_parity) is built in the
container where /root/reference exists and travels to the GPU box."""
import json
import os
import subprocess

import numpy as np
import pytest

import util
from gtsam_b200 import datasets

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "shim_parity")


def run(path, maxit=30, ceres=0):
    out = subprocess.check_output([BIN, path, str(maxit), str(ceres)], timeout=600)
    return json.loads(out.decode().strip().splitlines()[-1])


def check(r, rtol):
    assert r["launches"] > 0
    assert len(r["dev_errors"]) == len(r["ref_errors"])
    assert np.allclose(r["dev_errors"], r["ref_errors"], rtol=rtol, atol=1e-10)
    assert np.allclose(r["dev_lambdas"], r["ref_lambdas"], rtol=1e-12)
    assert r["dev_inner"] == r["ref_inner"]
    assert r["linearize_max_rel_diff"] <= 1e-12
    # one hook call per iterate() of the unmodified base-class loop (an iterate() that gives up
    # at lambdaUpperBound does not increment iterations())
    assert r["hook_calls"] == len(r["dev_errors"]) - 1 >= r["optimize_iterations"]
    assert abs(r["optimize_error"] - r["dev_errors"][-1]) <= 1e-9 * max(1.0, abs(r["dev_errors"][-1]))


@pytest.mark.skipif(not os.path.exists(BIN), reason="shim_parity not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", ["bal_tiny_s2", "bal_tiny_bundler", "sphere_small_colamd", "dubrovnik_3_7_unit", "bal_tiny_body_sensor",
                                  "sphere_tiny_huber", "bal_tiny_tukey", "sphere_tiny_interleaved", "pose3example"])
def test_shim_matches_stock_optimizer(case):
    r = run(os.path.join(util.GOLDEN, f"{case}.prob.bin"), 100 if case.startswith("dub") else 30, int(case in util.CERES_CASES))
    check(r, 1e-5 if case.startswith("dub") else 1e-7)
    assert r["max_value_diff"] <= (1e-2 if case.startswith("dub") else 1e-6)
    if case == "dubrovnik_3_7_unit":   # tests/testGeneralSFMFactorB.cpp:62 through the drop-in
        assert abs(r["optimize_error"] - 0.0199833) < 1e-5


@pytest.mark.skipif(not os.path.exists(BIN), reason="shim_parity not built")
def test_shim_mid_size_bal(tmp_path):
    prob = datasets.make("bal_tiny", ncams=23, npoints=3000, visibility="scattered")
    path = str(tmp_path / "p.bin")
    prob.save(path)
    r = run(path, 10)
    check(r, 1e-7)
    assert r["max_value_diff"] <= 1e-6


MBIN = os.path.join(os.path.dirname(BIN), "shim_marginals")


@pytest.mark.skipif(not os.path.exists(MBIN), reason="shim_marginals not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", ["bal_tiny_s2", "sphere_tiny_gaussian"])
def test_shim_marginals_match_stock_marginals(case):
    """gtsam_b200::B200Marginals (C++ drop-in over b200_marginal_covariance / b200_joint_marginal_covariance)
    against gtsam::Marginals on real GTSAM objects: every variable's covariance, one information matrix, one
    3-variable joint with unsorted keys; and gtsam_b200::B200DoglegOptimizer against gtsam::DoglegOptimizer."""
    try:
        out = subprocess.run([MBIN, os.path.join(util.GOLDEN, f"{case}.prob.bin")], capture_output=True, text=True, timeout=300)
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except (subprocess.SubprocessError, OSError, ValueError, IndexError) as e:
        pytest.fail(f"shim_marginals: did not complete: {e}")
    if not (r["worst_cov"] <= 1e-7 and r["worst_info"] <= 1e-6 and r["worst_joint"] <= 1e-7):
        pytest.fail(f"shim_marginals: off: {r}")
    # B200DoglegOptimizer against the stock DoglegOptimizer (5 iterations: errors, trust-region radii, final values)
    if not (r["dogleg_error"] <= 1e-7 and r["dogleg_delta"] <= 1e-6 and r["dogleg_values"] <= 1e-6):
        pytest.fail(f"B200DoglegOptimizer: off: {r}")


@pytest.mark.skipif(not os.path.exists(MBIN), reason="shim_marginals not built (needs /root/reference at build time)")
def test_reference_gnc_template_with_device_lm():
    """The reference's own gtsam::GncOptimizer template instantiated with gtsam_b200::B200LevenbergMarquardtParams
    (OptimizerType = the device LM) against the stock GncOptimizer<GncParams<LevenbergMarquardtParams>> on a Pose3
    graph with corrupted edges: same weights, same solution."""
    try:
        out = subprocess.run([MBIN, os.path.join(util.GOLDEN, "sphere_tiny_outliers.prob.bin"), "1"], capture_output=True, text=True, timeout=600)
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except (subprocess.SubprocessError, OSError, ValueError, IndexError) as e:
        pytest.fail(f"GNC through the shim: did not complete: {e}")
    if not (0 <= r["gnc_weights"] <= 1e-4 and 0 <= r["gnc_values"] <= 1e-5):
        pytest.fail(f"GNC through the shim: off: {r}")
