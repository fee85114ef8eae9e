"""The C++ drop-ins (gtsam_b200/shim: subclasses of the reference's own optimizers, B200Marginals, the GaussianFactorGraph
level) driven against the HOST-EMULATION build of the library (tests/test_library_emulation.py): the parity drivers that
tests/test_gpu_shim*.py run on the B200 are started here with LD_PRELOAD=libgtsam_b200_emu.so, so the b200_* symbols
resolve to the emulated library, and held to the same thresholds — real GTSAM objects on both sides, the stock optimizer
as the reference.  """
import json
import os
import subprocess

import numpy as np
import pytest

import util
from test_library_emulation import emu_jobs, emu_lib, emu_libs  # noqa: F401  (fixtures: build the emulated library, start every emulation job at once)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "shim_parity")), reason="shim drivers not built (need /root/reference at build time)")


def test_cpp_drop_ins_against_the_emulated_library(emu_jobs):  # noqa: F811
    r = {}
    for k in ("lm_bal", "lm_bundler", "lm_pose2", "lm_huber", "marg", "lin_nary", "lin_hess", "lin_sing", "pose2", "families", "gnc"):
        rc, out, err = emu_jobs["shim_" + k]
        assert rc == 0, (k, err[-600:])
        r[k] = json.loads(out.strip().splitlines()[-1])
    for k in ("lm_bal", "lm_bundler", "lm_pose2", "lm_huber"):      # B200LevenbergMarquardtOptimizer vs the stock optimizer
        x = r[k]
        assert x["launches"] > 0 and len(x["dev_errors"]) == len(x["ref_errors"])
        assert np.allclose(x["dev_errors"], x["ref_errors"], rtol=1e-7, atol=1e-10) and np.allclose(x["dev_lambdas"], x["ref_lambdas"], rtol=1e-12)
        assert x["dev_inner"] == x["ref_inner"] and x["linearize_max_rel_diff"] <= 1e-12 and x["max_value_diff"] <= 1e-6
        assert x["hook_calls"] == len(x["dev_errors"]) - 1 >= x["optimize_iterations"]
    m = r["marg"]                                                     # B200Marginals, B200DoglegOptimizer
    assert m["worst_cov"] <= 1e-7 and m["worst_info"] <= 1e-6 and m["worst_joint"] <= 1e-7
    assert m["dogleg_error"] <= 1e-7 and m["dogleg_delta"] <= 1e-6 and m["dogleg_values"] <= 1e-6
    for k in ("lin_nary", "lin_hess"):                                # optimizeOnDevice, B200LinearSolver, Bayes tree, linear Marginals
        x = r[k]
        assert x["ref_status"] == x["dev_status"] == 0 and 0 <= x["delta_rel_diff"] <= 1e-9 and 0 <= x["reuse_delta_rel_diff"] <= 1e-9
        assert 0 <= x["bayes_tree_diff"] <= 1e-9 and 0 <= x["marginals_diff"] <= 1e-7 and x["structure_builds"] == 1 and x["solves"] == 2
        assert 0 <= x["gradient_diff"] <= 1e-12          # B200LinearSolver::gradientAtZero vs gfg.gradientAtZero()
    assert r["lin_sing"]["ref_status"] == r["lin_sing"]["dev_status"] == 1     # IndeterminantLinearSystemException on both sides
    p2 = r["pose2"]                                                   # the solve() seam on a Pose2 graph
    assert np.allclose(p2["lm_dev_errors"], p2["lm_ref_errors"], rtol=1e-8) and np.allclose(p2["gn_dev_errors"], p2["gn_ref_errors"], rtol=1e-8)
    assert p2["lm_value_diff"] <= 1e-7 and p2["gn_value_diff"] <= 1e-7 and p2["lm_dev_inner"] == p2["lm_ref_inner"] and p2["structure_builds"] == 1
    # the reference's own GncOptimizer template instantiated with B200LevenbergMarquardtParams vs the stock instantiation
    assert 0 <= r["gnc"]["gnc_weights"] <= 1e-4 and 0 <= r["gnc"]["gnc_values"] <= 1e-5
    for name, x in r["families"].items():                             # GeneralSFMFactor2, smart factors, expression factors
        assert x["worst_error_rel_diff"] <= 1e-7 and x["value_diff"] <= 1e-6 and x["launches"] > 0 and x["builds"] == 1, (name, x)


def test_sharded_cpp_drop_in_against_the_emulated_library(emu_lib, tmp_path):  # noqa: F811
    """B200LevenbergMarquardtOptimizer with a B200Communicator (multi-GPU through the C++ drop-in, one process per rank): two
    ranks of oracle/_ref/shim_parity over the emulated library and tests/emu/fake_nccl.cpp — identical error / lambda /
    inner-iteration traces to the stock optimizer on every rank, and values() is the FULL estimate on every rank."""
    build = os.path.dirname(emu_lib)
    nccl = os.path.join(build, "libnccl.so.2")
    src = os.path.join(ROOT, "tests", "emu", "fake_nccl.cpp")
    if not os.path.exists(nccl) or os.path.getmtime(nccl) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", src, "-o", nccl, "-lrt", "-pthread"])
    env = dict(os.environ, LD_PRELOAD=emu_lib, B200_NO_GRAPH="1", LD_LIBRARY_PATH=build + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    for case in ("bal_tiny_s2", "sphere_small_colamd"):
        uid = str(tmp_path / (case + ".uid"))
        procs = [subprocess.Popen([os.path.join(REF, "shim_parity"), os.path.join(util.GOLDEN, case + ".prob.bin"), "12", "0", "0", "2", str(r), uid, "0"],   # (device 0 for both: the emulated runtime has one)
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for r in range(2)]
        for r, p in enumerate(procs):
            try:
                out, err = p.communicate(timeout=300)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            assert p.returncode == 0, (case, r, err[-800:])
            x = json.loads(out.strip().splitlines()[-1])
            assert x["world"] == 2 and x["rank"] == r and x["launches"] > 0
            assert len(x["dev_errors"]) == len(x["ref_errors"]) and np.allclose(x["dev_errors"], x["ref_errors"], rtol=1e-7, atol=1e-10)
            assert np.allclose(x["dev_lambdas"], x["ref_lambdas"], rtol=1e-12) and x["dev_inner"] == x["ref_inner"]
            assert x["max_value_diff"] <= 1e-6, (case, r, x["max_value_diff"])
