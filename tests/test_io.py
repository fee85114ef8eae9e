"""CPU: the dataset readers (gtsam_b200/io.py) against the reference's own loaders.  The text
files under tests/golden/data/ are synthetic (written by gtsam_b200.io); the *.prob.bin next to
them were produced by feeding those files to the unmodified reference (readG2o /
SfmData::FromBalFile through oracle/ref_harness g2ofile|balfile, see make_golden.py)."""
import os

import numpy as np

import util
from gtsam_b200 import io, problem as P
from oracle import oracle_py as O


def test_g2o_reader_matches_reference_loader():
    ref = util.load_case("synthetic_sphere_g2o")
    mine = io.read_g2o_3d(os.path.join(util.GOLDEN, "data", "synthetic_sphere.g2o"), ordering=ref.ordering)
    assert np.array_equal(mine.var_type, ref.var_type)
    assert np.abs(mine.values - ref.values).max() <= 1e-15
    for a, b in zip(mine.groups, ref.groups):
        assert a.type == b.type and np.array_equal(a.keys, b.keys)
        assert np.abs(a.meas - b.meas).max() <= 1e-15
        assert np.abs(a.noise.ravel() - b.noise.ravel()).max() <= 1e-12 * np.abs(b.noise).max()
    # and the problem means the same thing to the oracle
    assert abs(O.OracleProblem(mine).error() - O.OracleProblem(ref).error()) <= 1e-12 * O.OracleProblem(ref).error()


def test_bal_reader_matches_reference_loader():
    """Includes the reference's quirk of parsing every value through `float` (SfmData.cpp:206-240)."""
    ref = util.load_case("synthetic_bal")
    mine = io.read_bal(os.path.join(util.GOLDEN, "data", "synthetic_bal.txt"), noise_sigma=1.0, priors=True)
    assert np.array_equal(mine.var_type, ref.var_type)
    assert np.abs(mine.values - ref.values).max() <= 4e-16 * max(1.0, np.abs(ref.values).max())
    assert len(mine.groups) == len(ref.groups) == 3
    for a, b in zip(mine.groups, ref.groups):
        assert a.type == b.type and np.array_equal(a.keys, b.keys) and a.noise_kind == b.noise_kind
        assert np.abs(a.meas - b.meas).max() <= 4e-16 * max(1.0, np.abs(b.meas).max())
        assert np.allclose(a.noise, b.noise)


def test_pose3example_initial_error_golden():
    """examples/Pose3SLAMExample_g2o.cpp on examples/Data/pose3example.txt prints
    'initial error=64941.322888' (BASELINE.md §2); the fixture is that file through the reference's parser."""
    prob = util.load_case("pose3example")
    assert abs(O.OracleProblem(prob).error() - 64941.322888) < 1e-5
    assert abs(util.golden("pose3example", "dump0")["error"][0] - 64941.322888) < 1e-5
