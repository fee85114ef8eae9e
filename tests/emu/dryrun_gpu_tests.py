"""Dry run of the `-m gpu` test CODE on a machine without a GPU: the same test functions, but gtsam_b200.capi loads the
host-emulation build of the library (tests/emu/_build/libgtsam_b200_emu.so, built by tests/test_library_emulation.py).
It checks that the GPU tests themselves are sound (fixtures, shapes, tolerances that do not depend on the device) before
they meet hardware.  TEST INFRASTRUCTURE; not part of either pytest run.

    python tests/emu/dryrun_gpu_tests.py [pytest args, default: the in-process GPU test files]
"""
import os
import sys

os.environ["B200_NO_GRAPH"] = "1"          # CUDA graphs are not emulated
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gtsam_b200 import capi  # noqa: E402

capi.LIB_PATH = os.path.join(ROOT, "tests", "emu", "_build", "libgtsam_b200_emu.so")
import pytest  # noqa: E402

args = sys.argv[1:] or [os.path.join(ROOT, "tests", f) for f in ("test_gpu_parity.py", "test_gpu_marginals.py")]
# (tests that need the CUDA runtime through torch — pinned host buffers, bench — fail here by design)
sys.exit(pytest.main(["-m", "gpu", "-q", "-p", "no:cacheprovider"] + args))
