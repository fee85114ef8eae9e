"""The WHOLE library executed on the CPU: gtsam_b200/csrc/engine.cu + kernels.cuh compiled by g++ with -DB200_EMULATE
against tests/emu/cuda_emu_full.h (every CUDA thread a fiber of one host thread, __syncthreads / __syncwarp / warp shuffles as
barriers + exchanges, cp.async a plain copy) and tests/emu/cuda_fake_runtime.cpp (device memory = host
memory) — the same source the GPU build compiles, through the same C-ABI and the same Python mirror, against the same
golden vectors of the unmodified reference.

It is how the code written after this round's GPU budget was spent gets exercised end to end before its first hardware
run: the GaussianFactorGraph level (JacobianFactor / HessianFactor groups), the FP32-storage mode through the float
instantiations of the leaf kernels (cp.async staging included), the Pose2 factor family, METIS-ordered BAL, the joint
marginal kernel — next to paths that WERE validated on the B200 (the BAL point-leaf kernels, the panel / update chain,
the flag-chained back-substitution, Dogleg, Gauss-Newton, LM), which makes the emulation itself credible.  It checks
logic and arithmetic, not the GPU: no memory-model subtleties, no performance.  The scenario groups run as parallel
processes (tests/emu/run_scenarios.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "_build", "libgtsam_b200_emu.so")
CSRC = os.path.join(ROOT, "gtsam_b200", "csrc")

GROUPS = [
    # every typed golden case: both reference dumps (every stage incl. all conditionals) + the LM trace
    ["typed:bal_tiny_s2", "typed:bal_tiny_bundler", "typed:bal_tiny_colamd", "typed:bal_tiny_body_sensor", "typed:bal_tiny_tukey",
     "typed:bal_tiny_fair", "typed:bal_small_metis", "typed:pose2_ring", "typed:pose2_ring_colamd"],
    ["typed:sphere_tiny", "typed:sphere_tiny_gaussian", "typed:sphere_tiny_interleaved", "typed:pose3example"],
    ["typed:sphere_small_colamd", "typed:sphere_tiny_cauchy"],
    ["typed:sphere_small_metis", "typed:sphere_tiny_huber", "typed:dubrovnik_3_7_unit", "typed:dubrovnik_3_7_priors"],
    # FP32-storage mode (float instantiations, cp.async staging of floats in the point-leaf Schur kernel)
    ["fp32:bal_tiny_s2", "fp32:bal_tiny_bundler", "fp32:bal_small_metis", "fp32:sphere_tiny_gaussian", "fp32:pose2_ring",
     "marginals:bal_tiny_s2", "marginals:sphere_tiny", "marginals:bal_tiny_bundler", "marginals:pose2_ring"],
    # the GaussianFactorGraph level, Dogleg, Gauss-Newton
    ["linear:" + c for c in ("lin_pose2_toy", "lin_pose2_synth", "lin_random_nary", "lin_mixed_hessian", "lin_arity8", "lin_sphere_tiny",
                             "lin_bal_tiny", "lin_singular", "lin_family_sfm2", "lin_family_smart", "lin_family_expr")] +
    ["mirror:x", "dogleg:bal_tiny_s2", "dogleg:sphere_tiny", "dogleg:pose2_ring", "gn:sphere_tiny", "gn:pose2_ring"],
]


@pytest.fixture(scope="module")
def emu_lib():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "symbolic.cpp")] + [os.path.join(EMU, "cuda_fake_runtime.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("kernels.cuh", "engine.cuh", "factors.cuh", "geometry.cuh", "symbolic.h")] + \
        [os.path.join(EMU, "cuda_emu_full.h"), os.path.join(ROOT, "include", "gtsam_b200.h")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-w", "-fPIC", "-shared", "-DB200_EMULATE", "-x", "c++", "-I/usr/local/cuda/include",
                               "-I", EMU] + srcs + ["-o", LIB, "-pthread", "-ldl"], cwd=CSRC)
    return LIB


def test_whole_library_in_host_emulation(emu_lib):
    procs = [subprocess.Popen([sys.executable, os.path.join(EMU, "run_scenarios.py"), emu_lib] + g, stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for g in GROUPS]
    failures = []
    for g, p in zip(GROUPS, procs):
        try:
            out, err = p.communicate(timeout=900)
        except subprocess.TimeoutExpired:
            p.kill()
            failures.append((g, "timeout"))
            continue
        done = [l.split()[1] for l in out.splitlines() if l.startswith("EMU_OK")]
        if p.returncode != 0 or done != g:
            failures.append((g, done, err[-800:]))
    assert not failures, failures
