"""Which kernels of the GPU build do the emulated scenarios reach?  TEST INFRASTRUCTURE.

    python tests/emu/kernel_coverage.py        (after tests/test_library_emulation.py has built tests/emu/_build/)

Runs every scenario group of tests/test_library_emulation.py plus one sharded world with B200_EMU_TRACE_FILE set and
compares the traced kernel names with the kernel symbols of the sm_100a build (gtsam_b200/libgtsam_b200.so).  Prints the kernels never launched in emulation."""
import importlib.util
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "_build", "libgtsam_b200_emu.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"^void ", "", n).split("(")[0] for n in out]


def main():
    spec = importlib.util.spec_from_file_location("tle", os.path.join(ROOT, "tests", "test_library_emulation.py"))
    tle = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tle)
    trace = tempfile.NamedTemporaryFile(suffix=".txt", delete=False).name
    env = dict(os.environ, B200_EMU_TRACE_FILE=trace, LD_LIBRARY_PATH=os.path.dirname(LIB) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    jobs = [subprocess.Popen([sys.executable, os.path.join(EMU, "run_scenarios.py"), LIB] + g, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
            for g in tle.GROUPS]
    world = 2
    uid = (b"/b200emu_cov_%d" % os.getpid()).ljust(128, b"\0").hex()
    jobs += [subprocess.Popen([sys.executable, os.path.join(EMU, "run_sharded.py"), LIB, str(r), str(world), uid], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL, env=env) for r in range(world)]
    rcs = [j.wait() for j in jobs]
    counts = {}
    for line in open(trace):
        name, n = line.rsplit(" ", 1)
        counts[name] = counts.get(name, 0) + int(n)
    os.unlink(trace)
    reached = dict(zip(demangle(list(counts)), counts.values()))
    # the __global__ functions of the GPU build = the host-side kernel symbols of libgtsam_b200.so
    syms = subprocess.run(["nm", "--defined-only", os.path.join(ROOT, "gtsam_b200", "libgtsam_b200.so")], capture_output=True, text=True).stdout
    built = sorted({n for n in demangle([l.split()[-1] for l in syms.splitlines() if l.split()[-1].startswith("_ZN4b200")])
                    if re.match(r"b200::[a-z0-9_]*_kernel(<.*>)?$", n)})
    missing = [k for k in built if k not in reached]
    print("job exit codes:", rcs)
    print(f"{len(built)} kernels in the sm_100a build, {len(built) - len(missing)} launched in emulation, {len(missing)} never:")
    for k in missing:
        print("  ", k)
    return 0


if __name__ == "__main__":
    sys.exit(main())
