"""Profiling aid: B200_DF_TRACE=<file> python profiles/df_trace.py <workload> -> per-tile globaltimer stamps of one solve of
front_df_kernel; prints the timeline of the tiles on the critical path (the diagonal tiles of the widest front)."""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtsam_b200 import capi, datasets

def main():
    name = sys.argv[1]
    path = os.environ["B200_DF_TRACE"]
    prob = datasets.make(name)
    ctx = capi.Context(0)
    dev = capi.DeviceProblem(ctx, prob)
    dev.linearize()
    for _ in range(3):
        dev.solve(1e-5)
    raw = open(path, "rb").read()
    nt = struct.unpack("q", raw[:8])[0]
    tasks = np.frombuffer(raw[8:8 + 16 * nt], dtype=np.int32).reshape(nt, 4)
    st = np.frombuffer(raw[8 + 16 * nt:], dtype=np.uint64).reshape(nt, 32)
    code = (st >> np.uint64(56)).astype(np.int64)
    ns = (st & np.uint64((1 << 56) - 1)).astype(np.int64)
    t0 = ns[(ns > 0) & (code < 20) & (code > 0)].min()
    print("tasks", nt, "span us", (ns[(code < 20) & (code > 0)].max() - t0) / 1e3)
    fronts, counts = np.unique(tasks[:, 0], return_counts=True)
    big = fronts[np.argmax(counts)]
    print("front", big, "tiles", counts.max())
    for c in fronts:     # per front: first start, children arrived, last end
        sel = np.where(tasks[:, 0] == c)[0]
        arrived = [(ns[t, e] - t0) / 1e3 for t in sel for e in range(32) if code[t, e] == 2]
        ends = [(ns[t, e] - t0) / 1e3 for t in sel for e in range(32) if code[t, e] in (13, 14)]
        print("front %d tiles %d: children arrived %.1f .. %.1f us, finished %.1f us" % (c, len(sel), min(arrived), max(arrived), max(ends)))
    sel = np.where(tasks[:, 0] == big)[0]
    for t in sel:
        c, j, r, _ = tasks[t]
        if r != j // 4 or j > 12:
            continue    # diagonal-row tiles of the first columns
        ev = [(int(code[t, e]), (ns[t, e] - t0) / 1e3 if code[t, e] < 20 else int(ns[t, e])) for e in range(32) if code[t, e]]
        print("tile j=%d r=%d:" % (j, r), " ".join(("%d@%.1f" % e) if e[0] < 20 else ("[%s %d cyc]" % ("chol" if e[0] == 20 else "trsm", e[1])) for e in ev))

main()
