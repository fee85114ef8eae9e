"""Round 2: where a tile of front_df_kernel spends its slot — intervals between consecutive stamps of the per-tile trace
(B200_DF_TRACE=<file> python profiles/df_trace.py <workload>), summed over all tiles.
    python profiles/df_trace_intervals.py gpurun_out/df_trace_c5.bin [more.bin ...] > profiles/r02_df_trace_summary.md
Stamp codes (front_df.cuh): 1 ticket taken, 2 children arrived, 3 R_kk seen (pivot rows in the tile), 4 R_kk staged, 5 solve done,
6 piece published, 7 pieces of a step above the tile seen, 8 staged, 9 rank-32 update done, 10 diagonal block: Cholesky starts,
11 done, 12 published, 13 extend-add issued, 14 arrival signalled (tile ends)."""
import struct
import sys

import numpy as np

NAMES = {(9, 7): "wait for the next step's pieces (above the tile)", (7, 8): "stage the pieces (cp.async)", (8, 9): "rank-32 update (DMMA)",
         (2, 7): "load C + wait for the first pieces", (9, 13): "extend-add into the parent (FP64 atomics)", (13, 14): "barrier + fence + arrival counter",
         (1, 2): "wait for the children", (2, 3): "load C + wait for R_kk (pivot rows in the tile)", (4, 5): "triangular solve of the piece",
         (6, 9): "wait for row pieces + update (pivot rows in the tile)", (3, 4): "stage R_kk", (5, 6): "publish the piece", (9, 3): "next pivot block in the tile",
         (10, 11): "Cholesky of the diagonal block", (2, 10): "load C (diagonal tile of column 0)", (12, 13): "invert R_kk (off the chain)", (11, 12): "publish R_kk"}
for path in sys.argv[1:]:
    raw = open(path, "rb").read()
    nt = struct.unpack("q", raw[:8])[0]
    st = np.frombuffer(raw[8 + 16 * nt:], dtype=np.uint64).reshape(nt, 32)
    code = (st >> np.uint64(56)).astype(np.int64)
    ns = (st & np.uint64((1 << 56) - 1)).astype(np.int64)
    valid = (code > 0) & (code < 20)
    T = (ns - ns[valid].min()) / 1e3
    d = {}
    for t in range(nt):
        pc = pt = None
        for e in range(32):
            c = code[t, e]
            if c == 0 or c >= 20:
                continue
            if pc is not None:
                d.setdefault((int(pc), int(c)), []).append(T[t, e] - pt)
            pc, pt = c, T[t, e]
    tot = sum(np.sum(v) for v in d.values())
    print(f"## `{path}`: {nt} tiles, span {T[valid].max():.0f} us, {tot / 1e3:.0f} ms of slot time accounted (tiles with more than 32 stamps are cut)\n")
    print("| from -> to | what | count | p10 us | p50 us | p90 us | sum ms | share |")
    print("|---|---|---|---|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -np.sum(kv[1]))[:14]:
        x = np.array(v)
        p = np.percentile(x, [10, 50, 90])
        print(f"| {k[0]} -> {k[1]} | {NAMES.get(k, '')} | {len(x)} | {p[0]:.2f} | {p[1]:.2f} | {p[2]:.2f} | {x.sum() / 1e3:.1f} | {100 * x.sum() / tot:.0f}% |")
    print()
