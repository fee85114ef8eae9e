"""Summarise ncu outputs brought back in gpurun_out/ (run here, no GPU needed)."""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [x["Kernel Name"] for x in rows]
    idx = [i for i, n in enumerate(names) if "linearize_kernel" in n]
    start = idx[-2]
    agg, tot = collections.OrderedDict(), 0.0
    for x in rows[start:]:
        n = re.sub(r"\(.*", "", re.sub(r"<.*", "", x["Kernel Name"])).replace("b200::", "").replace("void ", "")
        v = float(x["Metric Value"]) / 1e3
        a = agg.setdefault(n, [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v); tot += v
    out = ["| kernel | launches / LM iteration | total us | max us | share |", "|---|---|---|---|---|"]
    for k, (c, v, m) in agg.items():
        out.append(f"| {k} | {c} | {v:.1f} | {m:.1f} | {100 * v / tot:.1f}% |")
    out.append(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | | |")
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "launch__grid_size", "launch__block_size"]


def full(path):
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    r = list(csv.reader(raw.splitlines()))
    h, u, v = r[0], r[1], r[2]
    out = ["| metric | unit | value |", "|---|---|---|"]
    for i, n in enumerate(h):
        if n in WANT:
            out.append(f"| {n} | {u[i]} | {v[i]} |")
    return "\n".join(out)


if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    print(launches(path) if kind == "launches" else full(path))
