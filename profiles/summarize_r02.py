"""Round 2: summarise the ncu outputs of profiles/collect_r02b.sh (run here, no GPU needed):
    python profiles/summarize_r02.py gpurun_out > profiles/r02_profile_summary.md
Also writes profiles/r02_kernel_traffic.json (dram bytes per launch of the kernels bench.py's roofline objects name)."""
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

D = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(n):
    return re.sub(r"\(.*", "", n).replace("b200::", "").replace("void ", "").strip()


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [short(x["Kernel Name"]) for x in rows]
    # the last complete LM iteration = from the last-but-one projection-factor linearize_kernel launch to the last one
    idx = [i for i, n in enumerate(names) if n.startswith("linearize_kernel<3") or n.startswith("linearize_kernel<4")]
    if len(idx) < 2:
        return "(no complete iteration in the launch list)", {}
    a, b = idx[-2], idx[-1]
    agg, tot = collections.OrderedDict(), 0.0
    for x, n in zip(rows[a:b], names[a:b]):
        n = re.sub(r"<.*", "", n)
        v = float(x["Metric Value"]) / 1e3
        e = agg.setdefault(n, [0, 0.0, 0.0])
        e[0] += 1; e[1] += v; e[2] = max(e[2], v); tot += v
    out = ["| kernel | launches | total us | max us | share |", "|---|---|---|---|---|"]
    for k, (c, v, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {c} | {v:.1f} | {m:.1f} | {100 * v / tot:.1f}% |")
    out.append(f"| **total** | {sum(e[0] for e in agg.values())} | {tot:.1f} | | |")
    return "\n".join(out), {k: v[1] / tot for k, v in agg.items()}


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__inst_executed.sum"]


def captures(rep):
    """rep: an .ncu-rep, or its raw page exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x.raw.csv`: the reports of
    several launches exceed what gpurun copies back)"""
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append((short(r[idx["Kernel Name"]]), {w: (r[idx[w]], units[idx[w]]) for w in WANT if w in idx}))
    return out


def main():
    print("# Round 2, final build: ncu evidence (one B200, `profiles/collect_r02f.sh`, commit 17da8ee)\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES with the bench's phase timers, not absolutes.\n")
    print("These captures are of commit 17da8ee — the second-pass kernels, BEFORE the projection groups were stored in leaf-visit order "
          "(commit 66544f1).  The 5.5 GB of DRAM reads of each leaf kernel at 10M factors below are what motivated that change; the bench "
          "records after it are profiles/r02_bench_1gpu_final.json (default) and r02_bench_1gpu_graph_order.json (B200_NO_FACTOR_REORDER=1); "
          "the GPU budget of the round did not allow a second ncu pass.\n")
    for path in sorted(glob.glob(os.path.join(D, "r02_launches_*.csv"))):
        w = os.path.basename(path)[len("r02_launches_"):-4]
        table, _ = launches(path)
        print(f"## Launch list of one LM iteration, {w}\n\n{table}\n")
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(D, "r02_*.ncu-rep")) + glob.glob(os.path.join(D, "r02_*.raw.csv"))):
        print(f"## Full captures: `{os.path.basename(rep)}` (`ncu --set full --import-source on --clock-control none`)\n")
        seen = set()
        for name, m in captures(rep):
            if name in seen:
                continue
            seen.add(name)
            print(f"### {name}\n\n| metric | value | unit |\n|---|---|---|")
            for k, (v, u) in m.items():
                print(f"| {k} | {v} | {u} |")
            print()
            try:
                fac = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
                rd = float(m["dram__bytes_read.sum"][0]) * fac.get(m["dram__bytes_read.sum"][1], 1.0)
                wr = float(m["dram__bytes_write.sum"][0]) * fac.get(m["dram__bytes_write.sum"][1], 1.0)
                traffic[re.sub(r"<.*", "", name) + ("<" + name.split("<", 1)[1] if "<" in name else "")] = {"dram_bytes": rd + wr, "report": os.path.basename(rep)}
            except Exception:
                pass
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "r02_ncu_dram_bytes.json"), "w"), indent=1)
    # per workload and bench phase (what bench.py's roofline objects report as `traffic`): dram bytes of ONE launch of the phase's kernel
    phase_of = {"front_df_kernel": "eliminate_large", "leaf_point_factor_kernel": "leaf_fused", "leaf_point_schur_mma_kernel": "leaf_schur",
                "leaf_point_schur_kernel": "leaf_schur", "linearize_kernel<3": "linearize", "linearize_kernel<4": "linearize",
                "linerr_kernel<3": "linear_error", "linerr_kernel<4": "linear_error", "error_kernel<3": "error", "error_kernel<4": "error"}
    by_phase = {}
    for name, rec in traffic.items():
        wl = re.sub(r"^r02[a-z]?_.*?_(bal_\w+?|sphere\w+?)\.(?:ncu-rep|raw\.csv)$", r"\1", rec["report"])
        for prefix, ph in phase_of.items():
            if name.startswith(prefix):
                by_phase.setdefault(wl, {})[ph] = rec["dram_bytes"]
    if "--traffic" in sys.argv:   # (profiles/r02_kernel_traffic.json is curated by hand: only captures that are still valid for the build)
        json.dump(by_phase, open(os.path.join(ROOT, "profiles", "r02_kernel_traffic.json"), "w"), indent=1)


main()
