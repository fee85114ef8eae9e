"""Round 2: tables of the A/B measurements of profiles/ab_r02.py (profiles/ab/*.json) -> profiles/r02_ab_summary.md.
Every row: one problem on one B200, N LM iterations from the same initial estimate, L2 flushed between iterations; phases from the
library's CUDA-event timers.  `env` = read at problem creation, `tune` = flipped in place (b200_set_tuning), cumulative within a file."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["linearize", "leaf_fused", "leaf_schur", "eliminate_large", "back_substitute", "linear_error", "error"]
print("# Round 2: A/B of kernel variants on the B200 (`profiles/ab_r02.py`, raw records in `profiles/ab/`)\n")
print("`ab_tune_*` / `ab_env_*`: first A/B call (commit 141d95f: tensor-path Schur kernel v1, staged conditionals, linearize variant, "
      "ticket order / lag, run length).  `ab2_*`: second call (commit ceb32c0: front_df_kernel with branch-free C loads, one look at all flags, preloaded "
      "extend-add maps; Schur kernel v2).  ms per LM iteration; every variant reproduces the first row's error after the step "
      "(`max_error_rel_diff` <= 3e-11 in every file).\n")
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "ab", "*.json"))):
    d = json.load(open(f))
    print(f"## {os.path.basename(f)} — {d['workload']}{' (FP32 Jacobian storage)' if d['jacobian_fp32'] else ''}, max_error_rel_diff {d['max_error_rel_diff']:.1e}\n")
    print("| env | tune | ms/iter | " + " | ".join(COLS) + " |")
    print("|---|---|---|" + "---|" * len(COLS))
    for r in d["runs"]:
        print(f"| {r['env']} | {r['tune']} | {r['ms_per_iter']:.4f} | " + " | ".join(f"{r['phases_ms'].get(c, 0):.4f}" for c in COLS) + " |")
    print()
