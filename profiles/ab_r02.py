"""A/B of kernel variants on one B200 (round 2): one problem per environment combination, variants flipped in place with
b200_set_tuning; every variant is timed the same way (library phase timers over N LM iterations from the same initial
estimate + one CUDA-event pair around N un-profiled iterations) and must reproduce the baseline's error after the step.

    python profiles/ab_r02.py --workload bal_1m --iters 10 > gpurun_out/ab_bal_1m.json

ENVS: environment combinations read at problem creation (ticket order of front_df_kernel, run length of the point leaves);
TUNE: b200_set_tuning switches.  The first entry of each list is the shipped default."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="bal_1m")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--envs", default="default")           # ';'-separated K=V,K=V groups; "default" = no override
ap.add_argument("--tune", default="base")              # ';'-separated key=value,key=value groups; "base" = untouched
ap.add_argument("--jacobian", default="auto")
args = ap.parse_args()

import torch  # noqa: E402
from gtsam_b200 import capi, datasets, optimizer  # noqa: E402

ctx = capi.Context(0)
prob = datasets.make(args.workload)
jac32 = args.jacobian == "fp32" or (args.jacobian == "auto" and args.workload.startswith("bal_c5"))
stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", 0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"workload": args.workload, "jacobian_fp32": jac32, "iters": args.iters, "runs": []}


def parse_group(g):
    return [] if g in ("default", "base") else [kv.split("=") for kv in g.split(",")]


for eg in args.envs.split(";"):
    env = parse_group(eg)
    for k, v in env:
        os.environ[k] = v
    t0 = time.perf_counter()
    dev = capi.DeviceProblem(ctx, prob)
    if jac32:
        dev.set_jacobian_precision(True)
    dev.synchronize()
    setup_s = time.perf_counter() - t0
    for k, _ in env:
        os.environ.pop(k)
    lm = optimizer.LevenbergMarquardtOptimizer(ctx, prob, device_problem=dev)
    dev.save_values()
    for tg in args.tune.split(";"):
        for k, v in parse_group(tg):
            dev.set_tuning(k, int(v))

        def reset():
            dev.restore_values()
            capi._check(dev.L.b200_lm_reset(lm.h))

        for _ in range(3):
            reset(); lm.iterate()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(args.iters):       # each iteration on its own event pair, L2 flushed in between (as bench.py)
            reset()
            with torch.cuda.stream(stream):
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); lm.iterate(); b.record(stream)
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        for it in range(args.iters):
            reset()
            dev.profile_enable(1 if it == 0 else 2)
            lm.iterate()
            dev.synchronize()
            dev.profile_enable(0)
        prof = dev.profile()
        rec = {"env": eg, "tune": tg, "setup_s": round(setup_s, 2), "ms_per_iter": ms / args.iters, "error_after": lm.error(),
               "phases_ms": {k: round(v[0] / args.iters, 4) for k, v in prof.items() if v[0] > 0}}
        out["runs"].append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
    del lm
    dev.close()

base = out["runs"][0]["error_after"]
out["max_error_rel_diff"] = max(abs(r["error_after"] - base) / abs(base) for r in out["runs"])
print(json.dumps(out))
