#!/usr/bin/env python
"""bench.py — LM iterations/sec (linearize + damped multifrontal solve + retract + error)
on the BAL-style workload of BASELINE.json, through the C-ABI library.

    python bench.py --gpus N --steps K --warmup W [--workload bal_c3] [--impl reference]

One "step" = one LevenbergMarquardtOptimizer::iterate() from the same initial
estimate (values restored, lambda reset), i.e. one linearize + >=1 damped
factor/solve/retract/error tries.  `value` times it with the inputs resident in
HBM; `e2e` times the same step through the public call with HOST buffers
(host->device copy of the Values, device->host read-back of the new Values and
the error inside the timed region).  `--impl reference` times the UNMODIFIED
reference (oracle/_ref, built from /root/reference by oracle/Makefile; falls
back to the plain-C oracle port when that binary is absent) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="bal_c3")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush-l2", action="store_true", help="time the K steps back to back with L2 left warm")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="N > 1: auto = BAL workloads scale weakly (N x the points), others strongly; strong = the named workload "
                         "itself sharded over the N ranks (BASELINE configs[3]/[4]: the 3M / 10M-factor graphs on 4 / 8 GPUs)")
    ap.add_argument("--jacobian-fp32", action="store_true",
                    help="BASELINE configs[4]'s precision mix: whitened Jacobians stored as floats (FP32 linearize output), FP64 solve")
    return ap.parse_args()


METRIC = "LM iterations/sec (linearize+solve) on BAL-style graph"
UNIT = "iterations/s"


def workload_config(prob, args):
    from gtsam_b200 import problem as P
    return {
        "workload": f"{args.workload}: {prob.name}", "factors": prob.nfactors, "variables": prob.nvars,
        "factor_types": sorted({int(g.type) for g in prob.groups}),
        "ordering": ("Schur (points, then cameras)" if prob.meta.get("ordering", "schur") == "schur" else
                     f"{prob.meta['ordering'].upper()} by the reference's Ordering::Create (shipped as data)") if prob.meta.get("kind") == "bal"
        else prob.meta.get("ordering", "natural"),
        "lm_params": "LevenbergMarquardtParams::LegacyDefaults (lambda0=1e-5, factor 10)",
        "jacobian_storage": "fp32 (b200_set_jacobian_precision: FP64 evaluation, float [A|b], FP64 solve)" if getattr(args, "jacobian_fp32", False) else "fp64",
        "cache": ("L2 left warm between iterations (--no-flush-l2)" if getattr(args, "no_flush_l2", False) else
                  "L2 flushed between timed iterations: 256 MiB memset on the stream, outside the per-iteration CUDA-event pairs"),
        "parallelism": "single GPU" if args.gpus == 1 else
        f"{args.gpus} ranks, one per GPU: junction-tree subtrees (BAL: points) + their factors sharded by rank, top of the tree "
        f"replicated, one in-place NCCL all-reduce of the top fronts per solve; BAL workloads scale weakly: the graph has "
        f"{args.gpus}x the points of the 1-GPU workload and `value` counts 1-GPU-sized units (iterations/s x {args.gpus}); "
        f"other workloads: the same graph sharded (strong scaling, value = iterations/s)",
    }


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    @staticmethod
    def _epoch(ts):
        import datetime
        try:
            return datetime.datetime.strptime(ts.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self, region=None):
        """region = (epoch start, epoch end) of the timed loop: only samples taken inside it count (a short
        region may hold none: then all samples of the run, warm-up included, are used and `in_region` is 0)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.rows
        in_region = 0
        if region is not None:
            inside = [r for r in rows if len(r) > 7 and self._epoch(r[7]) is not None
                      and region[0] - 0.02 <= self._epoch(r[7]) <= region[1] + 0.02]
            in_region = len(inside)
            if inside:
                rows = inside
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "in_region": in_region, "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
def reference_time(prob, steps, warmup):
    """Times LevenbergMarquardtOptimizer::iterate() of the reference on the host cores.
    Returns (seconds per iterate, dict describing the run)."""
    from oracle import refio
    ncores = os.cpu_count()
    if refio.have_ref():
        try:
            r = refio.time_lm(prob, steps, warmup)
            return r["mean_s"], {"kind": "reference", "cores": 1, "host_cores": ncores,
                                 "sample": f"{steps} x LevenbergMarquardtOptimizer::iterate() of the unmodified reference "
                                           f"(oracle/_ref, -O3, no TBB in this image => single thread) on the full workload, "
                                           f"{warmup} warm-up; linearize {r['linearize_s']:.3f}s, damped solve {r['solve_s']:.3f}s",
                                 "error_after": r["error_after"], "inner_iterations": r["inner_iterations"]}
        except Exception as e:  # binary present but not runnable on this box
            note = f"oracle/_ref not runnable here ({type(e).__name__}); "
    else:
        note = "oracle/_ref absent; "
    from gtsam_b200 import problem as P
    from oracle import oracle_py as O
    import ctypes as C
    op = O.OracleProblem(prob)
    v0 = prob.values.copy()
    prm = P.CLMParams(100, 1e-5, 1e-5, 0.0, 1e-5, 10.0, 1e5, 0.0, 1e-3, 0, 1, 1e-6, 1e32)
    ts = []
    for s in range(warmup + steps):
        op.set_values(v0)
        lm = op.lm(prm)
        t0 = time.perf_counter()
        op.lm_iterate(lm)
        if s >= warmup:
            ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), {"kind": "port", "cores": 1, "host_cores": ncores,
                               "sample": note + f"{steps} x LM iterate of the plain-C oracle port on the full workload"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gtsam_b200 import datasets
    prob = datasets.make(args.workload)
    # bounded: the reference needs 1.7 s (bal_c3) .. 40 s (bal_c4) per iterate
    est = {"bal_c3": 2.0, "bal_1m": 9.0, "bal_c4": 40.0, "bal_c5": 200.0, "sphere2500": 0.3,
           "bal_1m_metis": 9.0, "bal_c4_metis": 40.0, "bal_c5_metis": 200.0}.get(args.workload, 1.0)
    steps = max(1, min(args.steps, int(150.0 / est)))
    warmup = min(args.warmup, 1 if est > 1 else 3)
    sec, info = reference_time(prob, steps, warmup)
    val = 1.0 / sec
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference", "config": workload_config(prob, args),
            "cpu_baseline": dict(info, value=val, unit=UNIT),
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np
    import torch
    from gtsam_b200 import capi, datasets, optimizer

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — gtsam_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    over = {}
    weak = False
    if world > 1:   # weak scaling: per-GPU points fixed, cameras fixed (BAL); other workloads: strong scaling
        base = datasets.WORKLOADS[args.workload][1]
        if "npoints" in base and args.scaling != "strong" and base.get("ordering", "schur") == "schur":   # stored orderings fit one size
            over["npoints"] = base["npoints"] * world
            weak = True
    prob = datasets.make(args.workload, **over)
    ctx = capi.Context(local)
    if world > 1:
        ids = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
    dev = capi.DeviceProblem(ctx, prob)
    if args.jacobian_fp32:
        dev.set_jacobian_precision(True)
    jb = 4 if args.jacobian_fp32 else 8
    lm = optimizer.LevenbergMarquardtOptimizer(ctx, prob, device_problem=dev)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    L = dev.L
    # the step's host buffers are page-locked (contract: inputs come from pinned host memory)
    host_values = torch.from_numpy(prob.values.copy()).pin_memory().numpy()
    host_out = torch.empty(host_values.size, dtype=torch.float64).pin_memory().numpy()
    dev.save_values()

    def step_resident():
        dev.restore_values()                    # D2D: inputs already in HBM
        capi._check(L.b200_lm_reset(lm.h))      # state <- (values, lambda0); recomputes graph.error
        lm.iterate()

    def step_e2e():
        dev.set_values(host_values)             # pinned host -> device
        capi._check(L.b200_lm_reset(lm.h))
        lm.iterate()
        out = dev.get_values(host_out)          # device -> pinned host
        return out, lm.error()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # L2 flush between timed iterations: the working set of bal_c3 (Jacobians 48 MB + fronts 52 MB + tables) is about
    # the size of the 126 MB L2, so each iteration is timed on its own (event pair on the library's stream) and a
    # 256 MiB memset on the same stream evicts L2 in between, outside the event pairs.
    flush_buf = None if args.no_flush_l2 else torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", local))

    def flush_l2():
        with torch.cuda.stream(stream):
            flush_buf.zero_()

    def timed(fn, steps, warmup, flush=True):
        for _ in range(warmup):
            fn()
        barrier()
        flush = flush and flush_buf is not None
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps if flush else 1)]
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        timed.region = [time.time(), None]
        if flush:
            for a, b in pairs:
                flush_l2()
                a.record(stream)
                fn()
                b.record(stream)
        else:
            pairs[0][0].record(stream)
            for _ in range(steps):
                fn()
            pairs[0][1].record(stream)
        barrier()
        wall = time.perf_counter() - t0
        timed.region[1] = time.time()
        ms = sum(a.elapsed_time(b) for a, b in pairs)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, ctx.launch_count() - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, wall, launches = timed(step_resident, args.steps, max(3, args.warmup))
    clocks = sampler.stop(tuple(timed.region)) if rank == 0 else None
    ms_e2e, wall_e2e, _ = timed(step_e2e, args.steps, 1)
    ms_warm, _, _ = timed(step_resident, args.steps, 1, flush=False)   # information only: L2 left warm between iterations

    # phase profile (separate pass; event records add ~1 us per phase)
    dev.profile_enable(True)
    for _ in range(args.steps):
        step_resident()
    dev.synchronize()
    prof = dev.profile()
    dev.profile_enable(False)
    st = lm._state()

    if rank != 0:
        return
    units = world if weak else 1      # weak scaling: each iteration processes `world` 1-GPU-sized graphs
    value = units * args.steps / (ms * 1e-3)
    e2e = units * args.steps / (ms_e2e * 1e-3)
    info = dev.symbolic_info()
    peak, peak_src = measured_peaks()
    per_step = {k: (v[0] / args.steps, v[1] / args.steps) for k, v in prof.items()}
    # algorithmic bytes per step of the HBM-bound phases
    lin_bytes = prob.linearize_bytes(jb)
    jac_bytes = sum(g.count * jb * P_ncols(g) for g in prob.groups)
    alg_bytes = {
        "linearize": lin_bytes,                                  # 184 B/projection factor + Values
        "memset_fronts": info.front_bytes,
        "assemble": jac_bytes + info.front_bytes,                # read [A|b] once, touch every front entry once
        "eliminate_small": 2 * info.front_bytes,                 # read each front, write [R S d] + Schur update
        "eliminate_large": 2 * info.front_bytes,
        "leaf_fused": jac_bytes + info.front_bytes,              # read [A|b] of the leaf factors, write [R S d]
        # Schur SYRK of the point leaves: read [S' d'] (3 x DC per factor + 3 per point) and [A_c b] (2 x DC + 2 per factor)
        "leaf_schur": schur_bytes(prob, jb),
    }
    tries = max(1.0, per_step["leaf_fused"][1] if per_step["leaf_fused"][1] else per_step["back_substitute"][1])
    # phases that are ONE kernel launch (per group): the candidates for "the dominant kernel"
    single = {"linearize": "linearize_kernel", "leaf_fused": "leaf_point_factor_kernel / leaf_fused_kernel",
              "leaf_schur": "leaf_point_schur_kernel",
              "memset_fronts": "memset", "assemble": "assemble_kernel", "linear_error": "linerr_kernel", "error": "error_kernel"}
    dom = max(single, key=lambda k: per_step[k][0])
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")
    if os.path.exists(tpath) and args.workload == "bal_c3" and world == 1:
        traffic = json.load(open(tpath))

    def roof(name):
        """achieved = algorithmic bytes of one launch / its average duration (CUDA events on the
        launching stream, from the library's phase timers); linearize runs once per step, the
        solve phases once per lambda try.  traffic = dram read+write of one launch from ncu --set full."""
        ms_phase, _calls = per_step[name]
        if ms_phase <= 0 or name not in alg_bytes:
            return None
        units = 1.0 if name == "linearize" else tries
        nbytes = alg_bytes[name] / world     # each rank handles its shard
        ach = nbytes * units / (ms_phase * 1e-3) / 1e9
        return {"kernel": single.get(name, name), "phase": name, "bound": "hbm", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic.get(name, {}).get("dram_bytes"),
                "peak_source": peak_src, "ms_per_launch": ms_phase / units, "algorithmic_bytes_per_launch": nbytes,
                "note": "FP64 path: latency/instruction bound at this size (see profiles/); tensor pipe unused (no FP64 tcgen05 kind)"}

    large_ms = per_step["eliminate_large"][0] / tries
    large = {"flops_per_solve": info.factor_flops, "ms_per_solve": large_ms,
             "achieved_tflops": (info.factor_flops / (large_ms * 1e-3) / 1e12) if large_ms > 0 else None,
             "note": "panel Cholesky + TRSM + rank-k updates of the non-leaf fronts, FP64 FMA pipe"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if (weak or world == 1) else "strong",
        "vs_baseline": None, "dtype": "f32 Jacobians + f64 solve" if args.jacobian_fp32 else "f64", "data": "synthetic", "config": workload_config(prob, args),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(host_values.nbytes),
                "d2h_bytes_per_step": int(host_values.nbytes) + 64, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "warm_l2": {"ms_per_step": ms_warm / args.steps, "value": units * args.steps / (ms_warm * 1e-3),
                    "note": "same K steps back to back without the L2 flush (information only)"},
        "roofline": roof(dom) or roof("linearize"),
        "roofline_linearize": roof("linearize"),
        "large_fronts": large,
        "phases_ms_per_step": {k: round(v[0], 4) for k, v in per_step.items()},
        "lm": {"error_after": st.error, "lambda_after": st.lambda_, "tries_per_step": tries},
        "tree": {"cliques": info.ncliques, "levels": info.nlevels, "max_frontal": info.max_frontal_dim,
                 "max_separator": info.max_separator_dim, "factor_flops": info.factor_flops,
                 "front_bytes": info.front_bytes},
        "wall_ms_per_step": wall * 1e3 / args.steps,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            est = {"bal_c3": 2.0, "bal_1m": 9.0, "bal_c4": 40.0, "bal_1m_metis": 9.0, "bal_c4_metis": 40.0, "bal_c5_metis": 200.0}.get(args.workload, 1.0)
            n = max(1, min(5, int(20.0 / est)))
            sec, cinfo = reference_time(prob, n, 1 if est <= 10.0 else 0)   # 3M / 10M factors: 30-250 s per reference iterate
            line["cpu_baseline"] = dict(cinfo, value=1.0 / sec, unit=UNIT)
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def schur_bytes(prob, jb=8):
    """Algorithmic bytes of leaf_point_schur_kernel: read [S' d'] (3 x DC per factor + 3 per point, FP64) and
    [A_c b] (2 x DC + 2 per factor, jb bytes each); the run's extend-add output is negligible."""
    from gtsam_b200 import problem as P
    total = 24 * int((prob.var_type == P.VAR_POINT3).sum())
    for g in prob.groups:
        if g.type in (P.FACTOR_PROJECTION_CAL3S2, P.FACTOR_SFM_BUNDLER):
            dc = P.factor_ncols(g.type) - 4       # camera dofs: ncols = DC + 3 + 1
            total += g.count * (8 * 3 * dc + jb * (2 * dc + 2))
    return total


def P_ncols(g):
    from gtsam_b200 import problem as P
    return P.FACTOR_DIM[g.type] * P.factor_ncols(g.type)


if __name__ == "__main__":
    main()
