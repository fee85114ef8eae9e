"""Launched with torchrun (one rank per GPU): the sharded solve against the
single-GPU solve of the same problem.  Prints MULTI_GPU_OK on rank 0 when everything matches.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtsam_b200 import capi, datasets, optimizer, problem as P  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = capi.Context(local)
    ctx.comm_init(ids[0], rank, world)
    solo_ctx = capi.Context(local)          # same GPU, no communicator: the single-GPU answer
    ok = True
    for name, kw in (("bal_tiny", dict(ncams=23, npoints=3000, visibility="scattered")),
                     ("bal_tiny", dict(ncams=40, npoints=4000, visibility="banded", camera_model="bundler")),
                     ("sphere2500", {}),                                  # COLAMD: non-leaf subtrees per rank
                     ("sphere_tiny", dict(layers=14, per_ring=24))):          # a chain: (almost) everything is top
        prob = datasets.make(name, **kw)
        co, fo = capi.shard_plan(prob, world)
        sh, solo = capi.DeviceProblem(ctx, prob), capi.DeviceProblem(solo_ctx, prob)
        e_sh, e_solo = sh.error(), solo.error()
        ok &= abs(e_sh - e_solo) <= 1e-12 * e_solo
        sh.linearize(); solo.linearize()
        for lam, diag in ((1e-3, False), (1e-2, True)):
            st, a0, a1, _ = sh.solve(lam, diag)
            so, b0, b1, _ = solo.solve(lam, diag)
            ok &= st == so == 0 and abs(a0 - b0) <= 1e-12 * b0 and abs(a1 - b1) <= 1e-9 * b0
            d_sh, d_solo = sh.get_delta(), solo.get_delta()
            # this rank's view: its own leaf variables + the replicated top variables
            fp, fv, sp, sv, par = solo.cliques()
            dof = prob.dof_offsets()
            mine = np.zeros(d_solo.size, dtype=bool)
            for c in range(len(par)):
                if co[c] in (-1, rank):
                    for v in fv[fp[c]:fp[c + 1]]:
                        mine[dof[v]:dof[v + 1]] = True
            ok &= np.linalg.norm(d_sh[mine] - d_solo[mine]) <= 1e-7 * np.linalg.norm(d_solo[mine])
            ok &= np.all(d_sh[~mine] == 0)
            ok &= abs(sh.try_step() - solo.try_step()) <= 1e-9 * e_solo
        lm_sh = optimizer.LevenbergMarquardtOptimizer(ctx, prob, device_problem=sh)
        lm_solo = optimizer.LevenbergMarquardtOptimizer(solo_ctx, prob, device_problem=solo)
        for _ in range(4):
            lm_sh.iterate(); lm_solo.iterate()
            ok &= abs(lm_sh.error() - lm_solo.error()) <= 1e-8 * lm_solo.error()
            ok &= lm_sh.lambda_() == lm_solo.lambda_() and lm_sh.getInnerIterations() == lm_solo.getInnerIterations()
        sh.close(); solo.close()
    flag = torch.tensor([int(bool(ok))], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if flag.item() == 1 else "MULTI_GPU_MISMATCH", "world", world)
    # the GaussianFactorGraph level sharded the same way (written after the last 2-GPU run: reported on its own line,
    # it does not change the verdict above until it has run on hardware once)
    lin_ok = True
    try:
        import util
        for name in ("lin_sphere_tiny", "lin_bal_tiny", "lin_random_nary", "lin_mixed_hessian"):
            lp = util.load_linear_case(name)
            sh, solo = capi.LinearDeviceProblem(ctx, lp), capi.LinearDeviceProblem(solo_ctx, lp)
            lin_ok &= util.relmax(sh.hessian_diagonal(), solo.hessian_diagonal()) <= 1e-12
            for lam, diag in ((0.25, False), (1e-2, True)):
                st, a0, a1, _ = sh.solve(lam, diag)
                so, b0, b1, _ = solo.solve(lam, diag)
                lin_ok &= st == so == 0 and abs(a0 - b0) <= 1e-12 * max(1.0, b0) and abs(a1 - b1) <= 1e-9 * max(1.0, b0)
                d_sh, d_solo = sh.get_delta(), solo.get_delta()
                mine = d_sh != 0
                lin_ok &= bool(mine.any()) and np.linalg.norm(d_sh[mine] - d_solo[mine]) <= 1e-7 * np.linalg.norm(d_solo[mine])
            sh.close(); solo.close()
    except Exception as e:     # noqa: BLE001
        lin_ok = False
        print("rank", rank, "linear sharded check raised:", repr(e))
    lflag = torch.tensor([int(bool(lin_ok))], device="cuda")
    dist.all_reduce(lflag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_LINEAR_OK" if lflag.item() == 1 else "MULTI_GPU_LINEAR_MISMATCH", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
