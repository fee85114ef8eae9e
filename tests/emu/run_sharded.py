"""One RANK of the sharded solve in host emulation: `python run_sharded.py <emu lib> <rank> <world> <hex unique id>`.
The sharded problem (a communicator of `world` emulation processes over tests/emu/fake_nccl.cpp) against the same problem
solved alone in this process — the checks of tests/multi_gpu_check.py on small problems.  TEST INFRASTRUCTURE."""
import os
import sys

os.environ["B200_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gtsam_b200 import capi  # noqa: E402

capi.LIB_PATH = sys.argv[1]
rank, world, uid = int(sys.argv[2]), int(sys.argv[3]), bytes.fromhex(sys.argv[4])
import numpy as np  # noqa: E402
import util  # noqa: E402
from gtsam_b200 import datasets, optimizer  # noqa: E402

ctx = capi.Context(0)
ctx.comm_init(uid, rank, world)
solo_ctx = capi.Context(0)
ok = True
cases = [datasets.make("bal_tiny", ncams=12, npoints=300, visibility="scattered"),
         datasets.make("bal_tiny", ncams=10, npoints=120, visibility="banded", camera_model="bundler"),
         util.load_case("sphere_small_colamd"), util.load_case("bal_small_metis"), util.load_case("pose2_ring_colamd"),
         datasets.make("sphere_tiny", layers=6, per_ring=10)]
for prob in cases:
    co, fo = capi.shard_plan(prob, world)
    sh, solo = capi.DeviceProblem(ctx, prob), capi.DeviceProblem(solo_ctx, prob)
    e_sh, e_solo = sh.error(), solo.error()
    ok &= abs(e_sh - e_solo) <= 1e-12 * e_solo
    sh.linearize(); solo.linearize()
    ok &= util.relmax(sh.gradient_at_zero(), solo.gradient_at_zero()) <= 1e-12
    for lam, diag in ((1e-3, False), (1e-2, True)):
        st, a0, a1, _ = sh.solve(lam, diag)
        so, b0, b1, _ = solo.solve(lam, diag)
        ok &= st == so == 0 and abs(a0 - b0) <= 1e-12 * b0 and abs(a1 - b1) <= 1e-9 * b0
        d_sh, d_solo = sh.get_delta(), solo.get_delta()
        fp, fv, sp, sv, par = solo.cliques()
        dof = prob.dof_offsets()
        mine = np.zeros(d_solo.size, dtype=bool)
        for c in range(len(par)):
            if co[c] in (-1, rank):
                for v in fv[fp[c]:fp[c + 1]]:
                    mine[dof[v]:dof[v + 1]] = True
        ok &= np.linalg.norm(d_sh[mine] - d_solo[mine]) <= 1e-7 * max(1e-300, np.linalg.norm(d_solo[mine]))
        ok &= bool(np.all(d_sh[~mine] == 0))
        ok &= abs(sh.try_step() - solo.try_step()) <= 1e-9 * e_solo
    lm_sh = optimizer.LevenbergMarquardtOptimizer(ctx, prob, device_problem=sh)
    lm_solo = optimizer.LevenbergMarquardtOptimizer(solo_ctx, prob, device_problem=solo)
    for _ in range(3):
        lm_sh.iterate(); lm_solo.iterate()
        ok &= abs(lm_sh.error() - lm_solo.error()) <= 1e-8 * lm_solo.error()
        ok &= lm_sh.lambda_() == lm_solo.lambda_() and lm_sh.getInnerIterations() == lm_solo.getInnerIterations()
    del lm_sh, lm_solo
    # values views (b200_values_view): only the input view is uploaded — everything else on the device is poisoned —
    # and after an LM iteration the owned view equals the single-GPU values of those variables
    idx_in, idx_own = sh.view_index(0), sh.view_index(1)
    sh.set_values(np.full(prob.values.size, np.nan))
    sh.set_values_view(np.ascontiguousarray(prob.values[idx_in]))
    solo.set_values(prob.values)
    lm_sh = optimizer.LevenbergMarquardtOptimizer(ctx, prob, device_problem=sh)
    lm_solo = optimizer.LevenbergMarquardtOptimizer(solo_ctx, prob, device_problem=solo)
    capi._check(sh.L.b200_lm_reset(lm_sh.h)); capi._check(solo.L.b200_lm_reset(lm_solo.h))
    lm_sh.iterate(); lm_solo.iterate()
    ok &= abs(lm_sh.error() - lm_solo.error()) <= 1e-8 * lm_solo.error()
    mine_v = sh.get_values_view(np.empty(idx_own.size))
    ok &= bool(np.allclose(mine_v, solo.get_values()[idx_own], rtol=1e-6, atol=1e-6))   # (delta agrees to ~1e-7 relative, see above)
    del lm_sh, lm_solo
    sh.close(); solo.close()
    print("rank", rank, prob.name, "ok" if ok else "MISMATCH", "owned cliques", int((co == rank).sum()), "top", int((co == -1).sum()), flush=True)
# the GaussianFactorGraph level, sharded the same way: JacobianFactor / HessianFactor groups split by owning subtree
for name in ("lin_sphere_tiny", "lin_bal_tiny", "lin_random_nary", "lin_mixed_hessian", "lin_arity8"):
    lp = util.load_linear_case(name)
    sh, solo = capi.LinearDeviceProblem(ctx, lp), capi.LinearDeviceProblem(solo_ctx, lp)
    rng = np.random.default_rng(11)
    for rnd in range(2):
        if rnd == 1:    # new numbers, same structure: every rank is handed the whole group and stages its own share
            for gi, g in enumerate(lp.groups):
                Ab = np.asarray(g.Ab, dtype=np.float64) * (1.0 + 0.05 * rng.standard_normal(np.asarray(g.Ab).shape))
                for d in (sh, solo):
                    d.update(gi, Ab, g.sigmas)
            for hi, g in enumerate(lp.hgroups):
                info = np.asarray(g.info, dtype=np.float64).copy()
                info.reshape(g.count, g.ncols, g.ncols)[:, -1, :] *= 1.05     # rhs row / column and the constant
                info.reshape(g.count, g.ncols, g.ncols)[:, :-1, -1] *= 1.05
                for d in (sh, solo):
                    d.update_hessian(hi, info)
        ok &= util.relmax(sh.hessian_diagonal(), solo.hessian_diagonal()) <= 1e-12
        ok &= util.relmax(sh.gradient_at_zero(), solo.gradient_at_zero()) <= 1e-12
        xr = rng.standard_normal(sh.ndelta)
        ok &= abs(sh.linear_graph_error(xr) - solo.linear_graph_error(xr)) <= 1e-12 * solo.linear_graph_error(xr)
        for lam, diag in ((0.25, False), (1e-2, True)):
            st, a0, a1, _ = sh.solve(lam, diag)
            so, b0, b1, _ = solo.solve(lam, diag)
            ok &= st == so == 0 and abs(a0 - b0) <= 1e-12 * max(1.0, b0) and abs(a1 - b1) <= 1e-9 * max(1.0, b0)
            d_sh, d_solo = sh.get_delta(), solo.get_delta()
            mine = d_sh != 0      # delta of the variables this rank owns or shares (the rest stays zero)
            ok &= bool(mine.any()) and np.linalg.norm(d_sh[mine] - d_solo[mine]) <= 1e-7 * np.linalg.norm(d_solo[mine])
    sh.close(); solo.close()
    print("rank", rank, name, "ok" if ok else "MISMATCH", "delta entries here", int(mine.sum()), "of", mine.size, flush=True)
print("SHARDED_OK" if ok else "SHARDED_MISMATCH", rank, world, flush=True)
sys.exit(0 if ok else 1)
