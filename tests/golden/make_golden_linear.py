"""Regenerates tests/golden/lin_*.{lin,out0,out1}.bin: GaussianFactorGraph-level fixtures produced by the
UNMODIFIED reference (oracle/_ref/ref_harness linsolve / linearize2d).  Run from the repo root in the build
container:

    python tests/golden/make_golden_linear.py

* lin_pose2_toy   — BASELINE.json configs[0]: examples/Data/noisyToyGraph.txt (Pose2 g2o + the example's prior)
                    linearized BY THE REFERENCE at the file's initial estimate, COLAMD ordering.
* lin_pose2_synth — the same for tests/golden/data/synthetic_pose2.g2o (40 poses on a ring, odometry + loop closures,
                    written by synthetic_pose2() below); pose2_synth_reference.json records the reference's
                    Pose2SLAMExample_g2o path (GN as shipped, LM as configs[0] names it) on that file.
* lin_random_nary — random JacobianFactors of arity 1..4, rows 1..7, block widths 1/2/3/6, half of the groups with a
                    Diagonal model (sigmas), random elimination order.
* lin_arity8      — a few factors of the maximum supported arity (8).
* lin_sphere_tiny, lin_bal_tiny — the reference's own linearization of the sphere_tiny / bal_tiny_s2 graphs (the
                    whitened [A|b] of tests/golden/<case>.dump0.bin) re-posed as linear problems: must reproduce
                    those dumps' delta.
* lin_mixed_hessian — lin_random_nary with every third factor turned into a HessianFactor (augmented information
                    [A b]^T Sigma^-1 [A b] of the same numbers, plus a few genuinely full-rank quadratic factors): Jacobian
                    and Hessian factors interleaved in one graph.
* lin_family_sfm2 / _smart / _expr — the reference's linearizations of GeneralSFMFactor2 (ternary factors, a 5-dof
                    calibration variable), SmartProjectionPoseFactor (HessianFactors over 6 cameras) and ExpressionFactor
                    graphs of one small synthetic scene (tests/shim_families.cpp dumplin).
* lin_singular    — an under-constrained graph: the reference throws IndeterminantLinearSystemException.
Each case stores the problem (*.lin.bin, gtsam_b200.linear.LinearProblem.save) and the reference's outputs for
lambda = 0 (*.out0.bin: delta, hessianDiagonal, linear errors, Bayes tree + conditionals, marginal covariances) and
lambda = 0.25 (*.out1.bin: the damped system of LevenbergMarquardtState.h:125-156).
"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtsam_b200 import linear as LN, problem as P  # noqa: E402
from oracle import refio  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
H = refio.HARNESS
REF_DATA = "/root/reference/examples/Data"


def emit(name, lp=None):
    path = os.path.join(HERE, f"{name}.lin.bin")
    if lp is not None:
        lp.save(path)
    subprocess.check_call([H, "linsolve", path, os.path.join(HERE, f"{name}.out0.bin"), "0"])
    subprocess.check_call([H, "linsolve", path, os.path.join(HERE, f"{name}.out1.bin"), "0.25"])
    lp = LN.LinearProblem.load(path)
    print("wrote", name, lp.nvars, "vars", lp.nfactors, "factors", len(lp.groups), "+", len(lp.hgroups), "groups")


def random_nary(seed=0, nvars=40, nfac=90, max_arity=4):
    rng = np.random.default_rng(seed)
    var_dim = rng.choice([1, 2, 3, 6], size=nvars).astype(np.int32)
    factors = []   # (keys, rows, has_sigma)
    for v in range(nvars):   # a prior on every variable keeps the system well conditioned
        factors.append(([v], int(var_dim[v]), bool(v % 2)))
    for _ in range(nfac):
        ar = int(rng.integers(1, max_arity + 1))
        v0 = int(rng.integers(0, nvars))
        near = (v0 + rng.choice(np.arange(1, 9), size=ar - 1, replace=False)) % nvars   # locality => a real tree
        factors.append(([v0] + [int(x) for x in near], int(rng.integers(1, 8)), bool(rng.integers(0, 2))))
    order = rng.permutation(len(factors))
    buckets = {}
    for pos, fi in enumerate(order):
        keys, rows, sig = factors[fi]
        dims = tuple(int(var_dim[k]) for k in keys)
        b = buckets.setdefault((rows, sig) + dims, dict(keys=[], Ab=[], sig=[], pos=[]))
        nc = sum(dims) + 1
        A = rng.normal(size=(nc, rows))
        if len(keys) == 1 and rows == dims[0]:
            A[:rows, :] += 3.0 * np.eye(rows)
        b["keys"].append(keys)
        b["Ab"].append(A)
        b["sig"].append(rng.uniform(0.2, 3.0, size=rows))
        b["pos"].append(pos)
    groups = [LN.JacobianGroup(sig[0], sig[2:], np.array(b["keys"]), np.array(b["Ab"]),
                               np.array(b["sig"]) if sig[1] else None, graph_index=np.array(b["pos"]))
              for sig, b in buckets.items()]
    return LN.LinearProblem(var_dim, rng.permutation(nvars), groups)


def arity8(seed=3):
    rng = np.random.default_rng(seed)
    nv = 12
    var_dim = np.array([2, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3, 1], dtype=np.int32)
    pri = [LN.JacobianGroup(int(d), [int(d)], np.array([[v] for v in range(nv) if var_dim[v] == d]),
                            np.array([np.concatenate([2 * np.eye(d), rng.normal(size=(1, d))], 0) for v in range(nv) if var_dim[v] == d]))
           for d in (1, 2, 3)]
    keys = np.array([[0, 1, 2, 3, 4, 5, 6, 7], [3, 4, 5, 6, 7, 8, 9, 10]])
    dims = [int(var_dim[k]) for k in keys[0]]
    assert dims == [int(var_dim[k]) for k in keys[1]]
    big = LN.JacobianGroup(5, dims, keys, rng.normal(size=(2, sum(dims) + 1, 5)), rng.uniform(0.5, 2.0, size=(2, 5)))
    return LN.LinearProblem(var_dim, rng.permutation(nv), pri + [big])


def from_typed_dump(case):
    """The reference's linearization of a typed problem (whitened [A|b] of <case>.dump0.bin) as a linear problem."""
    prob = P.Problem.load(os.path.join(HERE, f"{case}.prob.bin"))
    ref = refio.read_out(os.path.join(HERE, f"{case}.dump0.bin"))
    vd = prob.var_dims
    groups = []
    for gi, g in enumerate(prob.groups):
        d, nc = P.FACTOR_DIM[g.type], P.factor_ncols(g.type)
        dims = [int(vd[k]) for k in g.keys[0]]
        groups.append(LN.JacobianGroup(d, dims, g.keys, ref[f"J{gi}"].reshape(g.count, nc, d), None,
                                       graph_index0=g.graph_index0, graph_index=g.graph_index))
    return LN.LinearProblem(vd.astype(np.int32), prob.ordering, groups)


def synthetic_pose2(path, n=40, seed=21):
    """A Pose2 g2o file in the format of examples/Data/noisyToyGraph.txt (VERTEX_SE2 / EDGE_SE2 with the upper
    triangle of the information matrix): n poses on a ring expressed in the frame of pose 0 (the example pins pose 0
    at the origin), odometry + two families of loop closures."""
    rng = np.random.default_rng(seed)
    th = 2 * np.pi * np.arange(n) / n
    ring = np.stack([10 * np.cos(th), 10 * np.sin(th), th + np.pi / 2], -1)

    def between(a, b):
        c, s = np.cos(a[2]), np.sin(a[2])
        d = b[:2] - a[:2]
        return np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], (b[2] - a[2] + np.pi) % (2 * np.pi) - np.pi])

    gt = np.array([between(ring[0], p) for p in ring])
    init = gt + rng.normal(size=gt.shape) * np.array([0.3, 0.3, 0.05])
    init[0] = 0
    edges = [(i, (i + 1) % n) for i in range(n)] + [(i, (i + 7) % n) for i in range(0, n, 3)] + [(i, (i + 19) % n) for i in range(1, n, 5)]
    lines = ["VERTEX_SE2 %d %.9f %.9f %.9f" % (i, *init[i]) for i in range(n)]
    for (i, j) in edges:
        z = between(gt[i], gt[j]) + rng.normal(size=3) * np.array([0.05, 0.05, 0.01])
        lines.append("EDGE_SE2 %d %d %.9f %.9f %.9f 400 0 0 400 0 10000" % (i, j, *z))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def mixed_hessian(seed=4):
    base = random_nary(seed=seed, nvars=30, nfac=70, max_arity=3)
    rng = np.random.default_rng(seed + 100)
    jg, hb = [], {}
    for g in base.groups:
        pos = g.graph_index
        W = g.whitened()                      # (count, rows, ncols)
        keepj = [i for i in range(g.count) if pos[i] % 3 != 0]
        toh = [i for i in range(g.count) if pos[i] % 3 == 0]
        if keepj:
            jg.append(LN.JacobianGroup(g.rows, g.dims, g.keys[keepj], g.Ab[keepj], None if g.sigmas is None else g.sigmas[keepj],
                                       graph_index=pos[keepj]))
        for i in toh:
            info = W[i].T @ W[i]
            if i % 2 == 0:                     # a genuinely full-rank quadratic, not just a Jacobian in disguise
                M = rng.normal(size=(info.shape[0] + 2, info.shape[0])) * 0.3
                info = info + M.T @ M
            b = hb.setdefault(tuple(int(d) for d in g.dims), dict(keys=[], info=[], pos=[]))
            b["keys"].append(g.keys[i]); b["info"].append(info); b["pos"].append(pos[i])
    hg = [LN.HessianGroup(dims, np.array(b["keys"]), np.array(b["info"]), graph_index=np.array(b["pos"])) for dims, b in hb.items()]
    return LN.LinearProblem(base.var_dim, base.ordering, jg, hg)


def singular():
    # x0 -- x1 -- x2 chain with no prior anywhere: A^T A is rank deficient
    rng = np.random.default_rng(9)
    Ab = np.zeros((2, 5, 2))
    for f in range(2):
        Ab[f, 0:2, :] = np.eye(2)
        Ab[f, 2:4, :] = -np.eye(2)
        Ab[f, 4, :] = rng.normal(size=2)
    return LN.LinearProblem(np.array([2, 2, 2], dtype=np.int32), np.array([0, 1, 2]),
                            [LN.JacobianGroup(2, [2, 2], np.array([[0, 1], [1, 2]]), Ab)])


def main():
    assert refio.have_ref(), "build oracle/_ref first: make -C oracle ref"
    subprocess.check_call([H, "linearize2d", os.path.join(REF_DATA, "noisyToyGraph.txt"), os.path.join(HERE, "lin_pose2_toy.lin.bin")])
    emit("lin_pose2_toy")
    g2o = os.path.join(HERE, "data", "synthetic_pose2.g2o")
    synthetic_pose2(g2o)
    subprocess.check_call([H, "linearize2d", g2o, os.path.join(HERE, "lin_pose2_synth.lin.bin")])
    emit("lin_pose2_synth")
    with open(os.path.join(HERE, "pose2_synth_reference.json"), "w") as f:
        f.write(subprocess.check_output([H, "pose2", g2o]).decode().strip().splitlines()[-1] + "\n")
    emit("lin_random_nary", random_nary())
    emit("lin_arity8", arity8())
    emit("lin_mixed_hessian", mixed_hessian())
    emit("lin_sphere_tiny", from_typed_dump("sphere_tiny"))
    emit("lin_bal_tiny", from_typed_dump("bal_tiny_s2"))
    emit("lin_singular", singular())
    # SURVEY 8(f) rank-4 families (GeneralSFMFactor2, smart factors in HESSIAN mode, expression factors) linearized by the
    # reference on a small synthetic scene (tests/shim_families.cpp)
    fam = os.path.join(os.path.dirname(H), "shim_families")
    if os.path.exists(fam):
        subprocess.check_call([fam, "dumplin", HERE])
        for name in ("sfm2", "smart", "expr"):
            emit(f"lin_family_{name}")


if __name__ == "__main__":
    main()
