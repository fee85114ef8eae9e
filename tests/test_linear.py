"""GaussianFactorGraph level (b200_linear_create, SURVEY 8b "GaussianFactorGraph::optimize-level entry"), CPU side:
the oracle's restatement and the product's host code against the UNMODIFIED reference's own
GaussianFactorGraph::optimize(ordering, EliminatePreferCholesky) on the same JacobianFactors
(tests/golden/lin_*.bin, written by tests/golden/make_golden_linear.py through oracle/_ref/ref_harness linsolve)."""
import ctypes as C

import numpy as np
import pytest

import util
from gtsam_b200 import capi, linear as LN, problem as P
from oracle import oracle_py as O


@pytest.mark.parametrize("case", util.LINEAR_CASES)
@pytest.mark.parametrize("which", [0, 1])
def test_oracle_linear_matches_reference(case, which):
    lp = util.load_linear_case(case)
    ref = util.golden(case, f"out{which}")
    util.check_linear_against_reference(O.OracleLinearProblem(lp), lp, ref, util.LINEAR_LAMBDA[which])


def test_singular_graph_is_indeterminate_in_the_reference():
    assert util.golden("lin_singular", "out0")["status"][0] == 1     # IndeterminantLinearSystemException
    assert util.golden("lin_singular", "out1")["status"][0] == 0     # damping makes it solvable


@pytest.mark.parametrize("case", ["lin_sphere_tiny", "lin_bal_tiny"])
def test_linear_level_reproduces_the_typed_dump(case):
    """The reference's linearization of a nonlinear graph, re-posed as JacobianFactors, has the solution of the
    nonlinear-level dump: both levels of the API describe the same system."""
    typed = {"lin_sphere_tiny": "sphere_tiny", "lin_bal_tiny": "bal_tiny_s2"}[case]
    a, b = util.golden(case, "out0"), util.golden(typed, "dump0")
    assert util.rel2(a["delta"], b["delta"]) <= 1e-12
    assert util.ref_clique_set(a) == util.ref_clique_set(b)


@pytest.mark.parametrize("case", util.LINEAR_CASES)
def test_host_symbolic_nary_matches_reference(case, built):
    """The product's host symbolic phase on n-ary factors builds the reference's Bayes-tree cliques."""
    lp = util.load_linear_case(case)
    ref = util.golden(case, "out1")      # out1 exists (solvable) for every case
    fp, fv, sp, sv, par = capi.linear_symbolic(lp)
    assert util.clique_set(fp, fv, sp, sv) == util.ref_clique_set(ref)
    ofp, ofv, osp, osv, opar = O.OracleLinearProblem(lp).cliques()
    assert (fp == ofp).all() and (fv == ofv).all() and (sp == osp).all() and (sv == osv).all() and (par == opar).all()


@pytest.mark.parametrize("case", ["lin_pose2_toy", "lin_random_nary", "lin_arity8"])
def test_oracle_linear_marginals_match_reference(case):
    """Marginal covariance of every variable (inverse of GaussianBayesTree::marginalFactor's information)."""
    lp = util.load_linear_case(case)
    ref = util.golden(case, "out0")["marginal_covariances"]
    o = O.OracleLinearProblem(lp)
    off = 0
    for v in range(lp.nvars):
        d = int(lp.var_dim[v])
        st, S = o.marginal_covariance(v)
        R = ref[off:off + d * d].reshape(d, d).T
        off += d * d
        assert st == 0 and np.abs(S - R).max() <= 1e-9 * np.abs(R).max()


def test_oracle_linear_update_reuses_structure():
    lp = util.load_linear_case("lin_random_nary")
    o = O.OracleLinearProblem(lp)
    rng = np.random.default_rng(1)
    g = lp.groups[3]
    newAb = g.Ab + 0.01 * rng.normal(size=g.Ab.shape)
    o.update(3, newAb, g.sigmas)
    st, _, _, _ = o.solve(0.0)
    lp2 = LN.LinearProblem(lp.var_dim, lp.ordering, [LN.JacobianGroup(h.rows, h.dims, h.keys, newAb if i == 3 else h.Ab, h.sigmas,
                                                                       h.graph_index0, h.graph_index) for i, h in enumerate(lp.groups)])
    o2 = O.OracleLinearProblem(lp2)
    st2, _, _, _ = o2.solve(0.0)
    assert st == st2 == 0 and np.array_equal(o.get_delta(), o2.get_delta())


def test_oracle_hessian_update_reuses_structure():
    lp = util.load_linear_case("lin_mixed_hessian")
    o = O.OracleLinearProblem(lp)
    h = lp.hgroups[1]
    new = h.info * 1.01
    o.update_hessian(1, new)
    assert o.solve(0.0)[0] == 0
    lp2 = LN.LinearProblem(lp.var_dim, lp.ordering, lp.groups,
                           [LN.HessianGroup(g.dims, g.keys, new if i == 1 else g.info, g.graph_index0, g.graph_index) for i, g in enumerate(lp.hgroups)])
    o2 = O.OracleLinearProblem(lp2)
    assert o2.solve(0.0)[0] == 0 and np.array_equal(o.get_delta(), o2.get_delta())


def test_python_graph_mirror_packs_like_the_harness():
    """gtsam_b200.linear.GaussianFactorGraph.to_problem: ids in ascending key order, factors grouped by shape with
    their graph positions — the same system as the LinearProblem it came from (checked through the oracle)."""
    lp = util.load_linear_case("lin_mixed_hessian")
    gfg = LN.GaussianFactorGraph()
    flat = {}
    for g in lp.groups:
        pos = g.graph_index if g.graph_index is not None else g.graph_index0 + np.arange(g.count)
        for i in range(g.count):
            M, blocks, c = g.Ab[i].T, [], 0
            for d in g.dims:
                blocks.append(M[:, c:c + d]); c += d
            flat[int(pos[i])] = LN.JacobianFactor([1000 + 7 * int(k) for k in g.keys[i]], blocks, M[:, c], None if g.sigmas is None else g.sigmas[i])
    for g in lp.hgroups:
        for i in range(g.count):
            flat[int(g.graph_index[i])] = LN.HessianFactor([1000 + 7 * int(k) for k in g.keys[i]], g.dims, g.info[i])
    for pos in sorted(flat):
        gfg.add(flat[pos])
    order_keys = [1000 + 7 * int(v) for v in lp.ordering]
    lp2, ids = gfg.to_problem(order_keys)
    assert gfg.size() == lp.nfactors and sorted(ids) == gfg.keys()
    a, b = O.OracleLinearProblem(lp), O.OracleLinearProblem(lp2)
    assert a.solve(0.0)[0] == b.solve(0.0)[0] == 0
    assert np.array_equal(a.get_delta(), b.get_delta())


def test_linear_exports_and_no_gpu_fails_loudly(built):
    L = capi.lib()
    for sym in ("b200_linear_create", "b200_linear_update", "b200_linear_symbolic_create"):
        assert hasattr(L, sym)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.B200Error) as e:
        LN.GaussianFactorGraph([LN.JacobianFactor([0], [np.eye(2)], np.ones(2))]).optimize()
    assert e.value.code == P.NO_DEVICE


def test_linear_description_is_validated(built):
    L = capi.lib()
    lp = util.load_linear_case("lin_arity8")

    def status(mutate):
        q = LN.LinearProblem(lp.var_dim.copy(), lp.ordering.copy(),
                             [LN.JacobianGroup(g.rows, g.dims.copy(), g.keys.copy(), g.Ab.copy(), None if g.sigmas is None else g.sigmas.copy(),
                                               g.graph_index0, g.graph_index) for g in lp.groups])
        mutate(q)
        desc, keep = q.c_desc()
        h = C.c_void_p()
        rc = L.b200_linear_symbolic_create(C.byref(desc), C.byref(h))
        if rc == 0:
            L.b200_symbolic_destroy(h)
        return rc

    assert status(lambda q: None) == P.OK
    assert status(lambda q: q.groups[3].sigmas.__setitem__((0, 0), 0.0)) == P.UNSUPPORTED_NOISE     # Constrained row
    assert status(lambda q: q.groups[3].keys.__setitem__((0, 0), 99)) == P.INVALID_ARGUMENT          # key out of range
    assert status(lambda q: q.groups[3].keys.__setitem__((0, 1), int(q.groups[3].keys[0, 0]))) == P.INVALID_ARGUMENT   # dims / duplicate
    assert status(lambda q: q.var_dim.__setitem__(0, 5)) == P.INVALID_ARGUMENT                        # block width mismatch
    assert status(lambda q: q.ordering.__setitem__(0, int(q.ordering[1]))) == P.INVALID_ARGUMENT      # not a permutation


SHIM_LINEAR = __import__("os").path.join(__import__("os").path.dirname(util.GOLDEN), "..", "oracle", "_ref", "shim_linear")


@pytest.mark.skipif(not __import__("os").path.exists(SHIM_LINEAR), reason="shim_linear not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", util.LINEAR_CASES)
def test_cpp_shim_packs_a_gaussian_factor_graph_like_the_reference(case):
    """C++ drop-in, host side only: gtsam_b200::symbolicOnHost packs a real gtsam::GaussianFactorGraph exactly as
    optimizeOnDevice does (ids by ascending Key, groups by shape, explicit graph positions) and runs the library's
    symbolic phase; the cliques equal those of the reference's eliminateMultifrontal."""
    import json
    import subprocess
    out = subprocess.check_output([SHIM_LINEAR, "hostpack", __import__("os").path.join(util.GOLDEN, f"{case}.lin.bin")], timeout=120)
    r = json.loads(out.decode().strip().splitlines()[-1])
    assert r["equal"] == 1 and r["cliques"] == r["reference_cliques"] > 0


SHIM_FAMILIES = __import__("os").path.join(__import__("os").path.dirname(util.GOLDEN), "..", "oracle", "_ref", "shim_families")


@pytest.mark.skipif(not __import__("os").path.exists(SHIM_FAMILIES), reason="shim_families not built (needs /root/reference at build time)")
def test_cpp_shim_accepts_the_rank4_factor_families():
    """SURVEY 8(f) rank 4 through the solve() seam, host side: GeneralSFMFactor2 (ternary Jacobians), smart projection
    factors (HessianFactors over all cameras of a point) and expression factors linearized by GTSAM; the shim packs each
    GaussianFactorGraph and the library's symbolic phase builds the reference's cliques."""
    import json
    import subprocess
    r = json.loads(subprocess.check_output([SHIM_FAMILIES, "host"], timeout=120).decode().strip().splitlines()[-1])
    assert set(r) == {"sfm2", "smart", "expr"}
    for name, f in r.items():
        assert f["cliques_equal"] == 1 and f["other"] == 0 and f["cliques"] > 0, (name, f)
    assert r["sfm2"]["max_arity"] == 3 and r["smart"]["hessian"] == 24 and r["smart"]["max_arity"] == 6


@pytest.mark.skipif(not __import__("os").path.exists(SHIM_LINEAR), reason="shim_linear not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", ["lin_pose2_toy", "lin_random_nary", "lin_mixed_hessian", "lin_family_smart", "lin_bal_tiny"])
def test_cpp_shim_builds_a_real_bayes_tree_from_flat_tables(case):
    """gtsam_b200::bayesTreeFromTables (what eliminateMultifrontalOnDevice calls on the device's cliques + conditionals),
    fed with tables extracted from the reference's own GaussianBayesTree: same clique count, optimize() and
    log-determinant."""
    import json
    import subprocess
    out = subprocess.check_output([SHIM_LINEAR, "bayestree", __import__("os").path.join(util.GOLDEN, f"{case}.lin.bin")], timeout=120)
    r = json.loads(out.decode().strip().splitlines()[-1])
    assert r["cliques"] == r["reference_cliques"] > 0 and r["optimize_rel_diff"] <= 1e-14 and r["log_determinant_diff"] <= 1e-10
