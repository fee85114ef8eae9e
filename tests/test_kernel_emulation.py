"""The library's GaussianFactorGraph-level kernels executed ON THE HOST, verbatim.

tests/test_gpu_linear.py is the real check (B200).  Until it has run, this test extracts the source text of
jacobian_load_kernel, assemble_{jacobian,hessian}_kernel, hdiag_{jacobian,hessian}_kernel and
linerr_{jacobian,hessian}_kernel from gtsam_b200/csrc/kernels.cuh (and the two view structs from engine.cuh), compiles
it with g++ behind a minimal emulation of the CUDA execution model (tests/emu/cuda_emu_prelude.h: sequential blocks and
threads, block reductions modelled, nothing else) and runs it on the linear fixtures with the tables the product's own
host symbolic phase produces: whitening, the assembled fronts (summed back into the global augmented Hessian), the
Hessian diagonal and both linear errors must equal the oracle's / numpy's.  It checks the kernels' indexing and
arithmetic, not their behaviour on the GPU."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import util
from gtsam_b200 import capi
from oracle import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ["jacobian_load_kernel", "assemble_jacobian_kernel", "hdiag_jacobian_kernel", "linerr_jacobian_kernel",
           "assemble_hessian_kernel", "hdiag_hessian_kernel", "linerr_hessian_kernel"]

WRAPPERS = r"""
extern "C" {
void emu_load(const double* Ab, const double* sigmas, int rows, int ncols, int count, double* J) {
  EMU_LAUNCH(jacobian_load_kernel, 7, 256, Ab, sigmas, rows, ncols, count, J);
}
static JacobianView mkview(int count, int rows, int arity, int ncols, const int* col0, const int* keys, const int* slots,
                           const int* clique, double* J) {
  JacobianView v;
  v.count = count; v.rows = rows; v.arity = arity; v.ncols = ncols;
  for (int a = 0; a < B200_JACOBIAN_MAX_ARITY + 2; a++) v.col0[a] = col0[a];
  v.keys = keys; v.slots = slots; v.clique = clique; v.J = J;
  return v;
}
static TreeView mktree(double* arena, const int64_t* off, const int* nf, const int* ns) {
  TreeView t;
  memset(&t, 0, sizeof t);
  t.arena = arena; t.off = off; t.nf = nf; t.ns = ns;
  return t;
}
void emu_assemble(int hessian, int count, int rows, int arity, int ncols, const int* col0, const int* keys, const int* slots,
                  const int* clique, double* J, double* arena, const int64_t* off, const int* nf, const int* ns) {
  JacobianView g = mkview(count, rows, arity, ncols, col0, keys, slots, clique, J);
  TreeView t = mktree(arena, off, nf, ns);
  const int nb = (count + 127) / 128;
  if (hessian) EMU_LAUNCH(assemble_hessian_kernel, nb, 128, g, t);
  else EMU_LAUNCH(assemble_jacobian_kernel, nb, 128, g, t);
}
void emu_hdiag(int hessian, int count, int rows, int arity, int ncols, const int* col0, const int* keys, const int* slots,
               const int* clique, double* J, const int* var_dof, double* hdiag) {
  JacobianView g = mkview(count, rows, arity, ncols, col0, keys, slots, clique, J);
  const int nb = (count + 127) / 128;
  if (hessian) EMU_LAUNCH(hdiag_hessian_kernel, nb, 128, g, var_dof, hdiag);
  else EMU_LAUNCH(hdiag_jacobian_kernel, nb, 128, g, var_dof, hdiag);
}
void emu_linerr(int hessian, int count, int rows, int arity, int ncols, const int* col0, const int* keys, const int* slots,
                const int* clique, double* J, const double* delta, const int* var_dof, double* out0, double* out1, int accumulate) {
  JacobianView g = mkview(count, rows, arity, ncols, col0, keys, slots, clique, J);
  double p0[64], p1[64];
  unsigned counters[2] = {0, 0};
  if (hessian) EMU_LAUNCH(linerr_hessian_kernel, 3, 256, g, delta, var_dof, p0, p1, counters, out0, out1, accumulate, 1.0);
  else EMU_LAUNCH(linerr_jacobian_kernel, 3, 256, g, delta, var_dof, p0, p1, counters, out0, out1, accumulate, 1.0);
}
}
"""


def _extract(text, start_pat):
    """source text from the line matching start_pat to the first line that is exactly '}' or '};'"""
    m = re.search(start_pat, text, re.M)
    assert m, start_pat
    end = re.compile(r"^\};?\s*$", re.M).search(text, m.start())
    return text[m.start():end.end()] + "\n"


@pytest.fixture(scope="module")
def emu():
    kern = open(os.path.join(ROOT, "gtsam_b200", "csrc", "kernels.cuh")).read()
    eng = open(os.path.join(ROOT, "gtsam_b200", "csrc", "engine.cuh")).read()
    src = '#include "cuda_emu_prelude.h"\n'
    src += _extract(eng, r"^struct JacobianView \{") + _extract(eng, r"^struct TreeView \{")
    for k in KERNELS:
        src += _extract(kern, r"^__global__ void __launch_bounds__\(\d+\) " + k + r"\(")
    src += WRAPPERS
    td = tempfile.mkdtemp()
    cpp, so = os.path.join(td, "emu.cpp"), os.path.join(td, "libemu.so")
    open(cpp, "w").write(src)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-fPIC", "-shared", "-I", os.path.join(ROOT, "tests", "emu"), cpp, "-o", so])
    return C.CDLL(so)


def _ip32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("case", [c for c in util.LINEAR_CASES if c != "lin_singular"])
def test_linear_kernels_emulated_on_host(emu, case, built):
    lp = util.load_linear_case(case)
    fp, fv, sp, sv, par, fclique, fptr, fslots = capi.linear_symbolic(lp, with_slots=True)
    dims = lp.var_dims.astype(np.int64)
    dof = lp.dof_offsets()
    nc = len(par)
    nf = np.array([dims[fv[fp[c]:fp[c + 1]]].sum() for c in range(nc)], dtype=np.int32)
    ns = np.array([dims[sv[sp[c]:sp[c + 1]]].sum() for c in range(nc)], dtype=np.int32)
    nn = (nf + ns + 1).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(nn * nn)]).astype(np.int64)      # the engine's layout without fused leaves
    arena = np.zeros(int(off[-1]))
    var_dof = dof.astype(np.int32)
    ndelta = int(dof[-1])
    hdiag = np.zeros(ndelta)
    orc = O.OracleLinearProblem(lp)
    assert orc.solve(0.25)[0] == 0          # damped: solvable for every fixture; delta only feeds the error check
    delta = orc.get_delta()
    e0, e1 = C.c_double(0), C.c_double(0)
    Hglob = np.zeros((ndelta + 1, ndelta + 1))   # numpy's own sum of [A b]^T [A b] / info over all factors
    first = True
    keep = []
    for g, hess in [(g, False) for g in lp.groups] + [(g, True) for g in lp.hgroups]:
        pos = g.graph_index if g.graph_index is not None else g.graph_index0 + np.arange(g.count)
        rows, ncols, ar = (g.ncols, g.ncols, g.arity) if hess else (g.rows, g.ncols, g.arity)
        src = np.ascontiguousarray(g.info if hess else g.Ab).ravel()
        sig = None if hess or g.sigmas is None else np.ascontiguousarray(g.sigmas).ravel()
        J = np.zeros(g.count * rows * ncols)
        emu.emu_load(_dp(src), _dp(sig) if sig is not None else None, rows, ncols, g.count, _dp(J))
        W = J.reshape(rows * ncols, g.count).T.reshape(g.count, ncols, rows).transpose(0, 2, 1)     # (count, rows, ncols)
        ref = (g.info.transpose(0, 2, 1) if hess else g.whitened())
        assert np.abs(W - ref).max() <= 1e-15 * max(1.0, np.abs(ref).max())
        col0 = np.zeros(10, dtype=np.int32)
        col0[1:ar + 1] = np.cumsum(g.dims)
        col0[ar + 1] = col0[ar] + 1
        keys32 = g.keys.astype(np.int32).ravel()
        slots = np.concatenate([fslots[fptr[p]:fptr[p + 1]] for p in pos]).astype(np.int32)
        clique = fclique[pos].astype(np.int32)
        keep += [J, col0, keys32, slots, clique]
        args = (int(hess), g.count, rows, ar, ncols, _ip32(col0), _ip32(keys32), _ip32(slots), _ip32(clique), _dp(J))
        emu.emu_assemble(*args, _dp(arena), off.ctypes.data_as(C.POINTER(C.c_int64)), _ip32(nf), _ip32(ns))
        emu.emu_hdiag(*args, _ip32(var_dof), _dp(hdiag))
        emu.emu_linerr(*args, _dp(delta), _ip32(var_dof), C.byref(e0), C.byref(e1), 0 if first else 1)
        first = False
        for i in range(g.count):   # numpy reference of the global augmented Hessian
            idx = np.concatenate([np.arange(dof[k], dof[k + 1]) for k in g.keys[i]] + [[ndelta]])
            M = np.triu(W[i]) + np.triu(W[i], 1).T if hess else W[i].T @ W[i]
            Hglob[np.ix_(idx, idx)] += M
    # fronts -> global: every front entry (i <= j) lands on (didx[i], didx[j])
    Hfront = np.zeros_like(Hglob)
    for c in range(nc):
        vars_c = list(fv[fp[c]:fp[c + 1]]) + list(sv[sp[c]:sp[c + 1]])
        didx = np.concatenate([np.arange(dof[v], dof[v + 1]) for v in vars_c] + [[ndelta]])
        F = arena[off[c]:off[c + 1]].reshape(nn[c], nn[c]).T          # column-major -> [row, col]
        assert np.abs(np.tril(F, -1)).max(initial=0.0) == 0.0         # only the upper triangle is written
        Fs = np.triu(F) + np.triu(F, 1).T
        Hfront[np.ix_(didx, didx)] += Fs
    scale = np.abs(Hglob).max()
    assert np.abs(Hfront - Hglob).max() <= 1e-12 * scale
    assert util.relmax(hdiag, orc.hessian_diagonal()) <= 1e-13
    assert util.relmax(hdiag, np.diag(Hglob)[:ndelta]) <= 1e-13
    x = np.concatenate([delta, [-1.0]])
    assert abs(e0.value - 0.5 * Hglob[ndelta, ndelta]) <= 1e-12 * abs(e0.value)
    assert abs(e1.value - 0.5 * x @ Hglob @ x) <= 1e-10 * abs(e0.value)


# ---- the marginal kernels (block-cooperative: real __syncthreads, no shuffles) in lockstep emulation ----------------
MARG_WRAPPERS = r"""
extern "C" {
static TreeView mktree2(double* arena, const int64_t* off, const int* nf, const int* ns, const int* ld, const int64_t* didx_ptr, const int* didx) {
  TreeView t;
  memset(&t, 0, sizeof t);
  t.arena = arena; t.off = off; t.nf = nf; t.ns = ns; t.ld = ld; t.didx_ptr = didx_ptr; t.didx = didx;
  return t;
}
void emu_marginal_path(double* arena, const int64_t* off, const int* nf, const int* ns, const int* ld, const int64_t* didx_ptr, const int* didx,
                       const int* path, int npath, int dof0, int d, double* work, int64_t ndelta, double* out) {
  TreeView t = mktree2(arena, off, nf, ns, ld, didx_ptr, didx);
  EMU_LAUNCH_LOCKSTEP(marginal_path_kernel, d, 256, t, path, npath, dof0, d, work, ndelta, out);
}
void emu_marginal_joint(double* arena, const int64_t* off, const int* nf, const int* ns, const int* ld, const int64_t* didx_ptr, const int* didx,
                        const int* path, int npath, const int* dofs, int D, double* work, int64_t ndelta, double* out) {
  TreeView t = mktree2(arena, off, nf, ns, ld, didx_ptr, didx);
  EMU_LAUNCH_LOCKSTEP(marginal_joint_kernel, D, 256, t, path, npath, dofs, D, work, ndelta, out);
}
}
"""


@pytest.fixture(scope="module")
def emu_marg():
    kern = open(os.path.join(ROOT, "gtsam_b200", "csrc", "kernels.cuh")).read()
    eng = open(os.path.join(ROOT, "gtsam_b200", "csrc", "engine.cuh")).read()
    src = '#include "cuda_emu_lockstep.h"\n' + _extract(eng, r"^struct TreeView \{")
    src += "constexpr int kMargMaxF = 4096;\n"
    for k in ("marginal_path_kernel", "marginal_joint_kernel"):
        src += _extract(kern, r"^__global__ void __launch_bounds__\(256\)\n" + k + r"\(")
    src += MARG_WRAPPERS
    td = tempfile.mkdtemp()
    cpp, so = os.path.join(td, "emu_marg.cpp"), os.path.join(td, "libemu_marg.so")
    open(cpp, "w").write(src)
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-w", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "tests", "emu"), cpp, "-o", so])
    return C.CDLL(so)


@pytest.mark.parametrize("case", ["bal_tiny_s2", "sphere_tiny"])
def test_marginal_kernels_emulated_on_host(emu_marg, case):
    """marginal_path_kernel (validated on the B200) and marginal_joint_kernel (first hardware run pending), verbatim, in
    lockstep emulation on the oracle's conditionals [R S d]: single-variable covariances and the joint covariances of
    the golden variable sets against the unmodified reference's Marginals."""
    prob = util.load_case(case)
    orc = O.OracleProblem(prob)
    orc.linearize()
    assert orc.solve(0.0)[0] == 0
    fp, fv, sp, sv, par = orc.cliques()
    dims = prob.var_dims.astype(np.int64)
    dof = prob.dof_offsets()
    nc, ndelta = len(par), int(dof[-1])
    nf = np.array([dims[fv[fp[c]:fp[c + 1]]].sum() for c in range(nc)], dtype=np.int32)
    ns = np.array([dims[sv[sp[c]:sp[c + 1]]].sum() for c in range(nc)], dtype=np.int32)
    ld = nf.copy()                                  # compact conditional storage, as for fused leaves: f x (f+s+1), ld = f
    off = np.concatenate([[0], np.cumsum(nf.astype(np.int64) * (nf + ns + 1))]).astype(np.int64)
    arena = np.concatenate([orc.conditional(c).T.ravel() for c in range(nc)])     # (f, n) -> column-major
    didx_ptr = np.concatenate([[0], np.cumsum(nf + ns)]).astype(np.int64)
    didx = np.concatenate([np.concatenate([np.arange(dof[v], dof[v + 1]) for v in list(fv[fp[c]:fp[c + 1]]) + list(sv[sp[c]:sp[c + 1]])])
                           for c in range(nc)]).astype(np.int32)
    var_clique = np.zeros(prob.nvars, dtype=np.int64)
    for c in range(nc):
        var_clique[fv[fp[c]:fp[c + 1]]] = c
    i64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))   # noqa: E731
    tree = (_dp(arena), i64(off), _ip32(nf), _ip32(ns), _ip32(ld), i64(didx_ptr), _ip32(didx))
    # single variables vs the reference's Marginals::marginalCovariance
    ref = util.golden(case, "marg")["marg_cov"]
    o = 0
    for v in range(prob.nvars):
        d = int(dims[v])
        R = ref[o:o + d * d].reshape(d, d).T
        o += d * d
        if v not in (0, prob.nvars // 2, prob.nvars - 1):   # three variables keep the 256-thread emulation quick
            continue
        path = []
        c = int(var_clique[v])
        while c >= 0:
            path.append(c); c = int(par[c])
        path = np.array(path, dtype=np.int32)
        work, out = np.zeros(d * ndelta), np.zeros(d * d)
        emu_marg.emu_marginal_path(*tree, _ip32(path), len(path), int(dof[v]), d, _dp(work), C.c_int64(ndelta), _dp(out))
        S = out.reshape(d, d).T
        assert np.abs(S - R).max() <= 1e-8 * np.abs(R).max(), v
    # joint sets vs Marginals::jointMarginalCovariance (tests/golden/<case>.joint<i>.bin)
    for k, vs in enumerate(util.JOINT_SETS[case]):
        vs = sorted(vs)
        on = np.zeros(nc, dtype=bool)
        for v in vs:
            c = int(var_clique[v])
            while c >= 0 and not on[c]:
                on[c] = True; c = int(par[c])
        path = np.nonzero(on)[0].astype(np.int32)
        dofs = np.concatenate([np.arange(dof[v], dof[v + 1]) for v in vs]).astype(np.int32)
        D = len(dofs)
        work, out = np.zeros(D * ndelta), np.zeros(D * D)
        emu_marg.emu_marginal_joint(*tree, _ip32(path), len(path), _ip32(dofs), D, _dp(work), C.c_int64(ndelta), _dp(out))
        S = out.reshape(D, D).T
        g = util.golden(case, f"joint{k}")
        R = next(iter(g.values())) if len(g) == 1 else g.get("joint_cov", g.get("cov"))
        R = np.asarray(R).reshape(D, D).T
        assert np.abs(S - R).max() <= 1e-8 * np.abs(R).max(), vs
