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


# ---- the factor evaluators: linearize_kernel<TYPE, JT> and error_kernel<TYPE>, verbatim, with the product's own
# ---- factors.cuh / geometry.cuh compiled for the host ----------------------------------------------------------------
EVAL_WRAPPERS = r"""
#define EMU_TYPES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
extern "C" {
static GroupView mkgroup(int type, int count, int noise_kind, int per_factor, int noise_size, int robust_kind, double robust_param,
                         const int* keys, const double* meas, const double* noise, const int* cal_index, const double* body, void* J) {
  GroupView g;
  memset(&g, 0, sizeof g);
  g.type = type; g.noise_kind = noise_kind; g.per_factor = per_factor; g.noise_size = noise_size; g.count = count;
  g.robust_kind = robust_kind; g.robust_param = robust_param;
  g.keys = (const int2*)keys; g.meas = meas; g.noise = noise; g.cal_index = cal_index; g.body = body; g.J = (double*)J;
  return g;
}
void emu_linearize(int type, int f32, int count, int noise_kind, int per_factor, int noise_size, int robust_kind, double robust_param,
                   const int* keys, const double* meas, const double* noise, const int* cal_index, const double* body, void* J,
                   const double* values, const int* val_off, const double* cal) {
  GroupView g = mkgroup(type, count, noise_kind, per_factor, noise_size, robust_kind, robust_param, keys, meas, noise, cal_index, body, J);
  EvalCtx c; c.values = values; c.val_off = val_off; c.cal = cal;
  const int nb = (count + 127) / 128;
#define X(T) if (type == T) { if (f32) EMU_LAUNCH((linearize_kernel<T, float>), nb, 128, g, c); else EMU_LAUNCH((linearize_kernel<T, double>), nb, 128, g, c); }
  EMU_TYPES(X)
#undef X
}
void emu_retract(const double* values, const double* delta, const int* val_off, const int* var_dof, const int* var_type, int nvars, double* out) {
  EMU_LAUNCH(retract_kernel, (nvars + 127) / 128, 128, values, delta, val_off, var_dof, var_type, nvars, out);
}
double emu_error(int type, int count, int noise_kind, int per_factor, int noise_size, int robust_kind, double robust_param,
                 const int* keys, const double* meas, const double* noise, const int* cal_index, const double* body,
                 const double* values, const int* val_off, const double* cal) {
  GroupView g = mkgroup(type, count, noise_kind, per_factor, noise_size, robust_kind, robust_param, keys, meas, noise, cal_index, body, nullptr);
  EvalCtx c; c.values = values; c.val_off = val_off; c.cal = cal;
  double partials[64], out = 0;
  unsigned counter = 0;
#define X(T) if (type == T) EMU_LAUNCH((error_kernel<T>), 5, 256, g, c, partials, &counter, &out, 0);
  EMU_TYPES(X)
#undef X
  return out;
}
}
"""


def _extract_template(text, first_line_pat):
    m = re.search(first_line_pat, text, re.M)
    assert m, first_line_pat
    end = re.compile(r"^\}\s*$", re.M).search(text, m.start())
    return text[m.start():end.end()] + "\n"


@pytest.fixture(scope="module")
def emu_eval():
    kern = open(os.path.join(ROOT, "gtsam_b200", "csrc", "kernels.cuh")).read()
    eng = open(os.path.join(ROOT, "gtsam_b200", "csrc", "engine.cuh")).read()
    src = ("#include <cuda_runtime.h>\n#undef __global__\n#undef __device__\n#undef __forceinline__\n#undef __launch_bounds__\n"
           "#undef __restrict__\n#undef __shared__\n#include \"cuda_emu_prelude.h\"\n#include \"factors.cuh\"\nusing namespace b200;\n")
    src += _extract(eng, r"^struct GroupView \{")
    src += _extract_template(kern, r"^template <int TYPE, typename JT = double, int MINB = [^\n]*\n__global__ void __launch_bounds__\(128, MINB\) linearize_kernel")
    src += _extract_template(kern, r"^template <int TYPE>\n__global__ void __launch_bounds__\(256\) error_kernel\(")
    src += _extract_template(kern, r"^__global__ void retract_kernel\(")
    src += EVAL_WRAPPERS
    td = tempfile.mkdtemp()
    cpp, so = os.path.join(td, "emu_eval.cpp"), os.path.join(td, "libemu_eval.so")
    open(cpp, "w").write(src)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-I", os.path.join(ROOT, "tests", "emu"),
                           "-I", os.path.join(ROOT, "gtsam_b200", "csrc"), cpp, "-o", so])
    lib = C.CDLL(so)
    lib.emu_error.restype = C.c_double
    return lib


@pytest.mark.parametrize("case", util.CASES + util.EXTRA_CASES)
def test_factor_evaluators_emulated_on_host(emu_eval, case):
    """Every device evaluator (Between / Prior on Pose3 and Pose2, projection with and without body_P_sensor, Bundler SfM,
    camera prior; Unit / Isotropic / Diagonal / Gaussian whitening, the robust losses) through linearize_kernel and
    error_kernel, verbatim, against the unmodified reference's whitened [A|b] and graph error; and the FP32-storage
    instantiation (float SoA) against the FP32 protocol."""
    from gtsam_b200 import problem as P
    prob = util.load_case(case)
    ref = util.golden(case, "dump0")
    val_off = prob.val_offsets().astype(np.int32)
    values = np.ascontiguousarray(prob.values)
    cal = np.ascontiguousarray(prob.cal).ravel() if prob.cal.size else np.zeros(5)
    total = 0.0
    for gi, g in enumerate(prob.groups):
        d, nc = P.FACTOR_DIM[g.type], P.factor_ncols(g.type)
        keys = np.full((g.count, 2), -1, dtype=np.int32)
        keys[:, :g.keys.shape[1]] = g.keys
        meas = np.ascontiguousarray(g.meas).ravel()
        noise = np.ascontiguousarray(g.noise).ravel() if g.noise is not None and g.noise.size else np.zeros(1)
        pay = P.noise_payload(g.noise_kind, d)
        ci = None if g.cal_index is None else np.ascontiguousarray(g.cal_index, dtype=np.int32)
        body = None if g.body_P_sensor is None else np.ascontiguousarray(g.body_P_sensor)
        common = (g.count, g.noise_kind, g.noise_per_factor, pay, g.robust_kind, C.c_double(g.robust_param), _ip32(keys), _dp(meas), _dp(noise),
                  _ip32(ci) if ci is not None else None, _dp(body) if body is not None else None)
        R = util.ref_jacobians(prob, ref, gi)
        for f32 in (0, 1):
            J = np.zeros(g.count * d * nc, dtype=np.float32 if f32 else np.float64)
            emu_eval.emu_linearize(g.type, f32, *common, J.ctypes.data_as(C.c_void_p), _dp(values), _ip32(val_off), _dp(cal))
            W = J.astype(np.float64).reshape(d * nc, g.count).T.reshape(g.count, nc, d).transpose(0, 2, 1)
            assert util.relmax(W, R) <= (1e-6 if f32 else 1e-12), (gi, f32)
        total += emu_eval.emu_error(g.type, *common, _dp(values), _ip32(val_off), _dp(cal))
    assert abs(total - ref["error"][0]) <= 1e-12 * abs(ref["error"][0])
    if ref["status"][0] == 0:      # retract_kernel: Values::retract(delta) of the reference's own delta
        out = np.zeros_like(values)
        var_dof = prob.dof_offsets().astype(np.int32)
        vt = np.ascontiguousarray(prob.var_type, dtype=np.int32)
        delta = np.ascontiguousarray(ref["delta"])
        emu_eval.emu_retract(_dp(values), _dp(delta), _ip32(val_off), _ip32(var_dof), _ip32(vt), prob.nvars, _dp(out))
        assert util.relmax(out, ref["new_values"]) <= 1e-12


# ---- the typed consumers of the Jacobians: assemble_kernel / hdiag_kernel / linerr_kernel <TYPE, JT>, FP64 and FP32 storage ----
CONSUMER_WRAPPERS = r"""
extern "C" {
static GroupView mkgroup2(int type, int count, const int* keys, const int* scat, void* J) {
  GroupView g;
  memset(&g, 0, sizeof g);
  g.type = type; g.count = count; g.keys = (const int2*)keys; g.scat = (const int4*)scat; g.J = (double*)J;
  return g;
}
void emu_consumers(int type, int f32, int count, const int* keys, const int* scat, void* J, double* arena, const int64_t* off, const int* nf,
                   const int* ns, const int* var_dof, double* hdiag, const double* delta, double* e0, double* e1, int accumulate) {
  GroupView g = mkgroup2(type, count, keys, scat, J);
  TreeView t;
  memset(&t, 0, sizeof t);
  t.arena = arena; t.off = off; t.nf = nf; t.ns = ns;
  double p0[64], p1[64];
  unsigned counters[2] = {0, 0};
  const int nb = (count + 127) / 128;
#define X(T)                                                                                                       \
  if (type == T) {                                                                                                 \
    if (f32) {                                                                                                     \
      EMU_LAUNCH((assemble_kernel<T, float>), nb, 128, g, t);                                                      \
      EMU_LAUNCH((hdiag_kernel<T, float>), nb, 128, g, var_dof, hdiag);                                            \
      EMU_LAUNCH((linerr_kernel<T, float>), 3, 256, g, delta, var_dof, p0, p1, counters, e0, e1, accumulate, 1.0);  \
    } else {                                                                                                       \
      EMU_LAUNCH((assemble_kernel<T, double>), nb, 128, g, t);                                                     \
      EMU_LAUNCH((hdiag_kernel<T, double>), nb, 128, g, var_dof, hdiag);                                           \
      EMU_LAUNCH((linerr_kernel<T, double>), 3, 256, g, delta, var_dof, p0, p1, counters, e0, e1, accumulate, 1.0); \
    }                                                                                                              \
  }
  X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#undef X
}
}
"""


@pytest.fixture(scope="module")
def emu_cons():
    kern = open(os.path.join(ROOT, "gtsam_b200", "csrc", "kernels.cuh")).read()
    eng = open(os.path.join(ROOT, "gtsam_b200", "csrc", "engine.cuh")).read()
    src = ("#include <cuda_runtime.h>\n#undef __global__\n#undef __device__\n#undef __forceinline__\n#undef __launch_bounds__\n"
           "#undef __restrict__\n#undef __shared__\n#include \"cuda_emu_prelude.h\"\n#include \"factors.cuh\"\nusing namespace b200;\n")
    src += _extract(eng, r"^struct GroupView \{") + _extract(eng, r"^struct TreeView \{")
    src += _extract_template(kern, r"^template <int D, int NA, int NB_>\n__device__ __forceinline__ void add_block\(")
    src += _extract_template(kern, r"^template <int TYPE, typename JT = double>\n__global__ void __launch_bounds__\(128\) assemble_kernel\(")
    src += _extract_template(kern, r"^template <int TYPE, typename JT = double>\n__global__ void __launch_bounds__\(128\) hdiag_kernel\(")
    src += _extract_template(kern, r"^template <int TYPE, typename JT = double>\n__global__ void __launch_bounds__\(256\) linerr_kernel\(")
    src += CONSUMER_WRAPPERS
    td = tempfile.mkdtemp()
    cpp, so = os.path.join(td, "emu_cons.cpp"), os.path.join(td, "libemu_cons.so")
    open(cpp, "w").write(src)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-I", os.path.join(ROOT, "tests", "emu"),
                           "-I", os.path.join(ROOT, "gtsam_b200", "csrc"), cpp, "-o", so])
    return C.CDLL(so)


@pytest.mark.parametrize("case", ["bal_tiny_s2", "bal_tiny_bundler", "sphere_tiny_gaussian", "pose2_ring", "bal_small_metis"])
@pytest.mark.parametrize("f32", [0, 1])
def test_typed_jacobian_consumers_emulated_on_host(emu_cons, case, f32, built):
    """assemble_kernel / hdiag_kernel / linerr_kernel <TYPE, JT> verbatim, in both storage modes, fed with the reference's
    whitened [A|b] (rounded to float for JT = float) and the scatter tables of the product's host symbolic phase: the
    assembled fronts sum to numpy's global augmented Hessian of the same (rounded) blocks, the Hessian diagonal and the
    two linear errors likewise."""
    from gtsam_b200 import problem as P
    prob = util.load_case(case)
    ref = util.golden(case, "dump0")
    L = capi.lib()
    desc, keep = prob.c_desc()
    h = C.c_void_p()
    capi._check(L.b200_symbolic_create(C.byref(desc), C.byref(h)))
    info = P.CSymbolicInfo()
    L.b200_symbolic_get_info(h, C.byref(info))
    # the scatter tables address SUPERNODES (the reference's cliques after relaxed amalgamation): what the device assembles into
    fp, sp = np.zeros(info.supernodes + 1, dtype=np.int64), np.zeros(info.supernodes + 1, dtype=np.int64)
    fv, sv = np.zeros(info.supernode_frontal_list_len, dtype=np.int64), np.zeros(max(1, info.supernode_separator_list_len), dtype=np.int64)
    par = np.zeros(info.supernodes, dtype=np.int64)
    L.b200_symbolic_get_supernodes(h, capi._ip(fp), capi._ip(fv), capi._ip(sp), capi._ip(sv), capi._ip(par))
    arity = np.zeros(prob.nfactors, dtype=np.int64)
    for g in prob.groups:
        pos = g.graph_index if g.graph_index is not None else g.graph_index0 + np.arange(g.count)
        arity[pos] = g.keys.shape[1]
    fptr = np.concatenate([[0], np.cumsum(arity)]).astype(np.int64)
    fclique, fslots = np.zeros(prob.nfactors, dtype=np.int32), np.zeros(int(fptr[-1]), dtype=np.int32)
    L.b200_symbolic_get_factor_slots(h, _ip32(fclique), _ip32(fslots))
    L.b200_symbolic_destroy(h)
    dims, dof = prob.var_dims.astype(np.int64), prob.dof_offsets()
    nc, ndelta = len(par), int(dof[-1])
    nf = np.array([dims[fv[fp[c]:fp[c + 1]]].sum() for c in range(nc)], dtype=np.int32)
    ns = np.array([dims[sv[sp[c]:sp[c + 1]]].sum() for c in range(nc)], dtype=np.int32)
    nn = (nf + ns + 1).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(nn * nn)]).astype(np.int64)
    arena, hdiag = np.zeros(int(off[-1])), np.zeros(ndelta)
    var_dof = dof.astype(np.int32)
    delta = np.ascontiguousarray(ref["delta"]) if ref["status"][0] == 0 else np.linspace(-0.1, 0.1, ndelta)
    e0, e1 = C.c_double(0), C.c_double(0)
    Hglob = np.zeros((ndelta + 1, ndelta + 1))
    keep2 = []
    for gi, g in enumerate(prob.groups):
        d, ncol = P.FACTOR_DIM[g.type], P.factor_ncols(g.type)
        W = util.ref_jacobians(prob, ref, gi)                          # (count, d, ncol) whitened, FP64
        if f32:
            W = W.astype(np.float32).astype(np.float64)
        soa = np.ascontiguousarray(W.transpose(0, 2, 1).reshape(g.count, d * ncol).T)     # [e = r + c*d][f]
        J = soa.astype(np.float32) if f32 else soa
        pos = g.graph_index if g.graph_index is not None else g.graph_index0 + np.arange(g.count)
        keys = np.full((g.count, 2), -1, dtype=np.int32)
        keys[:, :g.keys.shape[1]] = g.keys
        scat = np.zeros((g.count, 4), dtype=np.int32)
        scat[:, 0] = fclique[pos]
        scat[:, 1] = fslots[fptr[pos]]
        scat[:, 2] = np.where(arity[pos] == 2, fslots[np.minimum(fptr[pos] + 1, len(fslots) - 1)], -1)
        keep2 += [J, keys, scat]
        emu_cons.emu_consumers(g.type, f32, g.count, _ip32(keys), _ip32(scat), J.ctypes.data_as(C.c_void_p), _dp(arena),
                               off.ctypes.data_as(C.POINTER(C.c_int64)), _ip32(nf), _ip32(ns), _ip32(var_dof), _dp(hdiag), _dp(delta),
                               C.byref(e0), C.byref(e1), 0 if gi == 0 else 1)
        for i in range(g.count):
            idx = np.concatenate([np.arange(dof[k], dof[k + 1]) for k in g.keys[i]] + [[ndelta]])
            Hglob[np.ix_(idx, idx)] += W[i].T @ W[i]
    Hfront = np.zeros_like(Hglob)
    for c in range(nc):
        didx = np.concatenate([np.arange(dof[v], dof[v + 1]) for v in list(fv[fp[c]:fp[c + 1]]) + list(sv[sp[c]:sp[c + 1]])] + [[ndelta]])
        F = arena[off[c]:off[c + 1]].reshape(nn[c], nn[c]).T
        Hfront[np.ix_(didx, didx)] += np.triu(F) + np.triu(F, 1).T
    assert np.abs(Hfront - Hglob).max() <= 1e-12 * np.abs(Hglob).max()
    assert util.relmax(hdiag, np.diag(Hglob)[:ndelta]) <= 1e-13
    x = np.concatenate([delta, [-1.0]])
    assert abs(e0.value - 0.5 * Hglob[ndelta, ndelta]) <= 1e-12 * abs(e0.value)
    assert abs(e1.value - 0.5 * x @ Hglob @ x) <= 1e-9 * abs(e0.value)
    if not f32:
        assert util.relmax(hdiag, ref["hessian_diagonal"]) <= 1e-12
        if ref["status"][0] == 0:
            assert abs(e0.value - ref["linear_error_zero"][0]) <= 1e-11 * e0.value
