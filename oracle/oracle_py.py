"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module (see oracle/oracle.h).  The product
package gtsam_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gtsam_b200 import problem as P

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.orc_problem_create.argtypes = [C.POINTER(P.CProblemDesc), C.POINTER(C.c_void_p)]
        from gtsam_b200 import linear as LN
        L.orc_linear_create.argtypes = [C.POINTER(LN.CLinearDesc), C.POINTER(C.c_void_p)]
        L.orc_linear_update.argtypes = [C.c_void_p, C.c_int64, dp, dp]
        L.orc_linear_update_hessian.argtypes = [C.c_void_p, C.c_int64, dp]
        L.orc_problem_destroy.argtypes = [C.c_void_p]
        L.orc_set_values.argtypes = [C.c_void_p, dp]
        L.orc_get_values.argtypes = [C.c_void_p, dp]
        L.orc_values_size.argtypes = [C.c_void_p]
        L.orc_values_size.restype = C.c_int64
        L.orc_delta_size.argtypes = [C.c_void_p]
        L.orc_delta_size.restype = C.c_int64
        L.orc_error.argtypes = [C.c_void_p]
        L.orc_error.restype = C.c_double
        L.orc_linearize.argtypes = [C.c_void_p]
        L.orc_set_jacobian_fp32.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_jacobians.argtypes = [C.c_void_p, C.c_int64, dp]
        L.orc_hessian_diagonal.argtypes = [C.c_void_p, dp]
        L.orc_solve.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double, dp, dp, ip]
        L.orc_get_delta.argtypes = [C.c_void_p, dp]
        L.orc_try_step.argtypes = [C.c_void_p]
        L.orc_try_step.restype = C.c_double
        L.orc_accept_step.argtypes = [C.c_void_p]
        L.orc_gn_iterate.argtypes = [C.c_void_p, dp]
        L.orc_dogleg_iterate.argtypes = [C.c_void_p, dp, dp]
        L.orc_marginal_covariance.argtypes = [C.c_void_p, C.c_int64, dp]
        L.orc_solve_rhs.argtypes = [C.c_void_p, dp, dp]
        L.orc_joint_marginal_covariance.argtypes = [C.c_void_p, ip, C.c_int64, dp]
        L.orc_symbolic_info_get.argtypes = [C.c_void_p, C.POINTER(P.CSymbolicInfo)]
        L.orc_get_cliques.argtypes = [C.c_void_p, ip, ip, ip, ip, ip]
        L.orc_get_conditional.argtypes = [C.c_void_p, C.c_int64, dp]
        L.orc_cholesky_partial.argtypes = [dp, C.c_int64, C.c_int64]
        for fn in ("orc_so3_expmap", "orc_so3_logmap", "orc_pose3_expmap", "orc_pose3_logmap",
                   "orc_pose3_inverse", "orc_pose3_adjoint_map"):
            getattr(L, fn).argtypes = [dp, dp]
        L.orc_pose3_compose.argtypes = [dp, dp, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class OrcLM(C.Structure):
    _fields_ = [("prob", C.c_void_p), ("params", P.CLMParams), ("state", P.CLMState)]


def unary(fn, x, nout):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(nout)
    getattr(lib(), fn)(_dp(x), _dp(out))
    return out


def pose3_compose(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros(12)
    lib().orc_pose3_compose(_dp(a), _dp(b), _dp(out))
    return out


def cholesky_partial(M, nfrontal):
    """M: (n,n) symmetric (upper used). Returns (ok, result col-major as (n,n) array)."""
    A = np.asfortranarray(np.array(M, dtype=np.float64))
    ok = lib().orc_cholesky_partial(A.ctypes.data_as(C.POINTER(C.c_double)), A.shape[0], nfrontal)
    return bool(ok), np.array(A)


class OracleProblem:
    """CPU oracle with the same surface as gtsam_b200.capi.DeviceProblem."""

    def __init__(self, prob: P.Problem):
        self.prob = prob
        self.L = lib()
        desc, self._keep = prob.c_desc()
        h = C.c_void_p()
        st = self.L.orc_problem_create(C.byref(desc), C.byref(h))
        if st != 0:
            raise RuntimeError(f"orc_problem_create failed: {st}")
        self.h = h
        self.nval = self.L.orc_values_size(h)
        self.ndelta = self.L.orc_delta_size(h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_problem_destroy(self.h)
            self.h = None

    def set_values(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.nval
        self.L.orc_set_values(self.h, _dp(v))

    def get_values(self):
        out = np.zeros(self.nval)
        self.L.orc_get_values(self.h, _dp(out))
        return out

    def error(self):
        return float(self.L.orc_error(self.h))

    def linearize(self):
        self.L.orc_linearize(self.h)

    def set_jacobian_precision(self, fp32: bool):
        self.L.orc_set_jacobian_fp32(self.h, int(fp32))

    def get_jacobians(self, group):
        g = self.prob.groups[group]
        d, nc = P.FACTOR_DIM[g.type], P.factor_ncols(g.type)
        out = np.zeros(g.count * d * nc)
        self.L.orc_get_jacobians(self.h, group, _dp(out))
        # factor-major, each block column-major d x nc -> (count, d, nc)
        return out.reshape(g.count, nc, d).transpose(0, 2, 1)

    def hessian_diagonal(self):
        out = np.zeros(self.ndelta)
        self.L.orc_hessian_diagonal(self.h, _dp(out))
        return out

    def solve(self, lam=0.0, diagonal_damping=False, min_diagonal=1e-6, max_diagonal=1e32):
        e0, e1, fv = C.c_double(), C.c_double(), C.c_int64(-1)
        st = self.L.orc_solve(self.h, lam, int(diagonal_damping), min_diagonal, max_diagonal,
                              C.byref(e0), C.byref(e1), C.byref(fv))
        return st, e0.value, e1.value, fv.value

    def get_delta(self):
        out = np.zeros(self.ndelta)
        self.L.orc_get_delta(self.h, _dp(out))
        return out

    def try_step(self):
        return float(self.L.orc_try_step(self.h))

    def accept_step(self):
        self.L.orc_accept_step(self.h)

    def gn_iterate(self):
        e = C.c_double()
        st = self.L.orc_gn_iterate(self.h, C.byref(e))
        return st, e.value

    def dogleg_iterate(self, error, delta):
        e, d = C.c_double(error), C.c_double(delta)
        st = self.L.orc_dogleg_iterate(self.h, C.byref(e), C.byref(d))
        return st, e.value, d.value

    def marginal_covariance(self, var):
        d = int(self.prob.var_dims[var])
        out = np.zeros(d * d)
        st = self.L.orc_marginal_covariance(self.h, int(var), _dp(out))
        return st, out.reshape(d, d).T   # column-major -> (row, col)

    def joint_marginal_covariance(self, variables):
        vs = np.array(sorted(int(v) for v in variables), dtype=np.int64)
        D = int(sum(int(self.prob.var_dims[v]) for v in vs))
        out = np.zeros(D * D)
        st = self.L.orc_joint_marginal_covariance(self.h, _ip(vs), len(vs), _dp(out))
        return st, out.reshape(D, D).T

    def solve_rhs(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        x = np.zeros(self.ndelta)
        self.L.orc_solve_rhs(self.h, _dp(g), _dp(x))
        return x

    def symbolic_info(self):
        info = P.CSymbolicInfo()
        self.L.orc_symbolic_info_get(self.h, C.byref(info))
        return info

    def cliques(self):
        info = self.symbolic_info()
        fp = np.zeros(info.ncliques + 1, dtype=np.int64)
        sp = np.zeros(info.ncliques + 1, dtype=np.int64)
        fv = np.zeros(max(1, info.frontal_list_len), dtype=np.int64)
        sv = np.zeros(max(1, info.separator_list_len), dtype=np.int64)
        par = np.zeros(max(1, info.ncliques), dtype=np.int64)
        self.L.orc_get_cliques(self.h, _ip(fp), _ip(fv), _ip(sp), _ip(sv), _ip(par))
        return fp, fv[:info.frontal_list_len], sp, sv[:info.separator_list_len], par[:info.ncliques]

    def conditional(self, c):
        fp, fv, sp, sv, _ = self.cliques()
        dims = self.prob.var_dims
        f = int(dims[fv[fp[c]:fp[c + 1]]].sum())
        s = int(dims[sv[sp[c]:sp[c + 1]]].sum())
        out = np.zeros(f * (f + s + 1))
        self.L.orc_get_conditional(self.h, c, _dp(out))
        return out.reshape(f + s + 1, f).T

    # -- LM ---------------------------------------------------------------------
    def lm(self, params: P.CLMParams):
        lm = OrcLM()
        self.L.orc_lm_init(C.byref(lm), self.h, C.byref(params))
        return lm

    def lm_iterate(self, lm):
        return self.L.orc_lm_iterate(C.byref(lm))

    def lm_optimize(self, lm):
        return self.L.orc_lm_optimize(C.byref(lm))


class OracleLinearProblem(OracleProblem):
    """CPU oracle of the GaussianFactorGraph level (orc_linear_create): same surface as
    gtsam_b200.capi.LinearDeviceProblem."""

    def __init__(self, lprob):
        self.prob = lprob
        self.L = lib()
        desc, self._keep = lprob.c_desc()
        h = C.c_void_p()
        st = self.L.orc_linear_create(C.byref(desc), C.byref(h))
        if st != 0:
            raise RuntimeError(f"orc_linear_create failed: {st}")
        self.h = h
        self.nval = 0
        self.ndelta = self.L.orc_delta_size(h)

    def update(self, group, Ab, sigmas=None):
        Ab = np.ascontiguousarray(Ab, dtype=np.float64)
        sp = None
        if sigmas is not None:
            sigmas = np.ascontiguousarray(sigmas, dtype=np.float64)
            sp = _dp(sigmas)
        assert self.L.orc_linear_update(self.h, group, _dp(Ab), sp) == 0

    def update_hessian(self, hgroup, info):
        info = np.ascontiguousarray(info, dtype=np.float64)
        assert self.L.orc_linear_update_hessian(self.h, hgroup, _dp(info)) == 0

    def get_jacobians(self, group):
        g = self.prob.groups[group]
        out = np.zeros(g.count * g.rows * g.ncols)
        self.L.orc_get_jacobians(self.h, group, _dp(out))
        return out.reshape(g.count, g.ncols, g.rows).transpose(0, 2, 1)
