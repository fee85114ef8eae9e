"""ctypes binding of the C-ABI shared library (include/gtsam_b200.h).

The library is the product; this module only loads it and marshals numpy
buffers.  It fails loudly when the CUDA extension is missing or no GPU is
visible — there is no CPU fallback (the oracle under oracle/ is test
infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import problem as P

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgtsam_b200.so")
_LIB = None

EXPORTS = [
    "b200_var_storage", "b200_var_dim", "b200_factor_arity", "b200_factor_meas_size", "b200_factor_dim",
    "b200_ctx_create", "b200_ctx_destroy", "b200_last_error_string", "b200_launch_count", "b200_ctx_stream",
    "b200_problem_create", "b200_problem_destroy", "b200_set_values", "b200_set_group_noise", "b200_gradient_at_zero", "b200_linear_graph_error", "b200_get_values", "b200_values_size",
    "b200_delta_size", "b200_error", "b200_linearize", "b200_get_jacobians", "b200_hessian_diagonal",
    "b200_solve", "b200_get_delta", "b200_try_step", "b200_accept_step", "b200_lm_params_legacy",
    "b200_lm_params_ceres", "b200_lm_create", "b200_lm_destroy", "b200_lm_iterate", "b200_lm_optimize",
    "b200_lm_get_state", "b200_lm_reset", "b200_gn_iterate", "b200_symbolic_info_get", "b200_get_cliques",
    "b200_get_conditional", "b200_shared_front_buffer", "b200_save_values", "b200_restore_values",
    "b200_synchronize", "b200_profile_enable", "b200_profile_phase_count", "b200_profile_phase_name",
    "b200_profile_get", "b200_symbolic_create", "b200_symbolic_destroy", "b200_symbolic_get_info",
    "b200_symbolic_get_cliques", "b200_symbolic_get_levels", "b200_nccl_unique_id", "b200_ctx_comm_init",
    "b200_shard_plan", "b200_dl_create", "b200_dl_destroy", "b200_dl_iterate", "b200_dl_get_state", "b200_marginal_covariance", "b200_joint_marginal_covariance",
    "b200_linear_create", "b200_linear_update", "b200_linear_update_hessian", "b200_linear_symbolic_create",
    "b200_set_jacobian_precision", "b200_get_jacobian_precision", "b200_set_tuning", "b200_symbolic_get_factor_slots",
    "b200_get_supernodes", "b200_symbolic_get_supernodes", "b200_symbolic_get_clique_supernode", "b200_measure_fp64_peak",
    "b200_values_view", "b200_set_values_view", "b200_get_values_view", "b200_get_values_all",
]


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gtsam_b200 status {code}: {msg}")
        self.code = code


class IndeterminantLinearSystemException(B200Error):
    """Mirror of gtsam::IndeterminantLinearSystemException (gtsam/linear/linearExceptions.h)."""

    def __init__(self, var):
        B200Error.__init__(self, P.INDETERMINATE, f"indeterminant linear system near variable {var}")
        self.nearby_variable = var


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(gtsam_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        vp = C.c_void_p
        L.b200_last_error_string.restype = C.c_char_p
        L.b200_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.b200_ctx_destroy.argtypes = [vp]
        L.b200_launch_count.argtypes = [vp]
        L.b200_launch_count.restype = C.c_int64
        L.b200_ctx_stream.argtypes = [vp]
        L.b200_ctx_stream.restype = vp
        L.b200_problem_create.argtypes = [vp, C.POINTER(P.CProblemDesc), C.POINTER(vp)]
        L.b200_problem_destroy.argtypes = [vp]
        L.b200_set_values.argtypes = [vp, dp]
        L.b200_gradient_at_zero.argtypes = [vp, dp]
        L.b200_linear_graph_error.argtypes = [vp, dp, dp]
        L.b200_set_group_noise.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, dp]
        L.b200_get_values.argtypes = [vp, dp]
        L.b200_values_size.argtypes = [vp]
        L.b200_values_size.restype = C.c_int64
        L.b200_delta_size.argtypes = [vp]
        L.b200_delta_size.restype = C.c_int64
        L.b200_error.argtypes = [vp, dp]
        L.b200_linearize.argtypes = [vp]
        L.b200_get_jacobians.argtypes = [vp, C.c_int64, dp]
        L.b200_hessian_diagonal.argtypes = [vp, dp]
        L.b200_solve.argtypes = [vp, C.c_double, C.c_int, C.c_double, C.c_double, dp, dp, ip]
        L.b200_get_delta.argtypes = [vp, dp]
        L.b200_try_step.argtypes = [vp, dp]
        L.b200_accept_step.argtypes = [vp]
        L.b200_lm_params_legacy.argtypes = [C.POINTER(P.CLMParams)]
        L.b200_lm_params_ceres.argtypes = [C.POINTER(P.CLMParams)]
        L.b200_lm_create.argtypes = [vp, C.POINTER(P.CLMParams), C.POINTER(vp)]
        L.b200_lm_destroy.argtypes = [vp]
        L.b200_lm_iterate.argtypes = [vp]
        L.b200_lm_optimize.argtypes = [vp]
        L.b200_lm_get_state.argtypes = [vp, C.POINTER(P.CLMState)]
        L.b200_lm_reset.argtypes = [vp]
        L.b200_gn_iterate.argtypes = [vp, dp]
        L.b200_marginal_covariance.argtypes = [vp, C.c_int64, dp]
        L.b200_joint_marginal_covariance.argtypes = [vp, ip, C.c_int64, dp]
        L.b200_dl_create.argtypes = [vp, C.c_double, C.POINTER(vp)]
        L.b200_dl_destroy.argtypes = [vp]
        L.b200_dl_iterate.argtypes = [vp]
        L.b200_dl_get_state.argtypes = [vp, dp, dp, C.POINTER(C.c_int32)]
        L.b200_symbolic_info_get.argtypes = [vp, C.POINTER(P.CSymbolicInfo)]
        L.b200_get_cliques.argtypes = [vp, ip, ip, ip, ip, ip]
        L.b200_get_conditional.argtypes = [vp, C.c_int64, dp]
        L.b200_shared_front_buffer.argtypes = [vp, C.POINTER(vp), ip]
        L.b200_save_values.argtypes = [vp]
        L.b200_restore_values.argtypes = [vp]
        L.b200_synchronize.argtypes = [vp]
        L.b200_profile_enable.argtypes = [vp, C.c_int]
        L.b200_profile_phase_name.argtypes = [C.c_int]
        L.b200_profile_phase_name.restype = C.c_char_p
        L.b200_profile_get.argtypes = [vp, dp, ip]
        L.b200_nccl_unique_id.argtypes = [C.c_char_p]
        L.b200_ctx_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
        L.b200_shard_plan.argtypes = [C.POINTER(P.CProblemDesc), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.b200_symbolic_create.argtypes = [C.POINTER(P.CProblemDesc), C.POINTER(vp)]
        L.b200_symbolic_destroy.argtypes = [vp]
        L.b200_symbolic_get_info.argtypes = [vp, C.POINTER(P.CSymbolicInfo)]
        L.b200_set_jacobian_precision.argtypes = [vp, C.c_int]
        L.b200_get_jacobian_precision.argtypes = [vp]
        L.b200_set_tuning.argtypes = [vp, C.c_char_p, C.c_int64]
        from . import linear as LN
        L.b200_linear_create.argtypes = [vp, C.POINTER(LN.CLinearDesc), C.POINTER(vp)]
        L.b200_linear_update.argtypes = [vp, C.c_int64, dp, dp]
        L.b200_linear_update_hessian.argtypes = [vp, C.c_int64, dp]
        L.b200_linear_symbolic_create.argtypes = [C.POINTER(LN.CLinearDesc), C.POINTER(vp)]
        L.b200_symbolic_get_cliques.argtypes = [vp, ip, ip, ip, ip, ip]
        L.b200_symbolic_get_supernodes.argtypes = [vp, ip, ip, ip, ip, ip]
        L.b200_symbolic_get_clique_supernode.argtypes = [vp, C.POINTER(C.c_int32)]
        L.b200_values_view.argtypes = [vp, C.c_int, ip, ip, ip]
        L.b200_set_values_view.argtypes = [vp, dp]
        L.b200_get_values_view.argtypes = [vp, dp]
        L.b200_get_values_all.argtypes = [vp, dp]
        L.b200_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.b200_symbolic_get_factor_slots.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _check(rc):
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error_string().decode())


def nccl_unique_id() -> bytes:
    """128-byte NCCL id (rank 0 creates it; ship it to the other ranks)."""
    buf = C.create_string_buffer(128)
    _check(lib().b200_nccl_unique_id(buf))
    return buf.raw


def shard_plan(prob: P.Problem, world: int):
    """(clique_owner, factor_owner) of the sharding b200_problem_create applies at `world`
    ranks; host only (no GPU needed).  clique_owner = -1 for the replicated top."""
    L = lib()
    desc, keep = prob.c_desc()
    h = C.c_void_p()
    _check(L.b200_symbolic_create(C.byref(desc), C.byref(h)))
    info = P.CSymbolicInfo()
    L.b200_symbolic_get_info(h, C.byref(info))
    L.b200_symbolic_destroy(h)
    co = np.zeros(max(1, info.ncliques), dtype=np.int32)
    fo = np.zeros(max(1, prob.nfactors), dtype=np.int32)
    _check(L.b200_shard_plan(C.byref(desc), world, co.ctypes.data_as(C.POINTER(C.c_int32)),
                             fo.ctypes.data_as(C.POINTER(C.c_int32))))
    return co[:info.ncliques], fo[:prob.nfactors]


class Context:
    """One CUDA device + stream (b200_ctx)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        h = C.c_void_p()
        _check(self.L.b200_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self.rank, self.world = 0, 1
        self._problems = weakref.WeakSet()

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Join the NCCL communicator (one process per GPU); call before creating problems."""
        _check(self.L.b200_ctx_comm_init(self.h, unique_id, rank, world))
        self.rank, self.world = rank, world

    def measure_fp64_peak(self):
        """(DMMA TFLOP/s, FMA-pipe TFLOP/s) measured on this device: roofline denominators of the dense-front kernels."""
        a, b = C.c_double(), C.c_double()
        _check(self.L.b200_measure_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self) -> int:
        return int(self.L.b200_launch_count(self.h))

    def stream(self) -> int:
        return int(self.L.b200_ctx_stream(self.h) or 0)

    def close(self):
        if self.h:
            for p in list(self._problems):   # problems hold a pointer to the ctx: free them first
                p.close()
            self.L.b200_ctx_destroy(self.h)
            self.h = None


class DeviceProblem:
    """Device-resident problem (b200_problem): values, factor tables, junction tree."""

    def __init__(self, ctx: Context, prob: P.Problem):
        self.ctx, self.prob, self.L = ctx, prob, ctx.L
        desc, keep = prob.c_desc()
        h = C.c_void_p()
        _check(self.L.b200_problem_create(ctx.h, C.byref(desc), C.byref(h)))
        del keep
        self.h = h
        self.nval = int(self.L.b200_values_size(h))
        self.ndelta = int(self.L.b200_delta_size(h))
        ctx._problems.add(self)

    def close(self):
        if getattr(self, "h", None):
            if self.ctx.h:
                self.L.b200_problem_destroy(self.h)
            self.h = None

    # -- benchmark / profiling helpers ------------------------------------------------
    def save_values(self):
        _check(self.L.b200_save_values(self.h))

    def restore_values(self):
        _check(self.L.b200_restore_values(self.h))

    def synchronize(self):
        _check(self.L.b200_synchronize(self.h))

    def profile_enable(self, on=True):
        _check(self.L.b200_profile_enable(self.h, int(on)))

    def profile(self):
        """{phase: (milliseconds, calls)} accumulated since profile_enable()."""
        n = self.L.b200_profile_phase_count()
        ms, calls = np.zeros(16), np.zeros(16, dtype=np.int64)
        _check(self.L.b200_profile_get(self.h, _dp(ms), _ip(calls)))
        return {self.L.b200_profile_phase_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(n)}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_values(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.nval
        _check(self.L.b200_set_values(self.h, _dp(v)))

    def gradient_at_zero(self):
        """GaussianFactorGraph::gradientAtZero of the current linearization (-A'b), dof order."""
        out = np.empty(self.ndelta)
        _check(self.L.b200_gradient_at_zero(self.h, _dp(out)))
        return out

    def linear_graph_error(self, x):
        """GaussianFactorGraph::error(x) of the current linearization / of a linear problem's graph; x in dof order."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.ndelta
        e = C.c_double()
        _check(self.L.b200_linear_graph_error(self.h, _dp(x), C.cast(C.byref(e), C.POINTER(C.c_double))))
        return e.value

    def set_group_noise(self, gi: int, noise_kind: int, noise):
        """New noise model(s) on factor group ``gi`` (shared payload, or one per factor), as GncOptimizer's
        makeWeightedGraph does between outer iterations; the robust loss of the group is kept."""
        g = self.prob.groups[gi]
        pay = P.noise_payload(noise_kind, P.FACTOR_DIM[g.type])
        noise = np.zeros(0) if noise is None else np.ascontiguousarray(noise, dtype=np.float64).ravel()
        if noise.size not in (pay, pay * g.count):
            raise ValueError("noise payload size")
        per = int(pay > 0 and noise.size == pay * g.count and g.count > 1)
        _check(self.L.b200_set_group_noise(self.h, C.c_int64(gi), noise_kind, per, _dp(noise) if noise.size else None))

    def values_view(self, which: int):
        """Variable ids of a values view of a (sharded) problem: 0 = what this rank needs as input, 1 = what it owns."""
        nv, nd = C.c_int64(), C.c_int64()
        _check(self.L.b200_values_view(self.h, which, C.byref(nv), C.byref(nd), None))
        ids = np.zeros(max(1, nv.value), dtype=np.int64)
        _check(self.L.b200_values_view(self.h, which, C.byref(nv), C.byref(nd), _ip(ids)))
        return ids[:nv.value], nd.value

    def view_index(self, which: int):
        """Indices into the full packed Values of the doubles of a view (host-side gather / stitch)."""
        ids, nd = self.values_view(which)
        off = self.prob.val_offsets().astype(np.int64)
        lens = off[ids + 1] - off[ids]
        idx = np.repeat(off[ids], lens) + (np.arange(nd) - np.repeat(np.cumsum(lens) - lens, lens))
        assert idx.size == nd
        return idx.astype(np.int64)

    def set_values_view(self, packed):
        assert packed.dtype == np.float64 and packed.flags.c_contiguous
        _check(self.L.b200_set_values_view(self.h, _dp(packed)))

    def get_values_view(self, out):
        assert out.dtype == np.float64 and out.flags.c_contiguous
        _check(self.L.b200_get_values_view(self.h, _dp(out)))
        return out

    def get_values_all(self, out=None):
        """The whole packed Values on every rank of a sharded problem (owned views gathered by one all-reduce)."""
        if out is None:
            out = np.empty(self.nval)
        _check(self.L.b200_get_values_all(self.h, _dp(out)))
        return out

    def get_values(self, out=None):
        """Packed values; pass a page-locked float64 array as ``out`` for a direct D2H copy."""
        if out is None:
            out = np.empty(self.nval)
        assert out.dtype == np.float64 and out.size == self.nval and out.flags.c_contiguous
        _check(self.L.b200_get_values(self.h, _dp(out)))
        return out

    def error(self) -> float:
        e = C.c_double()
        _check(self.L.b200_error(self.h, C.byref(e)))
        return e.value

    def linearize(self):
        _check(self.L.b200_linearize(self.h))

    def set_jacobian_precision(self, fp32: bool):
        """FP32 storage of the whitened Jacobians ("FP32 linearize + FP64 solve", BASELINE configs[4])."""
        _check(self.L.b200_set_jacobian_precision(self.h, int(fp32)))

    def set_tuning(self, key: str, value: int):
        """Kernel-variant switch of this problem (b200_set_tuning): every variant computes the same result."""
        _check(self.L.b200_set_tuning(self.h, key.encode(), int(value)))

    def get_jacobians(self, group: int):
        g = self.prob.groups[group]
        d, nc = P.FACTOR_DIM[g.type], P.factor_ncols(g.type)
        out = np.zeros(g.count * d * nc)
        _check(self.L.b200_get_jacobians(self.h, group, _dp(out)))
        return out.reshape(g.count, nc, d).transpose(0, 2, 1)

    def hessian_diagonal(self):
        out = np.zeros(self.ndelta)
        _check(self.L.b200_hessian_diagonal(self.h, _dp(out)))
        return out

    def solve(self, lam=0.0, diagonal_damping=False, min_diagonal=1e-6, max_diagonal=1e32):
        """Returns (status, linear_error(0), linear_error(delta), fail_var)."""
        e0, e1, fv = C.c_double(), C.c_double(), C.c_int64(-1)
        rc = self.L.b200_solve(self.h, lam, int(diagonal_damping), min_diagonal, max_diagonal,
                               C.byref(e0), C.byref(e1), C.byref(fv))
        if rc not in (P.OK, P.INDETERMINATE):
            _check(rc)
        return rc, e0.value, e1.value, fv.value

    def get_delta(self):
        out = np.zeros(self.ndelta)
        _check(self.L.b200_get_delta(self.h, _dp(out)))
        return out

    def try_step(self) -> float:
        e = C.c_double()
        _check(self.L.b200_try_step(self.h, C.byref(e)))
        return e.value

    def accept_step(self):
        _check(self.L.b200_accept_step(self.h))

    def marginal_covariance(self, var: int):
        """(d, d) covariance of variable `var` at the current values (Marginals::marginalCovariance)."""
        d = int(self.prob.var_dims[var])
        out = np.zeros(d * d)
        rc = self.L.b200_marginal_covariance(self.h, int(var), _dp(out))
        if rc == P.INDETERMINATE:
            raise IndeterminantLinearSystemException(-1)
        _check(rc)
        return out.reshape(d, d).T      # column-major -> [row, col]

    def joint_marginal_covariance(self, variables):
        """(D, D) joint covariance, blocks in ascending variable order (Marginals::jointMarginalCovariance)."""
        vs = np.array(sorted(int(v) for v in variables), dtype=np.int64)
        D = int(sum(int(self.prob.var_dims[v]) for v in vs))
        out = np.zeros(D * D)
        rc = self.L.b200_joint_marginal_covariance(self.h, _ip(vs), len(vs), _dp(out))
        if rc == P.INDETERMINATE:
            raise IndeterminantLinearSystemException(-1)
        _check(rc)
        return out.reshape(D, D).T

    def gn_iterate(self):
        e = C.c_double()
        rc = self.L.b200_gn_iterate(self.h, C.byref(e))
        if rc not in (P.OK, P.INDETERMINATE):
            _check(rc)
        return rc, e.value

    def symbolic_info(self) -> P.CSymbolicInfo:
        info = P.CSymbolicInfo()
        _check(self.L.b200_symbolic_info_get(self.h, C.byref(info)))
        return info

    def cliques(self):
        info = self.symbolic_info()
        fp = np.zeros(info.ncliques + 1, dtype=np.int64)
        sp = np.zeros(info.ncliques + 1, dtype=np.int64)
        fv = np.zeros(max(1, info.frontal_list_len), dtype=np.int64)
        sv = np.zeros(max(1, info.separator_list_len), dtype=np.int64)
        par = np.zeros(max(1, info.ncliques), dtype=np.int64)
        _check(self.L.b200_get_cliques(self.h, _ip(fp), _ip(fv), _ip(sp), _ip(sv), _ip(par)))
        return fp, fv[:info.frontal_list_len], sp, sv[:info.separator_list_len], par[:info.ncliques]

    def supernodes(self):
        """The supernodes the device eliminates (the reference's cliques after relaxed amalgamation)."""
        info = self.symbolic_info()
        fp = np.zeros(info.supernodes + 1, dtype=np.int64)
        sp = np.zeros(info.supernodes + 1, dtype=np.int64)
        fv = np.zeros(max(1, info.supernode_frontal_list_len), dtype=np.int64)
        sv = np.zeros(max(1, info.supernode_separator_list_len), dtype=np.int64)
        par = np.zeros(max(1, info.supernodes), dtype=np.int64)
        _check(self.L.b200_get_supernodes(self.h, _ip(fp), _ip(fv), _ip(sp), _ip(sv), _ip(par)))
        return fp, fv[:info.supernode_frontal_list_len], sp, sv[:info.supernode_separator_list_len], par[:info.supernodes]

    def conditional(self, c: int):
        fp, fv, sp, sv, _ = self.cliques()
        dims = self.prob.var_dims
        f = int(dims[fv[fp[c]:fp[c + 1]]].sum())
        s = int(dims[sv[sp[c]:sp[c + 1]]].sum())
        out = np.zeros(f * (f + s + 1))
        _check(self.L.b200_get_conditional(self.h, c, _dp(out)))
        return out.reshape(f + s + 1, f).T


class LinearDeviceProblem(DeviceProblem):
    """Device-resident GaussianFactorGraph (b200_linear_create): JacobianFactor groups of any arity.
    solve / get_delta / hessian_diagonal / cliques / conditional / marginal covariances as on
    DeviceProblem; the calls that need Values raise (status B200_INVALID_ARGUMENT)."""

    def __init__(self, ctx: Context, lprob):
        self.ctx, self.prob, self.L = ctx, lprob, ctx.L
        desc, keep = lprob.c_desc()
        h = C.c_void_p()
        _check(self.L.b200_linear_create(ctx.h, C.byref(desc), C.byref(h)))
        del keep
        self.h = h
        self.nval = 0
        self.ndelta = int(self.L.b200_delta_size(h))
        ctx._problems.add(self)

    def update(self, group: int, Ab, sigmas=None):
        """New numbers for one group, same structure (b200_linear_update)."""
        g = self.prob.groups[group]
        Ab = np.ascontiguousarray(Ab, dtype=np.float64)
        assert Ab.size == g.count * g.rows * g.ncols
        sp = None
        if sigmas is not None:
            sigmas = np.ascontiguousarray(sigmas, dtype=np.float64)
            assert sigmas.size == g.count * g.rows
            sp = _dp(sigmas)
        _check(self.L.b200_linear_update(self.h, group, _dp(Ab), sp))

    def update_hessian(self, hgroup: int, info):
        """New augmented information matrices for one HessianFactor group (b200_linear_update_hessian)."""
        g = self.prob.hgroups[hgroup]
        info = np.ascontiguousarray(info, dtype=np.float64)
        assert info.size == g.count * g.ncols * g.ncols
        _check(self.L.b200_linear_update_hessian(self.h, hgroup, _dp(info)))

    def get_jacobians(self, group: int):
        """(count, rows, ncols) whitened [A|b] as stored on the device."""
        g = self.prob.groups[group]
        out = np.zeros(g.count * g.rows * g.ncols)
        _check(self.L.b200_get_jacobians(self.h, group, _dp(out)))
        return out.reshape(g.count, g.ncols, g.rows).transpose(0, 2, 1)


def linear_symbolic(lprob, with_slots=False):
    """Host-only junction tree of a linear problem: (frontal_ptr, frontal_vars, separator_ptr, separator_vars, parent)
    of the reference's cliques; with_slots returns the SUPERNODES instead (what the device eliminates: the cliques
    after relaxed amalgamation) plus (owning supernode per graph position, CSR pointer of the factors' keys, front slot
    of every key)."""
    L = lib()
    desc, keep = lprob.c_desc()
    h = C.c_void_p()
    _check(L.b200_linear_symbolic_create(C.byref(desc), C.byref(h)))
    info = P.CSymbolicInfo()
    L.b200_symbolic_get_info(h, C.byref(info))
    nc, nfl, nsl = ((info.supernodes, info.supernode_frontal_list_len, info.supernode_separator_list_len) if with_slots
                    else (info.ncliques, info.frontal_list_len, info.separator_list_len))
    fp = np.zeros(nc + 1, dtype=np.int64)
    sp = np.zeros(nc + 1, dtype=np.int64)
    fv = np.zeros(max(1, nfl), dtype=np.int64)
    sv = np.zeros(max(1, nsl), dtype=np.int64)
    par = np.zeros(max(1, nc), dtype=np.int64)
    (L.b200_symbolic_get_supernodes if with_slots else L.b200_symbolic_get_cliques)(h, _ip(fp), _ip(fv), _ip(sp), _ip(sv), _ip(par))
    if with_slots:
        arity = np.zeros(lprob.nfactors, dtype=np.int64)
        for g in list(lprob.groups) + list(lprob.hgroups):
            pos = g.graph_index if g.graph_index is not None else g.graph_index0 + np.arange(g.count)
            arity[pos] = g.arity
        fptr = np.concatenate([[0], np.cumsum(arity)]).astype(np.int64)
        clique = np.zeros(max(1, lprob.nfactors), dtype=np.int32)
        slots = np.zeros(max(1, int(fptr[-1])), dtype=np.int32)
        L.b200_symbolic_get_factor_slots(h, clique.ctypes.data_as(C.POINTER(C.c_int32)), slots.ctypes.data_as(C.POINTER(C.c_int32)))
    L.b200_symbolic_destroy(h)
    out = (fp, fv[:nfl], sp, sv[:nsl], par[:nc])
    return out + (clique[:lprob.nfactors], fptr, slots[:int(fptr[-1])]) if with_slots else out
