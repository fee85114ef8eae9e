"""GaussianFactorGraph level: the multifrontal solve on an ALREADY LINEARIZED graph.

Python mirror of ``b200_linear_desc`` / ``b200_jacobian_group`` (include/gtsam_b200.h) and of the
part of the reference's linear API that sits on the hot path:

* ``JacobianFactor(keys, blocks, b, sigmas)``   gtsam/linear/JacobianFactor.h:93-160
* ``HessianFactor(keys, dims, info)``            gtsam/linear/HessianFactor.h:99-110
* ``GaussianFactorGraph``                        gtsam/linear/GaussianFactorGraph.h:73-404
  ``.add / .push_back / .size / .keys``, ``.optimize(ordering)`` (GaussianFactorGraph.cpp:316-319, the
  multifrontal Cholesky path), ``.eliminateMultifrontal(ordering)``, ``.hessianDiagonal()`` (:279-287),
  ``.gradientAtZero()`` (:369-378), ``.error(x)`` (:71-78)
* ``VectorValues`` is a plain ``dict`` key -> 1-D array (gtsam/linear/VectorValues.h:77-78).

Factors of any arity and any block widths.  The numbers live in ``LinearProblem`` (flat groups of
same-shape factors, what crosses the C-ABI); this module holds no math: ``optimize`` packs, calls
``b200_linear_create`` + ``b200_solve`` + ``b200_get_delta`` and unpacks.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import problem as P

JACOBIAN_MAX_ARITY = 8


class CJacobianGroup(C.Structure):
    _fields_ = [("rows", C.c_int32), ("arity", C.c_int32), ("dims", C.POINTER(C.c_int32)), ("count", C.c_int64),
                ("graph_index0", C.c_int64), ("graph_index", C.POINTER(C.c_int64)), ("keys", C.POINTER(C.c_int64)),
                ("Ab", C.POINTER(C.c_double)), ("sigmas", C.POINTER(C.c_double))]


class CHessianGroup(C.Structure):
    _fields_ = [("arity", C.c_int32), ("dims", C.POINTER(C.c_int32)), ("count", C.c_int64), ("graph_index0", C.c_int64),
                ("graph_index", C.POINTER(C.c_int64)), ("keys", C.POINTER(C.c_int64)), ("info", C.POINTER(C.c_double))]


class CLinearDesc(C.Structure):
    _fields_ = [("nvars", C.c_int64), ("var_dim", C.POINTER(C.c_int32)), ("ordering", C.POINTER(C.c_int64)),
                ("ngroups", C.c_int64), ("groups", C.POINTER(CJacobianGroup)),
                ("nhgroups", C.c_int64), ("hgroups", C.POINTER(CHessianGroup))]


@dataclass
class JacobianGroup:
    """A run of JacobianFactors of one shape.  ``Ab[f, c, r]`` = entry (r, c) of factor f's [A1 .. Ak b]
    (each factor column-major, the layout of JacobianFactor::matrixObject())."""
    rows: int
    dims: Sequence[int]
    keys: np.ndarray                      # (count, arity) int64 variable ids
    Ab: np.ndarray                        # (count, sum(dims)+1, rows) float64
    sigmas: Optional[np.ndarray] = None   # (count, rows) Diagonal sigmas, None = unit
    graph_index0: int = -1
    graph_index: Optional[np.ndarray] = None

    def __post_init__(self):
        self.dims = np.ascontiguousarray(self.dims, dtype=np.int32)
        self.keys = np.ascontiguousarray(self.keys, dtype=np.int64).reshape(-1, self.arity)
        self.Ab = np.ascontiguousarray(self.Ab, dtype=np.float64).reshape(self.count, self.ncols, self.rows)
        if self.sigmas is not None:
            self.sigmas = np.ascontiguousarray(self.sigmas, dtype=np.float64).reshape(self.count, self.rows)
        if self.graph_index is not None:
            self.graph_index = np.ascontiguousarray(self.graph_index, dtype=np.int64)
            assert self.graph_index.size == self.count

    @property
    def arity(self) -> int:
        return int(self.dims.size)

    @property
    def ncols(self) -> int:
        return int(self.dims.sum()) + 1

    @property
    def count(self) -> int:
        return int(self.keys.shape[0])

    def whitened(self) -> np.ndarray:
        """(count, rows, ncols) whitened [A|b] — for tests; the library whitens on the device."""
        M = self.Ab.transpose(0, 2, 1)
        return M if self.sigmas is None else M * (1.0 / self.sigmas)[:, :, None]


@dataclass
class HessianGroup:
    """A run of HessianFactors with the same block widths.  ``info[f, c, r]`` = entry (r, c) of factor f's augmented
    information matrix [G g; g' f] (HessianFactor::info(); the upper triangle is what the library reads)."""
    dims: Sequence[int]
    keys: np.ndarray                      # (count, arity) int64 variable ids
    info: np.ndarray                      # (count, N+1, N+1) float64, N = sum(dims)
    graph_index0: int = -1
    graph_index: Optional[np.ndarray] = None

    def __post_init__(self):
        self.dims = np.ascontiguousarray(self.dims, dtype=np.int32)
        self.keys = np.ascontiguousarray(self.keys, dtype=np.int64).reshape(-1, self.arity)
        self.info = np.ascontiguousarray(self.info, dtype=np.float64).reshape(self.count, self.ncols, self.ncols)
        if self.graph_index is not None:
            self.graph_index = np.ascontiguousarray(self.graph_index, dtype=np.int64)
            assert self.graph_index.size == self.count

    @property
    def arity(self) -> int:
        return int(self.dims.size)

    @property
    def ncols(self) -> int:
        return int(self.dims.sum()) + 1

    rows = ncols

    @property
    def count(self) -> int:
        return int(self.keys.shape[0])


@dataclass
class LinearProblem:
    var_dim: np.ndarray        # (nvars,) int32 tangent dimensions
    ordering: np.ndarray       # (nvars,) int64 elimination order
    groups: List[JacobianGroup] = field(default_factory=list)
    hgroups: List[HessianGroup] = field(default_factory=list)   # HessianFactor groups (graph positions after / among the Jacobian ones)
    name: str = ""
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        self.var_dim = np.ascontiguousarray(self.var_dim, dtype=np.int32)
        self.ordering = np.ascontiguousarray(self.ordering, dtype=np.int64)
        nxt = 0
        for g in list(self.groups) + list(self.hgroups):
            if g.graph_index is not None:
                continue
            if g.graph_index0 < 0:
                g.graph_index0 = nxt
            nxt = g.graph_index0 + g.count

    @property
    def nvars(self) -> int:
        return int(self.var_dim.size)

    @property
    def nfactors(self) -> int:
        return sum(g.count for g in self.groups) + sum(g.count for g in self.hgroups)

    @property
    def var_dims(self) -> np.ndarray:
        return self.var_dim

    def dof_offsets(self) -> np.ndarray:
        return np.concatenate([[0], np.cumsum(self.var_dim)]).astype(np.int64)

    def c_desc(self):
        garr = (CJacobianGroup * max(1, len(self.groups)))()
        for i, g in enumerate(self.groups):
            garr[i].rows, garr[i].arity, garr[i].count = g.rows, g.arity, g.count
            garr[i].dims = P._ptr(g.dims, C.c_int32)
            garr[i].graph_index0 = g.graph_index0
            garr[i].graph_index = P._ptr(g.graph_index, C.c_int64)
            garr[i].keys = P._ptr(g.keys, C.c_int64)
            garr[i].Ab = P._ptr(g.Ab, C.c_double)
            garr[i].sigmas = P._ptr(g.sigmas, C.c_double)
        harr = (CHessianGroup * max(1, len(self.hgroups)))()
        for i, g in enumerate(self.hgroups):
            harr[i].arity, harr[i].count = g.arity, g.count
            harr[i].dims = P._ptr(g.dims, C.c_int32)
            harr[i].graph_index0 = g.graph_index0
            harr[i].graph_index = P._ptr(g.graph_index, C.c_int64)
            harr[i].keys = P._ptr(g.keys, C.c_int64)
            harr[i].info = P._ptr(g.info, C.c_double)
        d = CLinearDesc()
        d.nvars = self.nvars
        d.var_dim = P._ptr(self.var_dim, C.c_int32)
        d.ordering = P._ptr(self.ordering, C.c_int64)
        d.ngroups = len(self.groups)
        d.groups = garr
        d.nhgroups = len(self.hgroups)
        d.hgroups = harr
        return d, (garr, harr, self)

    # -- file exchange with oracle/ref_harness.cpp (oracle/linear_io.hpp) ------------------
    MAGIC = b"B200LIN1"

    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(self.MAGIC)
            f.write(struct.pack("<q", self.nvars))
            f.write(self.var_dim.tobytes())
            f.write(self.ordering.tobytes())
            f.write(struct.pack("<q", len(self.groups) + len(self.hgroups)))
            for g in self.groups:
                f.write(struct.pack("<ii", g.rows, g.arity))
                f.write(g.dims.tobytes())
                flags = (1 if g.sigmas is not None else 0) | (4 if g.graph_index is not None else 0)
                f.write(struct.pack("<qqi", g.count, g.graph_index0, flags))
                f.write(g.keys.tobytes())
                f.write(g.Ab.tobytes())
                if g.sigmas is not None:
                    f.write(g.sigmas.tobytes())
                if g.graph_index is not None:
                    f.write(g.graph_index.tobytes())
            for g in self.hgroups:
                f.write(struct.pack("<ii", g.ncols, g.arity))
                f.write(g.dims.tobytes())
                f.write(struct.pack("<qqi", g.count, g.graph_index0, 8 | (4 if g.graph_index is not None else 0)))
                f.write(g.keys.tobytes())
                f.write(g.info.tobytes())
                if g.graph_index is not None:
                    f.write(g.graph_index.tobytes())

    @classmethod
    def load(cls, path: str) -> "LinearProblem":
        with open(path, "rb") as f:
            buf = f.read()
        assert buf[:8] == cls.MAGIC
        o = 8

        def rd(fmt):
            nonlocal o
            v = struct.unpack_from(fmt, buf, o)
            o += struct.calcsize(fmt)
            return v

        def arr(dtype, n):
            nonlocal o
            a = np.frombuffer(buf, dtype=dtype, count=n, offset=o).copy()
            o += a.nbytes
            return a

        (nv,) = rd("<q")
        vd = arr(np.int32, nv)
        order = arr(np.int64, nv)
        (ng,) = rd("<q")
        groups, hgroups = [], []
        for _ in range(ng):
            rows, ar = rd("<ii")
            dims = arr(np.int32, ar)
            cnt, gi0, flags = rd("<qqi")
            nc = int(dims.sum()) + 1
            keys = arr(np.int64, cnt * ar)
            Ab = arr(np.float64, cnt * rows * nc)
            sig = arr(np.float64, cnt * rows) if flags & 1 else None
            gidx = arr(np.int64, cnt) if flags & 4 else None
            if flags & 8:
                hgroups.append(HessianGroup(dims, keys, Ab, gi0, gidx))
            else:
                groups.append(JacobianGroup(rows, dims, keys, Ab, sig, gi0, gidx))
        return cls(vd, order, groups, hgroups)


# ---- the reference's object-level API, mirrored ---------------------------------------------
class JacobianFactor:
    """|A1 x1 + .. + Ak xk - b|^2_Sigma with an optional Diagonal model (sigmas)."""

    def __init__(self, keys, blocks, b, sigmas=None):
        self._keys = [int(k) for k in keys]
        self.blocks = [np.atleast_2d(np.asarray(A, dtype=np.float64)) for A in blocks]
        self.b = np.asarray(b, dtype=np.float64).ravel()
        self.sigmas = None if sigmas is None else np.asarray(sigmas, dtype=np.float64).ravel()
        assert len(self._keys) == len(self.blocks) and len(set(self._keys)) == len(self._keys)
        assert all(A.shape[0] == self.b.size for A in self.blocks)

    def keys(self):
        return list(self._keys)

    def rows(self) -> int:
        return int(self.b.size)

    def getDim(self, i: int) -> int:
        return int(self.blocks[i].shape[1])

    def augmentedJacobianUnweighted(self) -> np.ndarray:
        return np.concatenate(self.blocks + [self.b[:, None]], axis=1)


class HessianFactor:
    """0.5 (f - 2 x'g + x'G x) over `keys` (gtsam/linear/HessianFactor.h:99-110): `info` is the symmetric augmented
    information matrix [G g; g' f], `dims` the block width of every key."""

    def __init__(self, keys, dims, info):
        self._keys = [int(k) for k in keys]
        self.dims = [int(d) for d in dims]
        self.info = np.asarray(info, dtype=np.float64)
        n1 = sum(self.dims) + 1
        assert self.info.shape == (n1, n1) and len(self.dims) == len(self._keys)

    def keys(self):
        return list(self._keys)

    def getDim(self, i: int) -> int:
        return self.dims[i]


VectorValues = Dict[int, np.ndarray]


class GaussianFactorGraph:
    def __init__(self, factors=()):
        self.factors: list = list(factors)   # JacobianFactor | HessianFactor

    def add(self, *args):
        """add(factor) or add(keys, blocks, b[, sigmas]) like the reference's overloads."""
        self.factors.append(args[0] if len(args) == 1 else JacobianFactor(*args))

    push_back = add

    def size(self) -> int:
        return len(self.factors)

    def keys(self):
        return sorted({k for f in self.factors for k in f.keys()})

    def to_problem(self, ordering=None):
        """(LinearProblem, ids): keys -> dense ids in ascending key order (as the C++ shim does); factors
        grouped by shape with explicit graph positions.  ordering = keys in elimination order (default:
        ascending keys)."""
        keys = self.keys()
        ids = {k: i for i, k in enumerate(keys)}
        dim = {}
        for f in self.factors:
            for i, k in enumerate(f.keys()):
                if dim.setdefault(k, f.getDim(i)) != f.getDim(i):
                    raise ValueError(f"variable {k} appears with two different dimensions")
        order = np.array([ids[k] for k in (keys if ordering is None else ordering)], dtype=np.int64)
        buckets: Dict[tuple, dict] = {}
        hbuckets: Dict[tuple, dict] = {}
        for pos, f in enumerate(self.factors):
            if isinstance(f, HessianFactor):
                b = hbuckets.setdefault(tuple(f.dims), dict(keys=[], info=[], pos=[]))
                b["keys"].append([ids[k] for k in f.keys()])
                b["info"].append(np.triu(f.info) + np.triu(f.info, 1).T)   # symmetric; stored column-major == row-major
                b["pos"].append(pos)
                continue
            sig = (f.rows(), f.sigmas is not None) + tuple(f.getDim(i) for i in range(len(f.keys())))
            b = buckets.setdefault(sig, dict(keys=[], Ab=[], sig=[], pos=[]))
            b["keys"].append([ids[k] for k in f.keys()])
            b["Ab"].append(f.augmentedJacobianUnweighted().T)     # (ncols, rows): column-major block
            if f.sigmas is not None:
                b["sig"].append(f.sigmas)
            b["pos"].append(pos)
        groups = [JacobianGroup(sig[0], sig[2:], np.array(b["keys"]), np.array(b["Ab"]),
                                np.array(b["sig"]) if sig[1] else None, graph_index=np.array(b["pos"]))
                  for sig, b in buckets.items()]
        hgroups = [HessianGroup(dims, np.array(b["keys"]), np.array(b["info"]), graph_index=np.array(b["pos"]))
                   for dims, b in hbuckets.items()]
        return LinearProblem(np.array([dim[k] for k in keys], dtype=np.int32), order, groups, hgroups), ids

    def hessianDiagonal(self, ctx=None) -> VectorValues:
        """GaussianFactorGraph::hessianDiagonal() on the device."""
        from . import capi
        own = ctx is None
        ctx = ctx or capi.Context(0)
        try:
            lp, ids = self.to_problem(None)
            dev = capi.LinearDeviceProblem(ctx, lp)
            h, off = dev.hessian_diagonal(), lp.dof_offsets()
            dev.close()
            return {k: h[off[i]:off[i + 1]].copy() for k, i in ids.items()}
        finally:
            if own:
                ctx.close()

    def error(self, x: VectorValues, ctx=None) -> float:
        """GaussianFactorGraph::error(x) (GaussianFactorGraph.cpp:71-78) on the device."""
        from . import capi
        own = ctx is None
        ctx = ctx or capi.Context(0)
        try:
            lp, ids = self.to_problem(None)
            off = lp.dof_offsets()
            v = np.zeros(off[-1])
            for k, i in ids.items():
                v[off[i]:off[i + 1]] = np.asarray(x[k], dtype=np.float64)
            dev = capi.LinearDeviceProblem(ctx, lp)
            e = dev.linear_graph_error(v)
            dev.close()
            return e
        finally:
            if own:
                ctx.close()

    def gradientAtZero(self, ctx=None) -> VectorValues:
        """GaussianFactorGraph::gradientAtZero() (GaussianFactorGraph.cpp:369-378) on the device: -A'b per key."""
        from . import capi
        own = ctx is None
        ctx = ctx or capi.Context(0)
        try:
            lp, ids = self.to_problem(None)
            dev = capi.LinearDeviceProblem(ctx, lp)
            g, off = dev.gradient_at_zero(), lp.dof_offsets()
            dev.close()
            return {k: g[off[i]:off[i + 1]].copy() for k, i in ids.items()}
        finally:
            if own:
                ctx.close()

    def eliminateMultifrontal(self, ordering=None, ctx=None):
        """GaussianFactorGraph::eliminateMultifrontal(ordering, EliminatePreferCholesky) on the device: the Bayes tree as a
        list of cliques in elimination order, each ``(frontal keys, separator keys, parent index or -1, [R S d])`` with
        ``[R S d]`` an f x (f+s+1) array (columns: frontals, separators in ascending key order, rhs)."""
        from . import capi
        own = ctx is None
        ctx = ctx or capi.Context(0)
        try:
            lp, ids = self.to_problem(ordering)
            key_of = {i: k for k, i in ids.items()}
            dev = capi.LinearDeviceProblem(ctx, lp)
            st, _, _, fv = dev.solve(0.0)
            if st == P.INDETERMINATE:
                raise capi.IndeterminantLinearSystemException(key_of.get(int(fv), -1))
            fp, fvars, sp, svars, par = dev.cliques()
            out = [([key_of[int(v)] for v in fvars[fp[c]:fp[c + 1]]], [key_of[int(v)] for v in svars[sp[c]:sp[c + 1]]], int(par[c]),
                    dev.conditional(c)) for c in range(len(par))]
            dev.close()
            return out
        finally:
            if own:
                ctx.close()

    def optimize(self, ordering=None, ctx=None) -> VectorValues:
        """GaussianFactorGraph::optimize(ordering, EliminatePreferCholesky) on the device."""
        from . import capi
        own = ctx is None
        ctx = ctx or capi.Context(0)
        try:
            lp, ids = self.to_problem(ordering)
            dev = capi.LinearDeviceProblem(ctx, lp)
            st, _, _, fv = dev.solve(0.0)
            if st == P.INDETERMINATE:
                raise capi.IndeterminantLinearSystemException(self.keys()[fv] if fv >= 0 else -1)
            delta, off = dev.get_delta(), lp.dof_offsets()
            dev.close()
            return {k: delta[off[i]:off[i + 1]].copy() for k, i in ids.items()}
        finally:
            if own:
                ctx.close()
