"""Host-side problem container: the flat description that crosses the C-ABI.

This is the Python mirror of ``b200_problem_desc`` / ``b200_factor_group``
(include/gtsam_b200.h).  It stands where the reference has a
``NonlinearFactorGraph`` + ``Values`` + ``Ordering`` triple
(gtsam/nonlinear/NonlinearFactorGraph.h, gtsam/nonlinear/Values.h:74-79,
gtsam/inference/Ordering.h:217-236): typed factor tables (SoA) instead of a
vector of shared_ptr<NonlinearFactor>, a packed value array instead of a
std::map<Key, Value>.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# ---- enums (include/gtsam_b200.h) -------------------------------------------
VAR_POSE3, VAR_POINT3, VAR_CAM_BUNDLER, VAR_POSE2 = 0, 1, 2, 3
VAR_STORAGE = (12, 3, 17, 3)
VAR_DIM = (6, 3, 9, 3)

(FACTOR_BETWEEN_POSE3, FACTOR_PRIOR_POSE3, FACTOR_PRIOR_POINT3, FACTOR_PROJECTION_CAL3S2,
 FACTOR_SFM_BUNDLER, FACTOR_PRIOR_CAM_BUNDLER, FACTOR_BETWEEN_POSE2, FACTOR_PRIOR_POSE2) = range(8)
FACTOR_ARITY = (2, 1, 1, 2, 2, 1, 2, 1)
FACTOR_MEAS = (12, 12, 3, 2, 2, 17, 3, 3)
FACTOR_DIM = (6, 6, 3, 2, 2, 9, 3, 3)
FACTOR_VAR_TYPES = ((0, 0), (0,), (1,), (0, 1), (2, 1), (2,), (3, 3), (3,))

NOISE_UNIT, NOISE_ISOTROPIC, NOISE_DIAGONAL, NOISE_GAUSSIAN = range(4)
ROBUST_NONE, ROBUST_HUBER, ROBUST_CAUCHY, ROBUST_TUKEY, ROBUST_FAIR = range(5)

(OK, INDETERMINATE, UNSUPPORTED_FACTOR, UNSUPPORTED_NOISE, INVALID_ARGUMENT, CUDA_ERROR,
 NCCL_ERROR, NO_DEVICE) = range(8)


def noise_payload(kind: int, d: int) -> int:
    return (0, 1, d, d * d)[kind] if 0 <= kind < 4 else 0   # unknown kinds are rejected by the library


def factor_ncols(ftype: int) -> int:
    """Columns of the whitened per-factor block [A1 A2 b]."""
    return sum(VAR_DIM[t] for t in FACTOR_VAR_TYPES[ftype]) + 1


# ---- ctypes mirrors -----------------------------------------------------------
class CFactorGroup(C.Structure):
    _fields_ = [("type", C.c_int32), ("noise_kind", C.c_int32), ("noise_per_factor", C.c_int32),
                ("robust_kind", C.c_int32), ("count", C.c_int64), ("graph_index0", C.c_int64),
                ("keys", C.POINTER(C.c_int64)), ("meas", C.POINTER(C.c_double)),
                ("noise", C.POINTER(C.c_double)), ("cal_index", C.POINTER(C.c_int32)),
                ("body_P_sensor", C.POINTER(C.c_double)), ("robust_param", C.c_double),
                ("graph_index", C.POINTER(C.c_int64))]


class CProblemDesc(C.Structure):
    _fields_ = [("nvars", C.c_int64), ("var_type", C.POINTER(C.c_int32)),
                ("values", C.POINTER(C.c_double)), ("ordering", C.POINTER(C.c_int64)),
                ("ncal", C.c_int64), ("cal", C.POINTER(C.c_double)), ("ngroups", C.c_int64),
                ("groups", C.POINTER(CFactorGroup))]


class CLMParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("relative_error_tol", C.c_double),
                ("absolute_error_tol", C.c_double), ("error_tol", C.c_double),
                ("lambda_initial", C.c_double), ("lambda_factor", C.c_double),
                ("lambda_upper_bound", C.c_double), ("lambda_lower_bound", C.c_double),
                ("min_model_fidelity", C.c_double), ("diagonal_damping", C.c_int32),
                ("use_fixed_lambda_factor", C.c_int32), ("min_diagonal", C.c_double),
                ("max_diagonal", C.c_double)]


class CLMState(C.Structure):
    _fields_ = [("error", C.c_double), ("lambda_", C.c_double), ("current_factor", C.c_double),
                ("iterations", C.c_int32), ("total_inner_iterations", C.c_int32)]


class CSymbolicInfo(C.Structure):
    _fields_ = [("ncliques", C.c_int64), ("nlevels", C.c_int64), ("total_dim", C.c_int64),
                ("max_frontal_dim", C.c_int64), ("max_separator_dim", C.c_int64),
                ("frontal_list_len", C.c_int64), ("separator_list_len", C.c_int64),
                ("factor_flops", C.c_double), ("front_bytes", C.c_int64),
                ("supernodes", C.c_int64), ("supernode_levels", C.c_int64), ("supernode_max_frontal_dim", C.c_int64),
                ("supernode_max_separator_dim", C.c_int64), ("supernode_frontal_list_len", C.c_int64),
                ("supernode_separator_list_len", C.c_int64), ("supernode_flops", C.c_double)]


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None or a.size == 0:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))


@dataclass
class FactorGroup:
    """A homogeneous run of factors (same type and noise kind)."""
    type: int
    keys: np.ndarray                 # (count, arity) int64 variable ids
    meas: np.ndarray                 # (count, meas_size) float64
    noise_kind: int = NOISE_UNIT
    noise: Optional[np.ndarray] = None  # shared (payload,) or per-factor (count, payload)
    cal_index: Optional[np.ndarray] = None
    graph_index0: int = -1
    body_P_sensor: Optional[np.ndarray] = None   # (12,) Pose3 shared by the group (projection factors)
    robust_kind: int = 0                          # ROBUST_*: noiseModel::Robust around the noise model
    robust_param: float = 0.0
    graph_index: Optional[np.ndarray] = None      # explicit graph positions (non-consecutive factors)

    def __post_init__(self):
        ar, ms = FACTOR_ARITY[self.type], FACTOR_MEAS[self.type]
        self.keys = np.ascontiguousarray(self.keys, dtype=np.int64).reshape(-1, ar)
        self.meas = np.ascontiguousarray(self.meas, dtype=np.float64).reshape(-1, ms)
        assert self.keys.shape[0] == self.meas.shape[0]
        pay = noise_payload(self.noise_kind, FACTOR_DIM[self.type])
        if pay == 0:
            self.noise = np.zeros(0)
        else:
            self.noise = np.ascontiguousarray(self.noise, dtype=np.float64)
            assert self.noise.size in (pay, pay * self.count), "noise payload size"
        if self.cal_index is not None:
            self.cal_index = np.ascontiguousarray(self.cal_index, dtype=np.int32)
        if self.body_P_sensor is not None:
            self.body_P_sensor = np.ascontiguousarray(self.body_P_sensor, dtype=np.float64).reshape(12)
        if self.graph_index is not None:
            self.graph_index = np.ascontiguousarray(self.graph_index, dtype=np.int64)
            assert self.graph_index.size == self.keys.shape[0]

    @property
    def count(self) -> int:
        return int(self.keys.shape[0])

    @property
    def noise_per_factor(self) -> int:
        pay = noise_payload(self.noise_kind, FACTOR_DIM[self.type])
        return int(pay > 0 and self.noise.size == pay * self.count and self.count > 1)


@dataclass
class Problem:
    var_type: np.ndarray             # (nvars,) int32
    values: np.ndarray               # packed float64
    ordering: np.ndarray             # (nvars,) int64 elimination order
    groups: List[FactorGroup] = field(default_factory=list)
    cal: np.ndarray = field(default_factory=lambda: np.zeros((0, 5)))
    name: str = ""
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        self.var_type = np.ascontiguousarray(self.var_type, dtype=np.int32)
        self.values = np.ascontiguousarray(self.values, dtype=np.float64).ravel()
        self.ordering = np.ascontiguousarray(self.ordering, dtype=np.int64)
        self.cal = np.ascontiguousarray(self.cal, dtype=np.float64).reshape(-1, 5)
        nxt = 0
        for g in self.groups:  # resolve graph positions
            if g.graph_index is not None:
                continue
            if g.graph_index0 < 0:
                g.graph_index0 = nxt
            nxt = g.graph_index0 + g.count

    # -- sizes --------------------------------------------------------------------
    @property
    def nvars(self) -> int:
        return int(self.var_type.size)

    @property
    def nfactors(self) -> int:
        return sum(g.count for g in self.groups)

    @property
    def var_dims(self) -> np.ndarray:
        """Tangent dimension of every variable."""
        return np.asarray(VAR_DIM)[self.var_type]

    def val_offsets(self) -> np.ndarray:
        st = np.asarray(VAR_STORAGE)[self.var_type]
        return np.concatenate([[0], np.cumsum(st)]).astype(np.int64)

    def dof_offsets(self) -> np.ndarray:
        st = np.asarray(VAR_DIM)[self.var_type]
        return np.concatenate([[0], np.cumsum(st)]).astype(np.int64)

    def linearize_bytes(self, jac_elem_bytes: int = 8) -> int:
        """Algorithmic bytes of one linearize pass, materialised-[A|b] definition
        (SURVEY.md §8d): per factor = measurement + 2 x int32 ids (+ noise payload
        when per-factor) in, whitened [A1 A2 b] out; plus one pass over Values.
        jac_elem_bytes = 4 in the FP32-storage mode (b200_set_jacobian_precision)."""
        total = int(self.values.size) * 8
        for g in self.groups:
            d = FACTOR_DIM[g.type]
            per = FACTOR_MEAS[g.type] * 8 + 4 * FACTOR_ARITY[g.type] + d * factor_ncols(g.type) * jac_elem_bytes
            if g.noise_per_factor:
                per += noise_payload(g.noise_kind, d) * 8
            total += per * g.count
        return total

    # -- C view -------------------------------------------------------------------
    def c_desc(self):
        """Returns (CProblemDesc, keepalive)."""
        garr = (CFactorGroup * max(1, len(self.groups)))()
        for i, g in enumerate(self.groups):
            garr[i].type = g.type
            garr[i].noise_kind = g.noise_kind
            garr[i].noise_per_factor = g.noise_per_factor
            garr[i].count = g.count
            garr[i].graph_index0 = g.graph_index0
            garr[i].keys = _ptr(g.keys, C.c_int64)
            garr[i].meas = _ptr(g.meas, C.c_double)
            garr[i].noise = _ptr(g.noise, C.c_double)
            garr[i].cal_index = _ptr(g.cal_index, C.c_int32)
            garr[i].body_P_sensor = _ptr(g.body_P_sensor, C.c_double)
            garr[i].robust_kind = g.robust_kind
            garr[i].robust_param = g.robust_param
            garr[i].graph_index = _ptr(g.graph_index, C.c_int64)
        d = CProblemDesc()
        d.nvars = self.nvars
        d.var_type = _ptr(self.var_type, C.c_int32)
        d.values = _ptr(self.values, C.c_double)
        d.ordering = _ptr(self.ordering, C.c_int64)
        d.ncal = self.cal.shape[0]
        d.cal = _ptr(self.cal, C.c_double)
        d.ngroups = len(self.groups)
        d.groups = garr
        return d, (garr, self)

    # -- file exchange with oracle/ref_harness.cpp ---------------------------------
    MAGIC = b"B200PRB1"

    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(self.MAGIC)
            f.write(struct.pack("<q", self.nvars))
            f.write(self.var_type.tobytes())
            f.write(struct.pack("<q", self.values.size))
            f.write(self.values.tobytes())
            f.write(self.ordering.tobytes())
            f.write(struct.pack("<q", self.cal.shape[0]))
            f.write(self.cal.tobytes())
            f.write(struct.pack("<q", len(self.groups)))
            for g in self.groups:
                f.write(struct.pack("<iiiiqq", g.type, g.noise_kind, g.noise_per_factor,
                                    int(g.cal_index is not None) | (2 if g.body_P_sensor is not None else 0) | (4 if g.graph_index is not None else 0) | (g.robust_kind << 8),
                                    g.count, g.graph_index0))
                f.write(g.keys.tobytes())
                f.write(g.meas.tobytes())
                f.write(struct.pack("<q", g.noise.size))
                f.write(g.noise.tobytes())
                if g.cal_index is not None:
                    f.write(g.cal_index.tobytes())
                if g.body_P_sensor is not None:
                    f.write(g.body_P_sensor.tobytes())
                if g.robust_kind:
                    f.write(struct.pack("<d", g.robust_param))
                if g.graph_index is not None:
                    f.write(g.graph_index.tobytes())

    @classmethod
    def load(cls, path: str) -> "Problem":
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
        vt = arr(np.int32, nv)
        (nval,) = rd("<q")
        vals = arr(np.float64, nval)
        order = arr(np.int64, nv)
        (ncal,) = rd("<q")
        cal = arr(np.float64, ncal * 5)
        (ng,) = rd("<q")
        groups = []
        for _ in range(ng):
            t, nk, npf, hc, cnt, gi0 = rd("<iiiiqq")
            keys = arr(np.int64, cnt * FACTOR_ARITY[t])
            meas = arr(np.float64, cnt * FACTOR_MEAS[t])
            (nn,) = rd("<q")
            noise = arr(np.float64, nn)
            ci = arr(np.int32, cnt) if hc & 1 else None
            body = arr(np.float64, 12) if hc & 2 else None
            rk = (hc >> 8) & 0xff
            (rp,) = rd("<d") if rk else (0.0,)
            gidx = arr(np.int64, cnt) if hc & 4 else None
            groups.append(FactorGroup(t, keys, meas, nk, noise, ci, gi0, body, rk, rp, gidx))
        return cls(vt, vals, order, groups, cal)
