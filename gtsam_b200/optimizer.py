"""Host-side mirror of the reference's optimizer interface for the hot path.

Same names, argument meaning and error behaviour as
``gtsam::LevenbergMarquardtParams`` (gtsam/nonlinear/LevenbergMarquardtParams.h:49-141),
``gtsam::LevenbergMarquardtOptimizer`` (gtsam/nonlinear/LevenbergMarquardtOptimizer.h:35-120)
``gtsam::GaussNewtonOptimizer`` (gtsam/nonlinear/GaussNewtonOptimizer.h) and
``gtsam::DoglegOptimizer`` (gtsam/nonlinear/DoglegOptimizer.h:33-128), ``gtsam::Marginals``
(gtsam/nonlinear/Marginals.h:31-100), so the
parity tests read like the reference's own (tests/testNonlinearOptimizer.cpp).
All numeric work happens in the C-ABI library on the GPU; this file holds no math.
The C++ subclass shim a GTSAM user links instead is gtsam_b200/shim/.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

from . import problem as P
from .capi import Context, DeviceProblem, IndeterminantLinearSystemException, _check, lib


class LevenbergMarquardtParams:
    """gtsam::LevenbergMarquardtParams (legacy defaults on construction)."""

    def __init__(self):
        self._c = P.CLMParams()
        lib().b200_lm_params_legacy(C.byref(self._c))
        self.iterationHook: Optional[Callable[[int, float, float], None]] = None

    @staticmethod
    def LegacyDefaults() -> "LevenbergMarquardtParams":
        return LevenbergMarquardtParams()

    @staticmethod
    def CeresDefaults() -> "LevenbergMarquardtParams":
        p = LevenbergMarquardtParams()
        lib().b200_lm_params_ceres(C.byref(p._c))
        return p

    # camelCase accessors as in the reference's wrapped API
    maxIterations = property(lambda s: s._c.max_iterations, lambda s, v: setattr(s._c, "max_iterations", int(v)))
    relativeErrorTol = property(lambda s: s._c.relative_error_tol, lambda s, v: setattr(s._c, "relative_error_tol", v))
    absoluteErrorTol = property(lambda s: s._c.absolute_error_tol, lambda s, v: setattr(s._c, "absolute_error_tol", v))
    errorTol = property(lambda s: s._c.error_tol, lambda s, v: setattr(s._c, "error_tol", v))
    lambdaInitial = property(lambda s: s._c.lambda_initial, lambda s, v: setattr(s._c, "lambda_initial", v))
    lambdaFactor = property(lambda s: s._c.lambda_factor, lambda s, v: setattr(s._c, "lambda_factor", v))
    lambdaUpperBound = property(lambda s: s._c.lambda_upper_bound, lambda s, v: setattr(s._c, "lambda_upper_bound", v))
    lambdaLowerBound = property(lambda s: s._c.lambda_lower_bound, lambda s, v: setattr(s._c, "lambda_lower_bound", v))
    minModelFidelity = property(lambda s: s._c.min_model_fidelity, lambda s, v: setattr(s._c, "min_model_fidelity", v))
    diagonalDamping = property(lambda s: bool(s._c.diagonal_damping), lambda s, v: setattr(s._c, "diagonal_damping", int(v)))
    useFixedLambdaFactor = property(lambda s: bool(s._c.use_fixed_lambda_factor),
                                    lambda s, v: setattr(s._c, "use_fixed_lambda_factor", int(v)))
    minDiagonal = property(lambda s: s._c.min_diagonal, lambda s, v: setattr(s._c, "min_diagonal", v))
    maxDiagonal = property(lambda s: s._c.max_diagonal, lambda s, v: setattr(s._c, "max_diagonal", v))


def checkConvergence(relativeErrorTreshold, absoluteErrorTreshold, errorThreshold, currentError, newError) -> bool:
    """gtsam::checkConvergence, gtsam/nonlinear/NonlinearOptimizer.cpp:182-231."""
    if newError <= errorThreshold:
        return True
    absoluteDecrease = currentError - newError
    relativeDecrease = absoluteDecrease / currentError
    return bool((relativeErrorTreshold and relativeDecrease <= relativeErrorTreshold)
                or absoluteDecrease <= absoluteErrorTreshold)


class LevenbergMarquardtOptimizer:
    """Drop-in for gtsam::LevenbergMarquardtOptimizer on a device-resident problem.

    ``graph``/``initialValues``/``ordering`` of the reference's constructor are the
    :class:`Problem` (typed factor tables, packed values, ordering)."""

    def __init__(self, ctx: Context, problem: P.Problem, params: Optional[LevenbergMarquardtParams] = None,
                 device_problem: Optional[DeviceProblem] = None):
        self.params_ = params or LevenbergMarquardtParams()
        self.dp = device_problem or DeviceProblem(ctx, problem)
        self.L = self.dp.L
        h = C.c_void_p()
        _check(self.L.b200_lm_create(self.dp.h, C.byref(self.params_._c), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.b200_lm_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _state(self) -> P.CLMState:
        s = P.CLMState()
        _check(self.L.b200_lm_get_state(self.h, C.byref(s)))
        return s

    def error(self) -> float:
        return self._state().error

    def iterations(self) -> int:
        return self._state().iterations

    def getInnerIterations(self) -> int:
        return self._state().total_inner_iterations

    # `lambda` is a Python keyword; the reference's Python wrapper calls it lambda_ too
    def lambda_(self) -> float:
        return self._state().lambda_

    def values(self):
        return self.dp.get_values()

    def params(self) -> LevenbergMarquardtParams:
        return self.params_

    def iterate(self):
        """LevenbergMarquardtOptimizer::iterate() (.cpp:273-308)."""
        _check(self.L.b200_lm_iterate(self.h))

    def optimize(self):
        """NonlinearOptimizer::defaultOptimize() (NonlinearOptimizer.cpp:62-117) incl. iterationHook."""
        prm = self.params_
        if prm.iterationHook is None:
            _check(self.L.b200_lm_optimize(self.h))
            return self.values()
        currentError = self.error()
        if currentError <= prm.errorTol or self.iterations() >= prm.maxIterations:
            return self.values()
        newError = currentError
        while True:
            currentError = newError
            self.iterate()
            newError = self.error()
            prm.iterationHook(self.iterations(), currentError, newError)
            if not (self.iterations() < prm.maxIterations
                    and not checkConvergence(prm.relativeErrorTol, prm.absoluteErrorTol, prm.errorTol,
                                             currentError, newError)
                    and currentError == currentError and abs(currentError) != float("inf")):
                break
        return self.values()


class GaussNewtonOptimizer:
    """Drop-in for gtsam::GaussNewtonOptimizer (gtsam/nonlinear/GaussNewtonOptimizer.cpp:44-67)."""

    def __init__(self, ctx: Context, problem: P.Problem, device_problem: Optional[DeviceProblem] = None):
        self.dp = device_problem or DeviceProblem(ctx, problem)
        self.error_ = self.dp.error()
        self.iterations_ = 0

    def error(self) -> float:
        return self.error_

    def iterations(self) -> int:
        return self.iterations_

    def values(self):
        return self.dp.get_values()

    def iterate(self):
        rc, e = self.dp.gn_iterate()
        if rc == P.INDETERMINATE:
            raise IndeterminantLinearSystemException(-1)
        self.error_ = e
        self.iterations_ += 1

    def optimize(self, params=None):
        return _default_optimize(self, params or NonlinearOptimizerParams())


class NonlinearOptimizerParams:
    """gtsam::NonlinearOptimizerParams defaults (gtsam/nonlinear/NonlinearOptimizerParams.h:48-53)."""

    def __init__(self):
        self.maxIterations = 100
        self.relativeErrorTol = 1e-5
        self.absoluteErrorTol = 1e-5
        self.errorTol = 0.0
        self.iterationHook: Optional[Callable[[int, float, float], None]] = None


GaussNewtonParams = NonlinearOptimizerParams


def _default_optimize(opt, prm):
    """NonlinearOptimizer::defaultOptimize(), gtsam/nonlinear/NonlinearOptimizer.cpp:62-117."""
    currentError = opt.error()
    if currentError <= prm.errorTol or opt.iterations() >= prm.maxIterations:
        return opt.values()
    newError = currentError
    while True:
        currentError = newError
        opt.iterate()
        newError = opt.error()
        if prm.iterationHook is not None:
            prm.iterationHook(opt.iterations(), currentError, newError)
        if not (opt.iterations() < prm.maxIterations
                and not checkConvergence(prm.relativeErrorTol, prm.absoluteErrorTol, prm.errorTol, currentError, newError)
                and currentError == currentError and abs(currentError) != float("inf")):
            break
    return opt.values()


class DoglegParams(NonlinearOptimizerParams):
    """gtsam::DoglegParams (gtsam/nonlinear/DoglegOptimizer.h:33-58)."""

    def __init__(self):
        super().__init__()
        self.deltaInitial = 1.0


class DoglegOptimizer:
    """Drop-in for gtsam::DoglegOptimizer (gtsam/nonlinear/DoglegOptimizer.cpp:60-121)."""

    def __init__(self, ctx: Context, problem: P.Problem, params: Optional[DoglegParams] = None,
                 device_problem: Optional[DeviceProblem] = None):
        self.params_ = params or DoglegParams()
        self.dp = device_problem or DeviceProblem(ctx, problem)
        self.L = self.dp.L
        h = C.c_void_p()
        _check(self.L.b200_dl_create(self.dp.h, float(self.params_.deltaInitial), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.b200_dl_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _state(self):
        e, d, it = C.c_double(), C.c_double(), C.c_int32()
        _check(self.L.b200_dl_get_state(self.h, C.byref(e), C.byref(d), C.byref(it)))
        return e.value, d.value, it.value

    def error(self) -> float:
        return self._state()[0]

    def getDelta(self) -> float:
        return self._state()[1]

    def iterations(self) -> int:
        return self._state()[2]

    def values(self):
        return self.dp.get_values()

    def params(self) -> DoglegParams:
        return self.params_

    def iterate(self):
        rc = self.L.b200_dl_iterate(self.h)
        if rc == P.INDETERMINATE:
            raise IndeterminantLinearSystemException(-1)
        _check(rc)

    def optimize(self):
        return _default_optimize(self, self.params_)


class Marginals:
    """gtsam::Marginals(graph, solution) for the CHOLESKY factorization (gtsam/nonlinear/Marginals.h:31-100):
    ``marginalCovariance`` / ``marginalInformation`` of one variable at the problem's current values."""

    def __init__(self, ctx: Context, problem: P.Problem, device_problem: Optional[DeviceProblem] = None):
        self.dp = device_problem or DeviceProblem(ctx, problem)

    def marginalCovariance(self, variable: int):
        return self.dp.marginal_covariance(variable)

    def marginalInformation(self, variable: int):
        import numpy as np
        return np.linalg.inv(self.marginalCovariance(variable))

    def jointMarginalCovariance(self, variables) -> "JointMarginal":
        vs = sorted(int(v) for v in variables)
        dims = [P.VAR_DIM[int(self.dp.prob.var_type[v])] for v in vs]
        return JointMarginal(self.dp.joint_marginal_covariance(vs), vs, dims)


class JointMarginal:
    """gtsam::JointMarginal (gtsam/nonlinear/Marginals.h:135-185): blocks addressed by variable, keys sorted."""

    def __init__(self, full, keys, dims):
        self._full, self._keys = full, list(keys)
        self._off = {}
        o = 0
        for k, d in zip(keys, dims):
            self._off[k] = (o, o + d)
            o += d

    def fullMatrix(self):
        return self._full

    def keys(self):
        return list(self._keys)

    def at(self, i: int, j: int):
        (a, b), (c, d) = self._off[int(i)], self._off[int(j)]
        return self._full[a:b, c:d]

    __call__ = at
