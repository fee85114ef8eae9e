"""Graduated non-convexity on top of the device LM: host-side mirror of ``gtsam::GncOptimizer``
(gtsam/nonlinear/GncOptimizer.h:43-472) and ``gtsam::GncParams`` (gtsam/nonlinear/GncParams.h:40-175).

GNC is control logic around a base optimizer: weights per factor, a weighted copy of the graph
(``makeWeightedGraph``: information scaled by the weight, GncOptimizer.h:391-411), the base optimizer run on
it from the SAME initial values every iteration (:214-217), the mu schedule and the three convergence tests.
All of that is host logic here; the numeric work (factor errors, the weighted LM solves) goes through a
backend.  :class:`DeviceBackend` runs it on the GPU through the existing C-ABI (a new device problem per
weighted graph: the per-factor noise payload carries the weights); the CPU tests drive the same host logic
with an oracle-backed backend against traces of the unmodified reference.
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence

import numpy as np

from . import problem as P
from .capi import Context, DeviceProblem
from .optimizer import LevenbergMarquardtOptimizer, LevenbergMarquardtParams

GM, TLS = "GM", "TLS"


def chi2inv(alpha: float, dofs: int) -> float:
    """Quantile of the chi-squared distribution (GncOptimizer.h:38-40 -> internal::chi_squared_quantile)."""
    from scipy.stats import chi2
    return float(chi2.ppf(alpha, dofs))


class GncParams:
    """gtsam::GncParams<LevenbergMarquardtParams> with the reference's defaults (GncParams.h:66-80)."""

    def __init__(self, baseOptimizerParams: Optional[LevenbergMarquardtParams] = None):
        self.baseOptimizerParams = baseOptimizerParams or LevenbergMarquardtParams()
        self.lossType = TLS
        self.maxIterations = 100
        self.muStep = 1.4
        self.relativeCostTol = 1e-5
        self.weightsTol = 1e-4
        self.knownInliers: List[int] = []
        self.knownOutliers: List[int] = []

    def setLossType(self, t):
        self.lossType = t

    def setMuStep(self, s):
        self.muStep = s

    def setRelativeCostTol(self, v):
        self.relativeCostTol = v

    def setWeightsTol(self, v):
        self.weightsTol = v

    def setKnownInliers(self, idx: Sequence[int]):
        self.knownInliers = sorted(int(i) for i in idx)

    def setKnownOutliers(self, idx: Sequence[int]):
        self.knownOutliers = sorted(int(i) for i in idx)


def graph_positions(g: P.FactorGroup) -> np.ndarray:
    return g.graph_index if g.graph_index is not None else g.graph_index0 + np.arange(g.count, dtype=np.int64)


def strip_robust(prob: P.Problem) -> P.Problem:
    """GncOptimizer's constructor drops a noiseModel::Robust wrapper and keeps its Gaussian part (:70-78)."""
    out = copy.copy(prob)
    out.groups = []
    for g in prob.groups:
        h = copy.copy(g)
        h.robust_kind, h.robust_param = 0, 0.0
        out.groups.append(h)
    return out


def weighted_problem(prob: P.Problem, weights: np.ndarray, values: Optional[np.ndarray] = None) -> P.Problem:
    """makeWeightedGraph (GncOptimizer.h:391-411): information <- w * information, i.e. sqrt-information
    <- sqrt(w) * R; expressed in this library's noise payloads (sigma / sigmas / R) per factor.  A zero
    weight gives an infinite sigma: the factor whitens to nothing, as Gaussian::Information(0) does."""
    out = copy.copy(prob)
    out.groups = []
    if values is not None:
        out.values = np.ascontiguousarray(values, dtype=np.float64).ravel()
    for g in prob.groups:
        d = P.FACTOR_DIM[g.type]
        w = np.asarray(weights, dtype=np.float64)[graph_positions(g)]
        sw = np.sqrt(w)
        h = copy.copy(g)
        with np.errstate(divide="ignore"):
            inv = 1.0 / sw                         # inf where w == 0
        if g.noise_kind == P.NOISE_UNIT:
            h.noise_kind, h.noise = P.NOISE_ISOTROPIC, inv.reshape(-1, 1).copy()
        elif g.noise_kind == P.NOISE_ISOTROPIC:
            h.noise = (np.broadcast_to(g.noise.reshape(-1, 1), (g.count, 1)) * inv[:, None]).copy()
        elif g.noise_kind == P.NOISE_DIAGONAL:
            h.noise = (np.broadcast_to(g.noise.reshape(-1, d), (g.count, d)) * inv[:, None]).copy()
        else:
            h.noise = (np.broadcast_to(g.noise.reshape(-1, d * d), (g.count, d * d)) * sw[:, None]).copy()
        if g.count == 1 and h.noise_kind != P.NOISE_UNIT:
            h.noise = h.noise.reshape(-1)
        out.groups.append(h)
    return out


class DeviceBackend:
    """Numeric work of GNC on the GPU through the C-ABI (nothing here computes)."""

    def __init__(self, ctx: Context, lm_params: LevenbergMarquardtParams):
        self.ctx, self.lm_params = ctx, lm_params
        self._eval = self._dp = None
        self._dp_sig = None

    def factor_errors(self, prob: P.Problem, values: np.ndarray) -> np.ndarray:
        """nfg_[k]->error(values) for every factor, graph order: 0.5 * |whitened residual|^2 = 0.5 * |b|^2 of
        the device linearization (b is the last column of the whitened block)."""
        if self._eval is None:
            self._eval = DeviceProblem(self.ctx, prob)
        self._eval.set_values(values)
        self._eval.linearize()
        out = np.zeros(prob.nfactors)
        for gi, g in enumerate(prob.groups):
            b = self._eval.get_jacobians(gi)[:, :, -1]
            out[graph_positions(g)] = 0.5 * np.sum(b * b, axis=1)
        return out

    def optimize(self, prob_w: P.Problem):
        """BaseOptimizer(graph_w, state_, params).optimize(); returns (values, graph_w.error(values)).
        The weighted graphs of one GNC run differ in their noise payloads and initial values only: the device
        problem (symbolic phase, keys, measurements) is built once and re-weighted through b200_set_group_noise."""
        sig = [g.keys for g in prob_w.groups]      # weighted_problem() shares the structural arrays (held: ids stay unique)
        if self._dp is None or len(sig) != len(self._dp_sig) or any(a is not b for a, b in zip(sig, self._dp_sig)):
            if self._dp is not None:
                self._dp.close()
            self._dp, self._dp_sig = DeviceProblem(self.ctx, prob_w), sig
        else:
            for gi, g in enumerate(prob_w.groups):
                self._dp.set_group_noise(gi, g.noise_kind, g.noise)
            self._dp.set_values(prob_w.values)
        dp = self._dp
        dp.prob = prob_w
        lm = LevenbergMarquardtOptimizer(self.ctx, prob_w, self.lm_params, device_problem=dp)
        lm.optimize()
        values, cost = dp.get_values(), lm.error()
        del lm
        return values, cost

    def close(self):
        for h in (self._eval, self._dp):
            if h is not None:
                h.close()
        self._eval = self._dp = None


class GncOptimizer:
    """Drop-in for gtsam::GncOptimizer<GncParams<LevenbergMarquardtParams>>."""

    def __init__(self, ctx: Optional[Context], problem: P.Problem, params: Optional[GncParams] = None, backend=None):
        self.params_ = params or GncParams()
        self.prob = strip_robust(problem)
        self.state_ = problem.values.copy()
        self.backend = backend or DeviceBackend(ctx, self.params_.baseOptimizerParams)
        n = self.prob.nfactors
        ki, ko = self.params_.knownInliers, self.params_.knownOutliers
        if set(ki) & set(ko):
            raise RuntimeError("GncOptimizer::constructor: the user has selected one or more measurements"
                               " to be BOTH a known inlier and a known outlier.")
        if any(i > n - 1 for i in ki) or any(i > n - 1 for i in ko):
            raise RuntimeError("GncOptimizer::constructor: known inliers / outliers that are not in the factor graph.")
        self.weights_ = self._initial_weights()
        self.setInlierCostThresholdsAtProbability(0.99)

    # -- thresholds (GncOptimizer.h:108-140) ---------------------------------------------------------------
    def setInlierCostThresholds(self, inth):
        inth = np.asarray(inth, dtype=np.float64)
        self.barcSq_ = np.full(self.prob.nfactors, float(inth)) if inth.ndim == 0 else inth.copy()

    def setInlierCostThresholdsAtProbability(self, alpha: float):
        self.barcSq_ = np.ones(self.prob.nfactors)
        for g in self.prob.groups:
            self.barcSq_[graph_positions(g)] = 0.5 * chi2inv(alpha, P.FACTOR_DIM[g.type])

    def getInlierCostThresholds(self):
        return self.barcSq_

    def getWeights(self):
        return self.weights_

    def setWeights(self, w):
        w = np.asarray(w, dtype=np.float64)
        if w.size != self.prob.nfactors:
            raise RuntimeError("GncOptimizer::setWeights: the number of specified weights does not match the size of the factor graph.")
        self.weights_ = w.copy()

    def _initial_weights(self):
        w = np.ones(self.prob.nfactors)
        w[self.params_.knownOutliers] = 0.0
        return w

    # -- GncOptimizer::optimize, :184-268 --------------------------------------------------------------------
    def optimize(self):
        p = self.params_
        result, prev_cost = self.backend.optimize(weighted_problem(self.prob, self.weights_, self.state_))
        mu = self.initializeMu()
        cost = 0.0
        self.mu_history = [mu]
        nr_unknown = self.prob.nfactors - (len(p.knownInliers) + len(p.knownOutliers))
        if mu <= 0 or nr_unknown == 0:
            return result
        for _ in range(p.maxIterations):
            self.weights_ = self.calculateWeights(result, mu)
            result, cost = self.backend.optimize(weighted_problem(self.prob, self.weights_, self.state_))
            if self.checkConvergence(mu, self.weights_, cost, prev_cost):
                break
            mu = self.updateMu(mu)
            self.mu_history.append(mu)
            prev_cost = cost
        return result

    def initializeMu(self) -> float:      # :271-311
        err0 = self.backend.factor_errors(self.prob, self.state_)
        if self.params_.lossType == GM:
            return float(np.max(2 * err0 / self.barcSq_, initial=0.0))
        den = 2 * err0 - self.barcSq_
        cand = self.barcSq_[den > 0] / den[den > 0]
        mu = float(cand.min()) if cand.size else float("inf")
        if 0 <= mu < 1e-6:
            mu = 1e-6
        return mu if (mu > 0 and np.isfinite(mu)) else -1.0

    def updateMu(self, mu: float) -> float:   # :314-327
        return max(1.0, mu / self.params_.muStep) if self.params_.lossType == GM else mu * self.params_.muStep

    def checkMuConvergence(self, mu: float) -> bool:
        return self.params_.lossType == GM and abs(mu - 1.0) < 1e-9

    def checkCostConvergence(self, cost: float, prev_cost: float) -> bool:
        return abs(cost - prev_cost) / max(prev_cost, 1e-7) < self.params_.relativeCostTol

    def checkWeightsConvergence(self, weights) -> bool:
        if self.params_.lossType != TLS:
            return False
        return bool(np.all(np.abs(weights - np.round(weights)) <= self.params_.weightsTol))

    def checkConvergence(self, mu, weights, cost, prev_cost) -> bool:
        return self.checkCostConvergence(cost, prev_cost) or self.checkWeightsConvergence(weights) or self.checkMuConvergence(mu)

    def calculateWeights(self, currentEstimate, mu: float):   # :414-468
        p = self.params_
        weights = self._initial_weights()
        known = np.zeros(self.prob.nfactors, dtype=bool)
        known[p.knownInliers] = True
        known[p.knownOutliers] = True
        u2 = self.backend.factor_errors(self.prob, currentEstimate)
        b = self.barcSq_
        if p.lossType == GM:
            w = ((mu * b) / (u2 + mu * b)) ** 2
        else:
            upper, lower = (mu + 1) / mu * b, mu / (mu + 1) * b
            with np.errstate(divide="ignore", invalid="ignore"):
                w = np.sqrt(b * mu * (mu + 1) / u2) - mu
            zero = (u2 >= upper) | (w < 0)
            one = ~zero & ((u2 <= lower) | (w > 1))
            w = np.where(zero, 0.0, np.where(one, 1.0, w))
        weights[~known] = w[~known]
        return weights
