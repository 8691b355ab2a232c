"""Stop-rule policy seam: same names, arguments and error behaviour as the reference's
``src/algorithms/dp_solver.py`` (/root/reference/src/algorithms/dp_solver.py:12-289), with
the arithmetic executed by libasd_b200.so (``asd_stop_rule_host`` for the scalar API the
pipeline calls per request, ``asd_stop_rule`` on the device for batches - see ``stop_rule_batch``).
Results are bit-exact with the reference (tests/test_stop_rule.py)."""
from __future__ import annotations

import ctypes
import functools
import logging
import operator
from typing import List, Tuple

import numpy as np

from .._lib import check, lib

logger = logging.getLogger(__name__)


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def optimal_stopping_rule(p: List[float], C: List[float], lam: float, risk_adjustment: bool = False,
                          alpha: float = 1.0, beta: float = 1.0) -> Tuple[int, List[float]]:
    """dp_solver.py:12-71.  Returns (k_star, J) with J of length L+1."""
    if len(p) != len(C):
        raise ValueError("p and C must have the same length")          # dp_solver.py:34-35
    L = len(C)
    if L == 0:
        # reference: reversed(range(0)) is empty, J = [0.0], next(...) default is L-1 = -1
        return -1, [0.0]
    pa, ca = _arr(p), _arr(C)
    J = np.zeros(L + 1, dtype=np.float64)
    k = lib().asd_stop_rule_host(pa.ctypes.data, ca.ctypes.data, L, float(lam), int(bool(risk_adjustment)),
                                 float(alpha), float(beta), J.ctypes.data)
    if k < 0:
        check(-1, "optimal_stopping_rule")
    return int(k), [float(x) for x in J]


def compute_expected_cost(p: List[float], C: List[float], lam: float, stopping_stage: int) -> float:
    """Expected cost of stopping at ``stopping_stage`` (dp_solver.py:74-103):
    sum_{i<=k} C_i + lam * (1 - prod_{i<=k} p_i).  Left-to-right folds keep the reference's binary64
    rounding (goldens: tests/golden/stop_rule_golden.json ``expected_cost``)."""
    k = stopping_stage + 1
    reach = functools.reduce(operator.mul, p[:k], 1.0)
    spent = functools.reduce(operator.add, C[:k], 0)
    return spent + lam * (1 - reach)


def bayesian_adjustment(p_hat: float, n_obs: int, alpha: float = 1.0, beta: float = 1.0) -> float:
    """dp_solver.py:106-130, evaluated by the library (binary64, same operation order)."""
    return float(lib().asd_bayesian_adjustment_host(float(p_hat), float(n_obs), float(alpha), float(beta)))


def stop_rule_batch(p, C, lam: float, risk_adjustment: bool = False, alpha: float = 1.0, beta: float = 1.0):
    """Device form: p, C are CUDA float64 tensors [n, L]; returns (k_star int32 [n], J float64 [n, L+1])
    without leaving the GPU (north-star item 3)."""
    import torch
    assert p.is_cuda and C.is_cuda and p.dtype == torch.float64 and C.dtype == torch.float64
    p, C = p.contiguous(), C.contiguous()
    n, L = p.shape
    if C.shape != p.shape:
        raise ValueError("p and C must have the same shape")
    k_star = torch.empty(n, dtype=torch.int32, device=p.device)
    J = torch.empty(n, L + 1, dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        rc = lib().asd_stop_rule(p.data_ptr(), C.data_ptr(), n, L, float(lam), int(bool(risk_adjustment)),
                                 float(alpha), float(beta), k_star.data_ptr(), J.data_ptr(),
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(rc, "asd_stop_rule")
    return k_star, J


def stop_rule_rows(p, C, lam_rows, risk_adjustment: bool = False, alpha: float = 1.0, beta: float = 1.0):
    """One library call for n independent stop decisions with one lambda per row.  numpy inputs ([n, L] float64,
    lam [n]) run through ``asd_stop_rule_rows_host``; CUDA tensors through ``asd_stop_rule_rows`` on the device.
    Returns (k_star int32 [n], J float64 [n, L+1]) of the same kind as the inputs."""
    if isinstance(p, np.ndarray) or isinstance(p, (list, tuple)):
        pa, ca, la = _arr(p), _arr(C), _arr(lam_rows)
        if pa.shape != ca.shape or pa.ndim != 2 or la.shape != (pa.shape[0],):
            raise ValueError("p and C must be [n, L] and lam [n]")
        n, L = pa.shape
        k = np.empty(n, dtype=np.int32)
        J = np.empty((n, L + 1), dtype=np.float64)
        if n:
            check(lib().asd_stop_rule_rows_host(pa.ctypes.data, ca.ctypes.data, la.ctypes.data, n, L,
                                                int(bool(risk_adjustment)), float(alpha), float(beta),
                                                k.ctypes.data, J.ctypes.data), "asd_stop_rule_rows_host")
        return k, J
    import torch
    assert p.is_cuda and p.dtype == torch.float64
    p, C, lam_rows = p.contiguous(), C.contiguous(), lam_rows.contiguous()
    n, L = p.shape
    k = torch.empty(n, dtype=torch.int32, device=p.device)
    J = torch.empty(n, L + 1, dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        rc = lib().asd_stop_rule_rows(p.data_ptr(), C.data_ptr(), lam_rows.data_ptr(), n, L, int(bool(risk_adjustment)),
                                      float(alpha), float(beta), k.data_ptr(), J.data_ptr(),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(rc, "asd_stop_rule_rows")
    return k, J


class OptimalStoppingTable:
    """Memo of stop decisions (API of dp_solver.py:133-210: ``precompute(cost_ratios, prob_grid)``,
    ``lookup(probabilities, lambda_value, fallback_to_dp)``, attributes ``lambda_values``, ``num_stages``,
    ``table[lam][prob_key]``).  Where the reference runs one Python DP per (lambda, scenario), ``precompute``
    here lays the whole lambda x grid product out as one [n, L] batch and evaluates it with a single library
    call (``stop_rule_rows``), then files the answers under the reference's keys (probabilities rounded to 2 dp)."""

    DEFAULT_COSTS = (1.0, 1.6, 4.2, 8.8)          # the fallback cost vector of dp_solver.py:205

    def __init__(self, lambda_values: List[float], num_stages: int = 4):
        self.lambda_values = lambda_values
        self.num_stages = num_stages
        self.table = {}

    @staticmethod
    def _key(probabilities) -> tuple:
        return tuple(round(float(x), 2) for x in probabilities)

    def precompute(self, cost_ratios: List[float], prob_grid: List[List[float]]):
        lams = list(self.lambda_values)
        for lam in lams:
            self.table[lam] = {}
        grid = [list(map(float, sc)) for sc in prob_grid]
        if not grid or not lams:
            return
        by_len = {}
        for sc in grid:                      # ragged grids are legal in the reference: batch per length
            if len(sc) != len(cost_ratios):
                raise ValueError("p and C must have the same length")
            by_len.setdefault(len(sc), []).append(sc)
        for L, rows in by_len.items():
            if L == 0:
                for lam in lams:
                    self.table[lam][()] = -1
                continue
            P = np.tile(_arr(rows), (len(lams), 1))
            Cm = np.broadcast_to(_arr(cost_ratios), P.shape)
            lam_rows = np.repeat(_arr(lams), len(rows))
            k_star, _ = stop_rule_rows(P, Cm, lam_rows)
            keys = [self._key(sc) for sc in rows]
            for li, lam in enumerate(lams):
                base = li * len(rows)
                # later scenarios overwrite earlier ones with the same rounded key, as in the reference's loop
                self.table[lam].update(zip(keys, (int(x) for x in k_star[base:base + len(rows)])))
        logger.info("optimal stopping table: %d lambdas x %d scenarios in %d batched call(s)", len(lams), len(grid),
                    len(by_len))

    def lookup(self, probabilities: List[float], lambda_value: float, fallback_to_dp: bool = True) -> int:
        nearest = min(self.lambda_values, key=lambda x: abs(x - lambda_value))
        hit = self.table.get(nearest, {}).get(self._key(probabilities))
        if hit is not None:
            return hit
        if not fallback_to_dp:
            return len(probabilities) - 1
        costs = list(self.DEFAULT_COSTS[:len(probabilities)])
        return optimal_stopping_rule(probabilities, costs, lambda_value)[0]


class AdaptiveStopping:
    """Online per-stage reward statistics with Hoeffding bounds (API of dp_solver.py:213-289:
    ``update_statistics``, ``get_confidence_bounds``, ``should_explore``; attributes ``lambda_value``,
    ``confidence_level``, ``stage_counts``, ``stage_rewards``, ``total_steps``)."""

    NUM_STAGES = 4
    MIN_PULLS = 10          # every stage is explored at least this often
    SLACK = 0.1             # a stage stays interesting while its upper bound is within SLACK of the best

    def __init__(self, initial_lambda: float = 1.0, confidence_level: float = 0.1):
        self.lambda_value = initial_lambda
        self.confidence_level = confidence_level
        self.stage_counts = np.zeros(self.NUM_STAGES)
        self.stage_rewards = np.zeros(self.NUM_STAGES)
        self.total_steps = 0

    def update_statistics(self, chosen_stage: int, observed_quality: float, observed_latency: float):
        reward = observed_quality - self.lambda_value * (observed_latency / 1000.0)     # latency in seconds
        seen = self.stage_counts[chosen_stage]
        # running mean in the reference's form ((n-1) * mean + reward) / n so the statistics agree to the last bit
        self.stage_rewards[chosen_stage] = (seen * self.stage_rewards[chosen_stage] + reward) / (seen + 1.0)
        self.stage_counts[chosen_stage] = seen + 1.0
        self.total_steps += 1

    def _radius(self) -> np.ndarray:
        with np.errstate(divide="ignore"):
            return np.sqrt(-np.log(self.confidence_level / 2) / (2 * self.stage_counts))

    def get_confidence_bounds(self, stage: int) -> Tuple[float, float]:
        if self.stage_counts[stage] == 0:
            return -np.inf, np.inf
        r = self._radius()[stage]
        return self.stage_rewards[stage] - r, self.stage_rewards[stage] + r

    def should_explore(self, stage: int) -> bool:
        if self.stage_counts[stage] < self.MIN_PULLS:
            return True
        upper = np.where(self.stage_counts > 0, self.stage_rewards + self._radius(), np.inf)
        return bool(upper[stage] >= upper.max() - self.SLACK)


class DynamicProgrammingSolver:
    """Imported by real_model_pipeline.py:39 but never defined by the reference
    (SURVEY.md Appendix A).  ``should_stop`` asks the DP whether stopping at ``stage_id`` is
    optimal given the quality estimate as this stage's acceptance probability and 1.0 for the
    later stages (the pipeline's own convention that the last stage always accepts,
    pipeline.py:241-242); the reference's stated fallback ``quality > 0.8 or last stage``
    (real_model_pipeline.py:473) is kept for degenerate inputs."""

    def __init__(self, num_stages: int, cost_vector: List[float]):
        self.num_stages = num_stages
        self.cost_vector = list(cost_vector)

    def update_costs(self, cost_vector: List[float]):
        self.cost_vector = list(cost_vector)

    def should_stop(self, stage_id: int, current_cost: float, quality_estimate: float, lambda_param: float) -> bool:
        if stage_id >= self.num_stages - 1:
            return True
        if not (0.0 <= quality_estimate <= 1.0) or len(self.cost_vector) != self.num_stages:
            return quality_estimate > 0.8
        C = self.cost_vector[stage_id:]
        p = [quality_estimate] + [1.0] * (len(C) - 1)
        k_star, _ = optimal_stopping_rule(p, C, lambda_param)
        return k_star == 0
