"""Stop-rule policy seam: same names, arguments and error behaviour as the reference's
``src/algorithms/dp_solver.py`` (/root/reference/src/algorithms/dp_solver.py:12-289), with
the arithmetic executed by libasd_b200.so (``asd_stop_rule_host`` for the scalar API the
pipeline calls per request, ``asd_stop_rule`` on the device for batches - see ``stop_rule_batch``).
Results are bit-exact with the reference (tests/test_stop_rule.py)."""
from __future__ import annotations

import ctypes
import logging
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
    """dp_solver.py:74-103 (host bookkeeping; a handful of binary64 operations in the
    reference's order)."""
    p_bar = 1.0
    for i in range(stopping_stage + 1):
        p_bar *= p[i]
    computation_cost = sum(C[:stopping_stage + 1])
    quality_loss = lam * (1 - p_bar)
    return computation_cost + quality_loss


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


class OptimalStoppingTable:
    """dp_solver.py:133-210: memo of stop decisions keyed on probabilities rounded to 2 dp."""

    def __init__(self, lambda_values: List[float], num_stages: int = 4):
        self.lambda_values = lambda_values
        self.num_stages = num_stages
        self.table = {}

    def precompute(self, cost_ratios: List[float], prob_grid: List[List[float]]):
        logger.info("Precomputing optimal stopping table...")
        for lam in self.lambda_values:
            self.table[lam] = {}
            for prob_scenario in prob_grid:
                k_star, _ = optimal_stopping_rule(prob_scenario, cost_ratios, lam)
                self.table[lam][tuple(round(p, 2) for p in prob_scenario)] = k_star
        logger.info(f"Precomputed table for {len(self.lambda_values)} lambda values")

    def lookup(self, probabilities: List[float], lambda_value: float, fallback_to_dp: bool = True) -> int:
        closest_lam = min(self.lambda_values, key=lambda x: abs(x - lambda_value))
        prob_key = tuple(round(p, 2) for p in probabilities)
        if closest_lam in self.table and prob_key in self.table[closest_lam]:
            return self.table[closest_lam][prob_key]
        if fallback_to_dp:
            logger.debug(f"Table miss for lambda={lambda_value}, computing DP")
            cost_ratios = [1.0, 1.6, 4.2, 8.8][:len(probabilities)]      # dp_solver.py:205
            k_star, _ = optimal_stopping_rule(probabilities, cost_ratios, lambda_value)
            return k_star
        return len(probabilities) - 1


class AdaptiveStopping:
    """dp_solver.py:213-289: Hoeffding / UCB bookkeeping around the stop rule (host statistics)."""

    def __init__(self, initial_lambda: float = 1.0, confidence_level: float = 0.1):
        self.lambda_value = initial_lambda
        self.confidence_level = confidence_level
        self.stage_counts = np.zeros(4)
        self.stage_rewards = np.zeros(4)
        self.total_steps = 0

    def update_statistics(self, chosen_stage: int, observed_quality: float, observed_latency: float):
        self.stage_counts[chosen_stage] += 1
        normalized_latency = observed_latency / 1000.0
        reward = observed_quality - self.lambda_value * normalized_latency
        n = self.stage_counts[chosen_stage]
        self.stage_rewards[chosen_stage] = ((n - 1) * self.stage_rewards[chosen_stage] + reward) / n
        self.total_steps += 1

    def get_confidence_bounds(self, stage: int) -> Tuple[float, float]:
        n = self.stage_counts[stage]
        if n == 0:
            return -np.inf, np.inf
        confidence_radius = np.sqrt(-np.log(self.confidence_level / 2) / (2 * n))
        mean_reward = self.stage_rewards[stage]
        return mean_reward - confidence_radius, mean_reward + confidence_radius

    def should_explore(self, stage: int) -> bool:
        if self.stage_counts[stage] < 10:
            return True
        upper_bounds = [self.get_confidence_bounds(i)[1] for i in range(4)]
        return upper_bounds[stage] >= max(upper_bounds) - 0.1


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
