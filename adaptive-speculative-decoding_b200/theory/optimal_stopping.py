"""Threshold variant of the stop rule: same names and results as the reference's
src/theory/optimal_stopping.py:15-91 (``derive_optimal_policy``), bit-exact in binary64
(tests/test_api_surface.py against goldens generated from the reference).  The regret / sample
complexity formulas of the same file are paper mathematics with no runtime role."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List


@dataclass
class TheoreticalParameters:
    n_stages: int = 4
    quality_bounds: List[float] = None
    cost_ratios: List[float] = None
    lambda_param: float = 1.0
    epsilon: float = 0.1
    delta: float = 0.05


class OptimalStoppingTheory:
    def __init__(self, params: TheoreticalParameters):
        self.params = params
        if params.quality_bounds is None:
            self.params.quality_bounds = [0.7, 0.8, 0.85, 0.9]       # optimal_stopping.py:40
        if params.cost_ratios is None:
            self.params.cost_ratios = [1.0, 2.0, 4.5, 10.0]          # :43

    def _compute_improvement_probability(self, stage: int) -> float:
        return 0.6 * (1 - self.params.quality_bounds[stage])         # :91

    def derive_optimal_policy(self) -> Dict[int, float]:
        n, q, c, lam = self.params.n_stages, self.params.quality_bounds, self.params.cost_ratios, self.params.lambda_param
        V = [0.0] * (n + 1)
        thresholds: Dict[int, float] = {}
        for s in range(n - 1, -1, -1):                               # :59-80
            r_stop = q[s] - lam * c[s]
            if s < n - 1:
                p_improve = self._compute_improvement_probability(s)
                r_continue = p_improve * V[s + 1] + (1 - p_improve) * r_stop
            else:
                r_continue = float("-inf")
            V[s] = max(r_stop, r_continue)
            if s < n - 1:
                thresholds[s] = (V[s + 1] + lam * c[s]) / (1 + lam * (c[s + 1] - c[s]))
            else:
                thresholds[s] = 0
        return thresholds

    def should_stop(self, stage: int, predicted_quality: float) -> bool:
        """the threshold test of src/minimal_adaptive_decoder.py:153-164"""
        return predicted_quality >= self.derive_optimal_policy()[stage]
