"""Lambda tuning around the cascade: same names, arguments and search procedures as the
reference's src/algorithms/optimizer.py (``OptimizationResult`` :15-21, ``LambdaOptimizer`` :24-205,
``find_optimal_lambda`` :208-258, ``GridSearchOptimizer`` :261-353).  Pure host control flow that calls
``pipeline.process_request`` - which runs on the B200 engine - and reads the result like a dict and
``pipeline.lambda_value`` like an attribute (optimizer.py:231,237-241), both of which our pipeline supports."""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)


@dataclass
class OptimizationResult:
    optimal_lambda: float
    achieved_latency: float
    achieved_quality: float
    constraint_satisfied: bool
    iterations: int


class LambdaOptimizer:
    def __init__(self, latency_constraint: Optional[float] = None, quality_constraint: Optional[float] = None,
                 lambda_bounds: Tuple[float, float] = (0.01, 100.0)):
        self.latency_constraint = latency_constraint
        self.quality_constraint = quality_constraint
        self.lambda_bounds = lambda_bounds

    def optimize_for_latency_constraint(self, evaluate_function: Callable[[float], Tuple[float, float]],
                                        tolerance: float = 1e-3, max_iterations: int = 50) -> OptimizationResult:
        """bisection on lambda: a satisfied constraint moves the upper end down (optimizer.py:80-107)"""
        if self.latency_constraint is None:
            raise ValueError("Latency constraint must be set")
        low, high = self.lambda_bounds
        best_lambda, iterations = low, 0
        for it in range(max_iterations):
            iterations = it + 1
            mid = (low + high) / 2
            latency, _quality = evaluate_function(mid)
            if latency <= self.latency_constraint:
                best_lambda, high = mid, mid
            else:
                low = mid
            if high - low < tolerance:
                break
        final_latency, final_quality = evaluate_function(best_lambda)
        return OptimizationResult(best_lambda, final_latency, final_quality, final_latency <= self.latency_constraint,
                                  iterations)

    def optimize_pareto_front(self, evaluate_function: Callable[[float], Tuple[float, float]],
                              num_points: int = 20) -> List[Tuple[float, float, float]]:
        """log-spaced sweep, keep the non-dominated (latency, quality) points (optimizer.py:124-155)"""
        lams = np.logspace(np.log10(self.lambda_bounds[0]), np.log10(self.lambda_bounds[1]), num_points)
        pts = [(float(l), *map(float, evaluate_function(float(l)))) for l in lams]
        front = [p for p in pts if not any((o[1] <= p[1] and o[2] >= p[2]) and (o[1] < p[1] or o[2] > p[2]) for o in pts)]
        return sorted(front, key=lambda p: p[1])

    def find_balanced_lambda(self, evaluate_function: Callable[[float], Tuple[float, float]],
                             quality_weight: float = 0.5) -> OptimizationResult:
        from scipy.optimize import minimize_scalar

        def objective(lam: float) -> float:
            latency, quality = evaluate_function(lam)
            return -(quality_weight * quality - (1 - quality_weight) * latency / 1000.0)   # optimizer.py:174-185

        res = minimize_scalar(objective, bounds=self.lambda_bounds, method="bounded")
        lat, q = evaluate_function(res.x)
        return OptimizationResult(float(res.x), lat, q, True, int(res.nit))


def find_optimal_lambda(pipeline, test_prompts: List[str], constraint_type: str = "latency",
                        constraint_value: float = 1000.0, num_evaluations: int = 10) -> float:
    """optimizer.py:208-258"""

    def evaluate_lambda(lam: float) -> Tuple[float, float]:
        pipeline.lambda_value = lam
        lat, qual = [], []
        for prompt in test_prompts[:num_evaluations]:
            r = pipeline.process_request(prompt)
            lat.append(r["latency_ms"])
            qual.append(min(1.0, len(r["output"].split()) / 50.0))
        return float(np.mean(lat)), float(np.mean(qual))

    opt = LambdaOptimizer()
    if constraint_type == "latency":
        opt.latency_constraint = constraint_value
        return opt.optimize_for_latency_constraint(evaluate_lambda).optimal_lambda
    return opt.find_balanced_lambda(evaluate_lambda).optimal_lambda


class GridSearchOptimizer:
    """optimizer.py:261-353: evaluate a lambda grid on a sample set, report the best by score."""

    def __init__(self, lambda_values: Optional[List[float]] = None,
                 evaluation_function: Optional[Callable] = None):
        self.lambda_values = lambda_values or [0.1, 0.5, 1.0, 2.0, 5.0, 10.0]
        self.evaluation_function = evaluation_function or self._default_evaluation

    def search(self, pipeline, samples: List[Dict], quality_weight: float = 0.5) -> Dict:
        results = {}
        for lam in self.lambda_values:
            pipeline.lambda_value = lam
            rows = [self.evaluation_function(pipeline, s) for s in samples]
            lat = float(np.mean([r["latency_ms"] for r in rows]))
            qual = float(np.mean([r["quality"] for r in rows]))
            cost = float(np.mean([r["cost"] for r in rows]))
            results[lam] = {"latency_ms": lat, "quality": qual, "cost": cost,
                            "score": quality_weight * qual - (1 - quality_weight) * lat / 1000.0}
        best = max(results, key=lambda l: results[l]["score"])
        return {"best_lambda": best, "results": results}

    def _default_evaluation(self, pipeline, sample: Dict) -> Dict:
        r = pipeline.process_request(sample["prompt"])
        return {"latency_ms": r["latency_ms"], "quality": min(1.0, len(r["output"].split()) / 50.0),
                "cost": sum(r["costs"]), "stopped_at_stage": r["stopped_at_stage"]}
