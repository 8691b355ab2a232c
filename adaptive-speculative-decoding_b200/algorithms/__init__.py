"""Mirror of the reference's ``src.algorithms`` package (src/algorithms/__init__.py)."""
from .dp_solver import (AdaptiveStopping, DynamicProgrammingSolver, OptimalStoppingTable, bayesian_adjustment,
                        compute_expected_cost, optimal_stopping_rule)

from .optimizer import GridSearchOptimizer, LambdaOptimizer, OptimizationResult, find_optimal_lambda

__all__ = ["LambdaOptimizer", "GridSearchOptimizer", "OptimizationResult", "find_optimal_lambda", "optimal_stopping_rule", "compute_expected_cost", "bayesian_adjustment", "OptimalStoppingTable",
           "AdaptiveStopping", "DynamicProgrammingSolver"]
