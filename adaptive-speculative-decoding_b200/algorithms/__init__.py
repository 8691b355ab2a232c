"""Mirror of the reference's ``src.algorithms`` package (src/algorithms/__init__.py)."""
from .dp_solver import (AdaptiveStopping, DynamicProgrammingSolver, OptimalStoppingTable, bayesian_adjustment,
                        compute_expected_cost, optimal_stopping_rule)

__all__ = ["optimal_stopping_rule", "compute_expected_cost", "bayesian_adjustment", "OptimalStoppingTable",
           "AdaptiveStopping", "DynamicProgrammingSolver"]
