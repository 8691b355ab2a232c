from .optimal_stopping import OptimalStoppingTheory, TheoreticalParameters

__all__ = ["OptimalStoppingTheory", "TheoreticalParameters"]
