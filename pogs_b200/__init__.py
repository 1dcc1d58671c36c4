"""pogs_b200 -- B200-native (sm_100a CUDA) implementation of the POGS graph-form
ADMM hot path behind the reference's C ABI and Python surface.

    from pogs_b200 import solve_lasso, Solver
"""
from .graph import (Function, FunctionObj, FunctionVector, Ordering, _solve_graph_form, solve_elastic_net,
                    solve_huber, solve_lasso, solve_logistic, solve_nonneg_ls, solve_ridge, solve_svm)
from .solver import Solver, lasso_path
from .cone import Cone, solve_cone

__version__ = "0.1.0"

__all__ = ["Cone", "solve_cone", "Function", "FunctionObj", "FunctionVector", "Ordering", "Solver", "lasso_path", "solve_elastic_net",
           "solve_huber", "solve_lasso", "solve_logistic", "solve_nonneg_ls", "solve_ridge", "solve_svm"]
