"""directtrajectoryoptimization.jl_b200 -- B200-native batched NLP callbacks for direct trajectory
optimization (the MOI-evaluator hot path of thowell/DirectTrajectoryOptimization.jl).

Import as `import dto_b200` (the directory name contains a dot, so the repo-root shim
`dto_b200.py` loads this package under that name).

Exports mirror /root/reference/src/DirectTrajectoryOptimization.jl:22-35.
"""
from .elements import Bound, Constraint, Cost, Dynamics, GeneralConstraint  # noqa: F401
from .evaluator import (BatchedNLPData, Model, Solver, get_trajectory, initialize_controls,  # noqa: F401
                        initialize_states, linear_interpolation, solve, solver_from)

Bounds = list
Constraints = list
Objective = list

__all__ = ["Cost", "Bound", "Bounds", "Constraint", "Constraints", "GeneralConstraint", "Dynamics", "Solver",
           "initialize_states", "initialize_controls", "solve", "get_trajectory", "linear_interpolation",
           "BatchedNLPData", "Model", "solver_from", "Objective"]
