"""ORACLE (test infrastructure only): the reference's exported names
(/root/reference/src/DirectTrajectoryOptimization.jl:22-35) bound to the CPU restatement."""
from .elements import Bound, Constraint, Cost, Dynamics, GeneralConstraint  # noqa: F401
from .nlp import NLPData, Solver, TrajectoryOptimizationData, linear_interpolation  # noqa: F401


def solver_from(model: dict, parameters=None) -> Solver:
    """Assemble a Solver from an examples.models builder dict."""
    return Solver(model["dynamics"], model["objective"], model["constraints"], model["bounds"],
                  evaluate_hessian=model.get("evaluate_hessian", False), general_constraint=model.get("general"),
                  parameters=parameters)
