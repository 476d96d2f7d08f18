"""Host-side mirror of the reference's element constructors (same names, argument order and
keyword names): Cost (/root/reference/src/costs.jl:13-45), Dynamics (src/dynamics.jl:18-101),
Constraint (src/constraints.jl:21-78), GeneralConstraint (src/general_constraint.jl:18-71),
Bound (src/bounds.jl:1-16).

Where the reference `eval`s Julia closures into `::Any` fields, these objects hold an
ElementSpec -- the traced expressions, structural patterns and derivative expressions -- that
the code generator lowers to CUDA device functions. Nothing here evaluates numbers.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import sympy as sp

from . import symbolic as S
from .codegen import ElementSpec, GeneralSpec


class Cost:
    def __init__(self, f: Callable, num_state: int, num_action: int, num_parameter: int = 0,
                 evaluate_hessian: bool = False):
        x, u, w = S.variables("x", num_state), S.variables("u", num_action), S.variables("w", num_parameter)
        evaluate = S.flatten(f(x, u, w))[0]
        xu = list(x) + list(u)
        self.num_state, self.num_action, self.num_parameter = num_state, num_action, num_parameter
        self.num_gradient = num_state + num_action
        hr, hc = [], []
        if evaluate_hessian:
            hr, hc = S.hessian_pattern(evaluate, xu)
        self.num_hessian = len(hr)
        self.sparsity = [hr, hc] if evaluate_hessian else [[]]
        self.spec = ElementSpec(role="cost", n_out=1, nx=num_state, nu=num_action, nw=num_parameter,
                                args={"x": list(x), "u": list(u), "w": list(w)}, evaluate=[evaluate],
                                jac_rows=[1] * len(xu), jac_cols=list(range(1, len(xu) + 1)),
                                has_hess=bool(evaluate_hessian), hess_rows=hr, hess_cols=hc, vars=xu)


class Dynamics:
    """Dynamics(f, num_next_state, num_state, num_action; num_parameter, evaluate_hessian) with
    f(y, x, u, w); or Dynamics(f, jacobian, ny, nx, nu) where jacobian(J, y, x, u, w) fills a dense
    ny x (nx+nu+ny) matrix (src/dynamics.jl:59-101). Both closures must be traceable (pure
    arithmetic on their arguments): arbitrary host closures cannot run on the GPU and there is
    no CPU fallback."""

    def __init__(self, f: Callable, *rest, num_parameter: int = 0, evaluate_hessian: bool = False):
        user_jac = None
        if rest and callable(rest[0]):
            user_jac, rest = rest[0], rest[1:]
        num_next_state, num_state, num_action = rest
        y, x = S.variables("y", num_next_state), S.variables("x", num_state)
        u, w = S.variables("u", num_action), S.variables("w", num_parameter)
        try:
            evaluate = S.flatten(f(y, x, u, w))
        except Exception as e:  # noqa: BLE001
            raise TypeError("Dynamics: the function could not be traced symbolically; the CUDA backend needs "
                            "expressions, not opaque host closures") from e
        xuy = list(x) + list(u) + list(y)
        self.num_next_state, self.num_state, self.num_action = num_next_state, num_state, num_action
        self.num_parameter = num_parameter
        hr, hc = [], []
        jv = None
        if user_jac is not None:
            nv = len(xuy)
            M = np.zeros((num_next_state, nv), dtype=object)
            user_jac(M, y, x, u, w)
            jr = [i for j in range(1, nv + 1) for i in range(1, num_next_state + 1)]
            jc = [j for j in range(1, nv + 1) for i in range(1, num_next_state + 1)]
            jv = [sp.sympify(M[i - 1, j - 1]) for i, j in zip(jr, jc)]
            evaluate_hessian = False
        else:
            jr, jc = S.jacobian_pattern(evaluate, xuy)
        lam = S.variables("lam", num_next_state)
        if evaluate_hessian:
            lag = S.dot(lam, evaluate)
            hr, hc = S.hessian_pattern(lag, xuy)
        self.num_jacobian, self.num_hessian = len(jr), len(hr)
        self.jacobian_sparsity = [jr, jc]
        self.hessian_sparsity = [hr, hc] if evaluate_hessian else [[]]
        self.spec = ElementSpec(role="dyn", n_out=num_next_state, nx=num_state, nu=num_action, nw=num_parameter,
                                args={"y": list(y), "x": list(x), "u": list(u), "w": list(w), "lam": list(lam)},
                                evaluate=evaluate, jac_rows=jr, jac_cols=jc, has_hess=bool(evaluate_hessian),
                                hess_rows=hr, hess_cols=hc, vars=xuy, lam=list(lam), _jac=jv, user_jac=jv is not None)


class Constraint:
    def __init__(self, f: Optional[Callable] = None, num_state: int = 0, num_action: int = 0, num_parameter: int = 0,
                 indices_inequality: Sequence[int] = (), evaluate_hessian: bool = False):
        self.indices_inequality = list(indices_inequality)
        self.num_state, self.num_action, self.num_parameter = num_state, num_action, num_parameter
        if f is None:  # Constraint(): empty (src/constraints.jl:66-78)
            self.num_constraint = self.num_jacobian = self.num_hessian = 0
            self.jacobian_sparsity = [[], []]
            self.hessian_sparsity = [[], []]
            self.spec = None
            return
        x, u, w = S.variables("x", num_state), S.variables("u", num_action), S.variables("w", num_parameter)
        evaluate = S.flatten(f(x, u, w))
        xu = list(x) + list(u)
        jr, jc = S.jacobian_pattern(evaluate, xu)
        lam = S.variables("lam", len(evaluate))
        hr, hc = [], []
        if evaluate_hessian:
            lag = S.dot(lam, evaluate)
            hr, hc = S.hessian_pattern(lag, xu)
        self.num_constraint, self.num_jacobian, self.num_hessian = len(evaluate), len(jr), len(hr)
        self.jacobian_sparsity = [jr, jc]
        self.hessian_sparsity = [hr, hc] if evaluate_hessian else [[]]
        self.spec = ElementSpec(role="stage", n_out=len(evaluate), nx=num_state, nu=num_action, nw=num_parameter,
                                args={"x": list(x), "u": list(u), "w": list(w), "lam": list(lam)}, evaluate=evaluate,
                                jac_rows=jr, jac_cols=jc, has_hess=bool(evaluate_hessian), hess_rows=hr,
                                hess_cols=hc, ineq=list(indices_inequality), vars=xu, lam=list(lam))


class GeneralConstraint:
    def __init__(self, f: Optional[Callable] = None, num_variables: int = 0, num_parameter: int = 0,
                 indices_inequality: Sequence[int] = (), evaluate_hessian: bool = False):
        self.indices_inequality = list(indices_inequality)
        self.num_variables, self.num_parameter = num_variables, num_parameter
        if f is None:
            self.num_constraint = self.num_jacobian = self.num_hessian = 0
            self.spec = None
            return
        z, w = S.variables("z", num_variables), S.variables("w", num_parameter)
        evaluate = S.flatten(f(z, w))
        jr, jc = S.jacobian_pattern(evaluate, list(z))
        jv = S.jacobian_values(evaluate, list(z), jr, jc)
        lam = S.variables("lam", len(evaluate))
        hr, hc, hv = [], [], []
        if evaluate_hessian:
            lag = S.dot(lam, evaluate)
            hr, hc = S.hessian_pattern(lag, list(z))
            hv = S.hessian_values(lag, list(z), hr, hc)
        self.num_constraint, self.num_jacobian, self.num_hessian = len(evaluate), len(jv), len(hv)
        self.jacobian_sparsity = [jr, jc]
        self.hessian_sparsity = [hr, hc] if evaluate_hessian else [[]]
        self.spec = GeneralSpec(num_variables=num_variables, num_parameter=num_parameter,
                                args={"z": list(z), "w": list(w), "lam": list(lam)}, evaluate=evaluate, jac_rows=jr,
                                jac_cols=jc, jac=jv, has_hess=bool(evaluate_hessian), hess_rows=hr, hess_cols=hc,
                                hess=hv, ineq=list(indices_inequality))


class Bound:
    def __init__(self, num_state: int = 0, num_action: int = 0, state_lower=None, state_upper=None,
                 action_lower=None, action_upper=None):
        self.state_lower = np.full(num_state, -np.inf) if state_lower is None else np.asarray(state_lower, float)
        self.state_upper = np.full(num_state, np.inf) if state_upper is None else np.asarray(state_upper, float)
        self.action_lower = np.full(num_action, -np.inf) if action_lower is None else np.asarray(action_lower, float)
        self.action_upper = np.full(num_action, np.inf) if action_upper is None else np.asarray(action_upper, float)
