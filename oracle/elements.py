"""ORACLE (test infrastructure only). CPU restatement of the reference's element
constructors and per-trajectory loops:

  Cost              /root/reference/src/costs.jl:1-107
  Dynamics          /root/reference/src/dynamics.jl:1-211
  Constraint        /root/reference/src/constraints.jl:1-183
  GeneralConstraint /root/reference/src/general_constraint.jl:1-139
  Bound             /root/reference/src/bounds.jl:1-16

Everything is 1-based exactly like the Julia source (index lists hold Julia
indices; python code subtracts one at the point of use) so that structures can be
compared bit-for-bit with what a Julia run would print. Loops are serial per knot,
one closure call per knot, cache -> slice copy, as in the reference.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import symbolics as S


def _vec(e) -> list:
    """Flatten a user-function return value (scalar / list / ndarray) to a list."""
    if isinstance(e, np.ndarray):
        return list(e.reshape(-1))
    if isinstance(e, (list, tuple)):
        out = []
        for v in e:
            out.extend(_vec(v))
        return out
    return [e]


def _dot(a, b):
    s = 0
    for x, y in zip(a, b):
        s = s + x * y
    return s


# ------------------------------------------------------------------ Cost
class Cost:
    """src/costs.jl:1-45. evaluate(out1, x, u, w); gradient(out, x, u, w) dense over
    [x; u]; hessian(out, x, u, w) = nzval of sparsehessian (NOT scaled)."""

    def __init__(self, f: Callable, num_state: int, num_action: int, num_parameter: int = 0,
                 evaluate_hessian: bool = False):
        x = S.variables("x", num_state)
        u = S.variables("u", num_action)
        w = S.variables("w", num_parameter)
        evaluate = _vec(f(x, u, w))[0]
        xu = list(x) + list(u)
        grad = S.gradient(evaluate, xu)
        self.sym = dict(x=x, u=u, w=w, evaluate=[evaluate], gradient=grad)
        self.num_state, self.num_action, self.num_parameter = num_state, num_action, num_parameter
        self.num_gradient = num_state + num_action
        self.evaluate = S.build_function([evaluate], x, u, w)
        self.gradient = S.build_function(grad, x, u, w)
        if evaluate_hessian:
            rows, cols, nz = S.sparsehessian(evaluate, xu)
            self.hessian = S.build_function(nz, x, u, w)
            self.sparsity = [rows, cols]
            self.num_hessian = len(nz)
            self.sym["hessian"] = nz
        else:
            self.hessian = None  # Expr(:null) -- calling it throws (Q9)
            self.sparsity = [[]]
            self.num_hessian = 0
        self.evaluate_cache = np.zeros(1)
        self.gradient_cache = np.zeros(self.num_gradient)
        self.hessian_cache = np.zeros(self.num_hessian)


def cost(objective: Sequence[Cost], states, actions, parameters) -> float:
    """src/costs.jl:49-56 -- serial sum in t order."""
    J = 0.0
    for t, c in enumerate(objective):
        c.evaluate(c.evaluate_cache, states[t], actions[t], parameters[t])
        J += c.evaluate_cache[0]
    return J


def gradient_(gradient, indices, objective, states, actions, parameters) -> None:
    """src/costs.jl:58-64"""
    for t, c in enumerate(objective):
        c.gradient(c.gradient_cache, states[t], actions[t], parameters[t])
        gradient[np.asarray(indices[t], dtype=np.int64) - 1] += c.gradient_cache


def hessian_(hessian, indices, objective, states, actions, parameters, scaling) -> None:
    """src/costs.jl:66-73 -- scaling applied in the cache, then accumulated."""
    for t, c in enumerate(objective):
        c.hessian(c.hessian_cache, states[t], actions[t], parameters[t])  # throws if None (Q9)
        c.hessian_cache *= scaling
        idx = np.asarray(indices[t], dtype=np.int64) - 1
        for k in range(len(idx)):  # .+= with possibly repeated slots is sequential per entry
            hessian[idx[k]] += c.hessian_cache[k]


def _shift(t0: int, num_state, num_action) -> int:
    """(t > 1 ? sum(num_state[1:t-1]) + sum(num_action[1:t-1]) : 0) with t = t0+1."""
    return int(sum(num_state[:t0]) + sum(num_action[:t0]))


def sparsity_hessian_objective(objective, num_state, num_action) -> List[Tuple[int, int]]:
    """src/costs.jl:75-86"""
    row, col = [], []
    for t, c in enumerate(objective):
        if len(c.sparsity[0]) > 0:
            sh = _shift(t, num_state, num_action)
            row.extend(r + sh for r in c.sparsity[0])
            col.extend(cc + sh for cc in c.sparsity[1])
    return list(zip(row, col))


def _findfirst(key, rc):
    """[findfirst(x -> x == i, key) for i in rc] (1-based)."""
    pos = {}
    for i, k in enumerate(key):
        pos.setdefault(k, i + 1)
    return [pos[i] for i in rc]


def hessian_indices_objective(objective, key, num_state, num_action):
    """src/costs.jl:88-104"""
    indices = []
    for t, c in enumerate(objective):
        if len(c.sparsity[0]) > 0:
            sh = _shift(t, num_state, num_action)
            rc = [(r + sh, cc + sh) for r, cc in zip(c.sparsity[0], c.sparsity[1])]
            indices.append(_findfirst(key, rc))
        else:
            indices.append([])
    return indices


# ------------------------------------------------------------------ Dynamics
class Dynamics:
    """src/dynamics.jl:1-57. Variables ordered [x; u; y]; closures take (out, y, x, u, w[, λ])."""

    def __init__(self, f: Callable, *rest, num_parameter: int = 0, evaluate_hessian: bool = False,
                 jacobian: Optional[Callable] = None):
        if rest and callable(rest[0]):
            # Dynamics(f, jacobian, ny, nx, nu) (src/dynamics.jl:59): value-returning python functions are
            # wrapped into the in-place closures the reference's second constructor stores
            user_f, user_j, rest = f, rest[0], rest[1:]

            def f(out, y, x, u, w, _f=user_f):  # noqa: F811
                out[:] = np.asarray(_f(y, x, u, w), dtype=float)

            jacobian = user_j
        num_next_state, num_state, num_action = rest
        self.num_next_state, self.num_state, self.num_action = num_next_state, num_state, num_action
        self.num_parameter = num_parameter
        if jacobian is not None:
            self._init_user(f, jacobian)
            return
        y = S.variables("y", num_next_state)
        x = S.variables("x", num_state)
        u = S.variables("u", num_action)
        w = S.variables("w", num_parameter)
        evaluate = _vec(f(y, x, u, w))
        xuy = list(x) + list(u) + list(y)
        rows, cols, nz = S.sparsejacobian(evaluate, xuy)
        self.sym = dict(y=y, x=x, u=u, w=w, evaluate=evaluate, jacobian=nz)
        self.evaluate = S.build_function(evaluate, y, x, u, w)
        self.jacobian = S.build_function(nz, y, x, u, w)
        self.num_jacobian = len(nz)
        self.jacobian_sparsity = [rows, cols]
        if evaluate_hessian:
            lam = S.variables("λ", num_next_state)
            lag = _dot(lam, evaluate)
            hr, hc, hnz = S.sparsehessian(lag, xuy)
            self.hessian = S.build_function(hnz, y, x, u, w, lam)
            self.hessian_sparsity = [hr, hc]
            self.num_hessian = len(hnz)
            self.sym.update(lam=lam, hessian=hnz)
        else:
            self.hessian = None
            self.hessian_sparsity = [[]]
            self.num_hessian = 0
        self.evaluate_cache = np.zeros(num_next_state)
        self.jacobian_cache = np.zeros(self.num_jacobian)
        self.hessian_cache = np.zeros(self.num_hessian)

    def _init_user(self, constraint, constraint_jacobian):
        """Second constructor, src/dynamics.jl:59-101: arbitrary closures, dense
        column-major ny x (nx+nu+ny) Jacobian pattern, no Hessian (Q14)."""
        ny = self.num_next_state
        nv = self.num_state + self.num_action + ny

        def jac(J, y, x, u, w):
            M = np.zeros((ny, nv))
            constraint_jacobian(M, y, x, u, w)
            J[:] = M.reshape(-1, order="F")

        self.evaluate = constraint
        self.jacobian = jac
        self.num_jacobian = ny * nv
        row, col = [], []
        for j in range(1, nv + 1):
            for i in range(1, ny + 1):
                row.append(i)
                col.append(j)
        self.jacobian_sparsity = [row, col]
        self.hessian = None
        self.hessian_sparsity = [[]]
        self.num_hessian = 0
        self.sym = None
        self.evaluate_cache = np.zeros(ny)
        self.jacobian_cache = np.zeros(self.num_jacobian)
        self.hessian_cache = np.zeros(0)


def constraints_dynamics(violations, indices, dynamics, states, actions, parameters):
    """src/dynamics.jl:103-109 -- argument order (y, x, u, w)."""
    for t, con in enumerate(dynamics):
        con.evaluate(con.evaluate_cache, states[t + 1], states[t], actions[t], parameters[t])
        violations[np.asarray(indices[t], dtype=np.int64) - 1] = con.evaluate_cache
        con.evaluate_cache[:] = 0.0


def jacobian_dynamics(jacobians, indices, dynamics, states, actions, parameters):
    """src/dynamics.jl:111-117"""
    for t, con in enumerate(dynamics):
        con.jacobian(con.jacobian_cache, states[t + 1], states[t], actions[t], parameters[t])
        jacobians[np.asarray(indices[t], dtype=np.int64) - 1] = con.jacobian_cache
        con.jacobian_cache[:] = 0.0


def hessian_lagrangian_dynamics(hessians, indices, dynamics, states, actions, parameters, duals):
    """src/dynamics.jl:119-127"""
    for t, con in enumerate(dynamics):
        if len(con.hessian_cache) > 0:
            con.hessian(con.hessian_cache, states[t + 1], states[t], actions[t], parameters[t], duals[t])
            idx = np.asarray(indices[t], dtype=np.int64) - 1
            for k in range(len(idx)):
                hessians[idx[k]] += con.hessian_cache[k]
            con.hessian_cache[:] = 0.0


def sparsity_jacobian_dynamics(dynamics, num_state, num_action, row_shift=0):
    """src/dynamics.jl:129-142"""
    row, col = [], []
    for t, con in enumerate(dynamics):
        cs = _shift(t, num_state, num_action)
        row.extend(r + row_shift for r in con.jacobian_sparsity[0])
        col.extend(c + cs for c in con.jacobian_sparsity[1])
        row_shift += con.num_next_state
    return list(zip(row, col))


def sparsity_hessian_dynamics(dynamics, num_state, num_action):
    """src/dynamics.jl:144-155"""
    row, col = [], []
    for t, con in enumerate(dynamics):
        if len(con.hessian_sparsity[0]) > 0:
            sh = _shift(t, num_state, num_action)
            row.extend(r + sh for r in con.hessian_sparsity[0])
            col.extend(c + sh for c in con.hessian_sparsity[1])
    return list(zip(row, col))


def num_state_action_next_state(dynamics):
    return sum(d.num_state + d.num_action for d in dynamics) + dynamics[-1].num_next_state


def num_constraint_dynamics(dynamics):
    return sum(d.num_next_state for d in dynamics)


def num_jacobian_dynamics(dynamics):
    return sum(d.num_jacobian for d in dynamics)


def constraint_indices_dynamics(dynamics, shift=0):
    """src/dynamics.jl:162-165"""
    out, acc = [], 0
    for d in dynamics:
        out.append([shift + acc + i for i in range(1, d.num_next_state + 1)])
        acc += d.num_next_state
    return out


def jacobian_indices_dynamics(dynamics, shift=0):
    """src/dynamics.jl:167-170"""
    out, acc = [], 0
    for d in dynamics:
        out.append([shift + acc + i for i in range(1, d.num_jacobian + 1)])
        acc += d.num_jacobian
    return out


def hessian_indices_dynamics(dynamics, key, num_state, num_action):
    """src/dynamics.jl:172-186"""
    indices = []
    for t, con in enumerate(dynamics):
        if len(con.hessian_sparsity[0]) > 0:
            sh = _shift(t, num_state, num_action)
            rc = [(r + sh, c + sh) for r, c in zip(con.hessian_sparsity[0], con.hessian_sparsity[1])]
            indices.append(_findfirst(key, rc))
        else:
            indices.append([])
    return indices


def state_indices(dynamics):
    """src/dynamics.jl:188-191"""
    out, acc = [], 0
    for d in dynamics:
        out.append([acc + i for i in range(1, d.num_state + 1)])
        acc += d.num_state + d.num_action
    out.append([acc + i for i in range(1, dynamics[-1].num_next_state + 1)])
    return out


def action_indices(dynamics):
    """src/dynamics.jl:193-195 (T-1 entries)"""
    out, acc = [], 0
    for d in dynamics:
        out.append([acc + d.num_state + i for i in range(1, d.num_action + 1)])
        acc += d.num_state + d.num_action
    return out


def state_action_indices(dynamics):
    """src/dynamics.jl:197-200"""
    out, acc = [], 0
    for d in dynamics:
        out.append([acc + i for i in range(1, d.num_state + d.num_action + 1)])
        acc += d.num_state + d.num_action
    out.append([acc + i for i in range(1, dynamics[-1].num_next_state + 1)])
    return out


def state_action_next_state_indices(dynamics):
    """src/dynamics.jl:202-204"""
    out, acc = [], 0
    for d in dynamics:
        out.append([acc + i for i in range(1, d.num_state + d.num_action + d.num_next_state + 1)])
        acc += d.num_state + d.num_action
    return out


def dimensions(dynamics, parameters=None):
    """src/dynamics.jl:206-211"""
    states = [d.num_state for d in dynamics] + [dynamics[-1].num_next_state]
    actions = [d.num_action for d in dynamics] + [0]
    if parameters is None:
        parameters = [0 for _ in range(len(dynamics) + 1)]
    return states, actions, parameters


# ------------------------------------------------------------------ Constraint
class Constraint:
    """src/constraints.jl:1-78. Variables [x; u]; closures (out, x, u, w[, λ])."""

    def __init__(self, f: Optional[Callable] = None, num_state: int = 0, num_action: int = 0,
                 num_parameter: int = 0, indices_inequality: Sequence[int] = (),
                 evaluate_hessian: bool = False):
        self.indices_inequality = list(indices_inequality)
        if f is None:  # Constraint(): zero-size no-op (src/constraints.jl:66-78)
            self.evaluate = lambda out, x, u, w: None
            self.jacobian = lambda out, x, u, w: None
            self.hessian = lambda out, x, u, w, lam=None: None
            self.num_state = self.num_action = self.num_parameter = 0
            self.num_constraint = self.num_jacobian = self.num_hessian = 0
            self.jacobian_sparsity = [[], []]
            self.hessian_sparsity = [[], []]
            self.evaluate_cache = np.zeros(0)
            self.jacobian_cache = np.zeros(0)
            self.hessian_cache = np.zeros(0)
            self.sym = None
            return
        x = S.variables("x", num_state)
        u = S.variables("u", num_action)
        w = S.variables("w", num_parameter)
        evaluate = _vec(f(x, u, w))
        xu = list(x) + list(u)
        rows, cols, nz = S.sparsejacobian(evaluate, xu)
        self.sym = dict(x=x, u=u, w=w, evaluate=evaluate, jacobian=nz)
        self.evaluate = S.build_function(evaluate, x, u, w)
        self.jacobian = S.build_function(nz, x, u, w)
        self.num_state, self.num_action, self.num_parameter = num_state, num_action, num_parameter
        self.num_constraint = len(evaluate)
        self.num_jacobian = len(nz)
        self.jacobian_sparsity = [rows, cols]
        if evaluate_hessian:
            lam = S.variables("λ", self.num_constraint)
            lag = _dot(lam, evaluate)
            hr, hc, hnz = S.sparsehessian(lag, xu)
            self.hessian = S.build_function(hnz, x, u, w, lam)
            self.hessian_sparsity = [hr, hc]
            self.num_hessian = len(hnz)
            self.sym.update(lam=lam, hessian=hnz)
        else:
            self.hessian = None
            self.hessian_sparsity = [[]]
            self.num_hessian = 0
        self.evaluate_cache = np.zeros(self.num_constraint)
        self.jacobian_cache = np.zeros(self.num_jacobian)
        self.hessian_cache = np.zeros(self.num_hessian)


def constraints_stage(violations, indices, constraints, states, actions, parameters):
    """src/constraints.jl:80-86"""
    for t, con in enumerate(constraints):
        con.evaluate(con.evaluate_cache, states[t], actions[t], parameters[t])
        violations[np.asarray(indices[t], dtype=np.int64) - 1] = con.evaluate_cache
        con.evaluate_cache[:] = 0.0


def jacobian_stage(jacobians, indices, constraints, states, actions, parameters):
    """src/constraints.jl:88-94"""
    for t, con in enumerate(constraints):
        con.jacobian(con.jacobian_cache, states[t], actions[t], parameters[t])
        jacobians[np.asarray(indices[t], dtype=np.int64) - 1] = con.jacobian_cache
        con.jacobian_cache[:] = 0.0


def hessian_lagrangian_stage(hessians, indices, constraints, states, actions, parameters, duals):
    """src/constraints.jl:96-104"""
    for t, con in enumerate(constraints):
        if len(con.hessian_cache) > 0:
            con.hessian(con.hessian_cache, states[t], actions[t], parameters[t], duals[t])
            idx = np.asarray(indices[t], dtype=np.int64) - 1
            for k in range(len(idx)):
                hessians[idx[k]] += con.hessian_cache[k]
            con.hessian_cache[:] = 0.0


def sparsity_jacobian_stage(constraints, num_state, num_action, row_shift=0):
    """src/constraints.jl:106-120"""
    row, col = [], []
    for t, con in enumerate(constraints):
        cs = _shift(t, num_state, num_action)
        row.extend(r + row_shift for r in con.jacobian_sparsity[0])
        col.extend(c + cs for c in con.jacobian_sparsity[1])
        row_shift += con.num_constraint
    return list(zip(row, col))


def sparsity_hessian_stage(constraints, num_state, num_action):
    """src/constraints.jl:122-135"""
    row, col = [], []
    for t, con in enumerate(constraints):
        if len(con.hessian_sparsity[0]) > 0:
            sh = _shift(t, num_state, num_action)
            row.extend(r + sh for r in con.hessian_sparsity[0])
            col.extend(c + sh for c in con.hessian_sparsity[1])
    return list(zip(row, col))


def num_constraint_stage(constraints):
    return sum(c.num_constraint for c in constraints)


def num_jacobian_stage(constraints):
    return sum(c.num_jacobian for c in constraints)


def constraint_indices_stage(constraints, shift=0):
    """src/constraints.jl:141-152"""
    out = []
    for con in constraints:
        out.append([shift + i for i in range(1, con.num_constraint + 1)])
        shift += con.num_constraint
    return out


def jacobian_indices_stage(constraints, shift=0):
    """src/constraints.jl:154-165"""
    out = []
    for con in constraints:
        out.append([shift + i for i in range(1, con.num_jacobian + 1)])
        shift += con.num_jacobian
    return out


def hessian_indices_stage(constraints, key, num_state, num_action):
    """src/constraints.jl:167-183"""
    indices = []
    for t, con in enumerate(constraints):
        if len(con.hessian_sparsity[0]) > 0:
            sh = _shift(t, num_state, num_action)
            rc = [(r + sh, c + sh) for r, c in zip(con.hessian_sparsity[0], con.hessian_sparsity[1])]
            indices.append(_findfirst(key, rc))
        else:
            indices.append([])
    return indices


# ------------------------------------------------------------------ GeneralConstraint
class GeneralConstraint:
    """src/general_constraint.jl:1-71. One block over the whole z and the flat
    parameter vector. Hessian closure built as f!(out, z, w, λ) (:35); the reference
    calls it with the wrong arity (:87, Q7) -- the oracle implements the intended
    (z, w, λ) semantics."""

    def __init__(self, f: Optional[Callable] = None, num_variables: int = 0, num_parameter: int = 0,
                 indices_inequality: Sequence[int] = (), evaluate_hessian: bool = False):
        self.indices_inequality = list(indices_inequality)
        if f is None:
            self.evaluate = lambda out, z, w: None
            self.jacobian = lambda out, z, w: None
            self.hessian = lambda out, z, w, lam=None: None
            self.num_variables = self.num_parameter = 0
            self.num_constraint = self.num_jacobian = self.num_hessian = 0
            self.jacobian_sparsity = [[], []]
            self.hessian_sparsity = [[], []]
            self.evaluate_cache = np.zeros(0)
            self.jacobian_cache = np.zeros(0)
            self.hessian_cache = np.zeros(0)
            self.sym = None
            return
        z = S.variables("z", num_variables)
        w = S.variables("w", num_parameter)
        evaluate = _vec(f(z, w))
        rows, cols, nz = S.sparsejacobian(evaluate, list(z))
        self.sym = dict(z=z, w=w, evaluate=evaluate, jacobian=nz)
        self.evaluate = S.build_function(evaluate, z, w)
        self.jacobian = S.build_function(nz, z, w)
        self.num_variables, self.num_parameter = num_variables, num_parameter
        self.num_constraint = len(evaluate)
        self.num_jacobian = len(nz)
        self.jacobian_sparsity = [rows, cols]
        if evaluate_hessian:
            lam = S.variables("λ", self.num_constraint)
            lag = _dot(lam, evaluate)
            hr, hc, hnz = S.sparsehessian(lag, list(z))
            self.hessian = S.build_function(hnz, z, w, lam)
            self.hessian_sparsity = [hr, hc]
            self.num_hessian = len(hnz)
            self.sym.update(lam=lam, hessian=hnz)
        else:
            self.hessian = None
            self.hessian_sparsity = [[]]
            self.num_hessian = 0
        self.evaluate_cache = np.zeros(self.num_constraint)
        self.jacobian_cache = np.zeros(self.num_jacobian)
        self.hessian_cache = np.zeros(self.num_hessian)


def constraints_general(violations, indices, general, variables, parameters):
    """src/general_constraint.jl:73-77"""
    general.evaluate(general.evaluate_cache, variables, parameters)
    violations[np.asarray(indices, dtype=np.int64) - 1] = general.evaluate_cache
    general.evaluate_cache[:] = 0.0


def jacobian_general(jacobians, indices, general, variables, parameters):
    """src/general_constraint.jl:79-83"""
    general.jacobian(general.jacobian_cache, variables, parameters)
    jacobians[np.asarray(indices, dtype=np.int64) - 1] = general.jacobian_cache
    general.jacobian_cache[:] = 0.0


def hessian_lagrangian_general(hessian, indices, general, variables, parameters, duals):
    """src/general_constraint.jl:85-91 with the intended arity (Q7)."""
    if len(general.hessian_cache) > 0:
        general.hessian(general.hessian_cache, variables, parameters, duals)
        idx = np.asarray(indices, dtype=np.int64) - 1
        for k in range(len(idx)):
            hessian[idx[k]] += general.hessian_cache[k]
        general.hessian_cache[:] = 0.0


def sparsity_jacobian_general(general, num_variables, row_shift=0):
    """src/general_constraint.jl:93-103"""
    return list(zip([r + row_shift for r in general.jacobian_sparsity[0]],
                    list(general.jacobian_sparsity[1])))


def sparsity_hessian_general(general, num_variables):
    """src/general_constraint.jl:105-116"""
    if len(general.hessian_sparsity[0]) > 0:
        return list(zip(general.hessian_sparsity[0], general.hessian_sparsity[1]))
    return []


def hessian_indices_general(general, key, num_variables):
    """src/general_constraint.jl:126-139"""
    if len(general.hessian_sparsity[0]) > 0:
        rc = list(zip(general.hessian_sparsity[0], general.hessian_sparsity[1]))
        return _findfirst(key, rc)
    return []


# ------------------------------------------------------------------ Bound
class Bound:
    """src/bounds.jl:1-16"""

    def __init__(self, num_state: int = 0, num_action: int = 0, state_lower=None, state_upper=None,
                 action_lower=None, action_upper=None):
        self.state_lower = np.full(num_state, -np.inf) if state_lower is None else np.asarray(state_lower, float)
        self.state_upper = np.full(num_state, np.inf) if state_upper is None else np.asarray(state_upper, float)
        self.action_lower = np.full(num_action, -np.inf) if action_lower is None else np.asarray(action_lower, float)
        self.action_upper = np.full(num_action, np.inf) if action_upper is None else np.asarray(action_upper, float)
