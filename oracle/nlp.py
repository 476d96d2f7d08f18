"""ORACLE (test infrastructure only). CPU restatement of problem assembly and the
five MOI callbacks:

  TrajectoryOptimizationData / indices / NLPData  /root/reference/src/data.jl:1-220
  trajectory!, duals!                              /root/reference/src/data.jl:258-278
  MOI.eval_* + structure methods                   /root/reference/src/moi.jl:1-125
  Solver (constructor defaults only)               /root/reference/src/solver.jl:6-21

Serial per-knot loops, one problem at a time; a batch is a python loop over
problems. Index lists are 1-based (Julia) everywhere.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import elements as E


class TrajectoryOptimizationData:
    """src/data.jl:1-40"""

    def __init__(self, objective, dynamics, constraints, bounds, parameters=None):
        sd, ad, pd = E.dimensions(dynamics)
        if parameters is None:
            parameters = [np.zeros(n) for n in pd]
        self.state_dimensions, self.action_dimensions, self.parameter_dimensions = sd, ad, pd
        self.states = [np.zeros(n) for n in sd]
        self.actions = [np.zeros(n) for n in ad]
        self.parameters = [np.asarray(p, dtype=float) for p in parameters]
        self.objective, self.dynamics, self.constraints, self.bounds = objective, dynamics, constraints, bounds
        self.duals_dynamics = [np.zeros(n) for n in sd[1:]]
        self.duals_constraints = [np.zeros(c.num_constraint) for c in constraints]


class TrajectoryOptimizationIndices:
    """src/data.jl:44-104"""

    def __init__(self, objective, dynamics, constraints, general, key, num_state, num_action, num_trajectory):
        nd = E.num_constraint_dynamics(dynamics)
        njd = E.num_jacobian_dynamics(dynamics)
        self.dynamics_constraints = E.constraint_indices_dynamics(dynamics, shift=0)
        self.dynamics_jacobians = E.jacobian_indices_dynamics(dynamics, shift=0)
        self.stage_constraints = E.constraint_indices_stage(constraints, shift=nd)
        self.stage_jacobians = E.jacobian_indices_stage(constraints, shift=njd)
        ns = E.num_constraint_stage(constraints)
        njs = E.num_jacobian_stage(constraints)
        self.general_constraint = [nd + ns + i for i in range(1, general.num_constraint + 1)]
        self.general_jacobian = [njd + njs + i for i in range(1, general.num_jacobian + 1)]
        self.objective_hessians = E.hessian_indices_objective(objective, key, num_state, num_action)
        self.dynamics_hessians = E.hessian_indices_dynamics(dynamics, key, num_state, num_action)
        self.stage_hessians = E.hessian_indices_stage(constraints, key, num_state, num_action)
        self.general_hessian = E.hessian_indices_general(general, key, num_trajectory)
        self.states = E.state_indices(dynamics)
        self.actions = E.action_indices(dynamics)
        self.state_action = E.state_action_indices(dynamics)
        self.state_action_next_state = E.state_action_next_state_indices(dynamics)


def primal_bounds(bounds, num_variables, state_indices, action_indices):
    """src/data.jl:123-133"""
    lower = np.full(num_variables, -np.inf)
    upper = np.full(num_variables, np.inf)
    for t, b in enumerate(bounds):
        if len(b.state_lower) > 0:
            lower[np.asarray(state_indices[t]) - 1] = b.state_lower
        if len(b.state_upper) > 0:
            upper[np.asarray(state_indices[t]) - 1] = b.state_upper
        if len(b.action_lower) > 0:
            lower[np.asarray(action_indices[t]) - 1] = b.action_lower
        if len(b.action_upper) > 0:
            upper[np.asarray(action_indices[t]) - 1] = b.action_upper
    return lower, upper


def constraint_bounds(constraints, general, num_dynamics, num_stage, idx):
    """src/data.jl:135-148 -- inequality rows are (-Inf, 0]."""
    total = num_dynamics + num_stage + general.num_constraint
    lower, upper = np.zeros(total), np.zeros(total)
    for t, con in enumerate(constraints):
        for i in con.indices_inequality:
            lower[idx.stage_constraints[t][i - 1] - 1] = -np.inf
    for i in general.indices_inequality:
        lower[num_dynamics + num_stage + i - 1] = -np.inf
    return lower, upper


def trajectory_(states, actions, trajectory, state_indices, action_indices):
    """src/data.jl:258-267"""
    for t, idx in enumerate(state_indices):
        states[t][:] = trajectory[np.asarray(idx, dtype=np.int64) - 1]
    for t, idx in enumerate(action_indices):
        actions[t][:] = trajectory[np.asarray(idx, dtype=np.int64) - 1]


def duals_(duals_dynamics, duals_constraints, duals_general, duals, dyn_idx, con_idx, gen_idx):
    """src/data.jl:269-278"""
    for t, idx in enumerate(dyn_idx):
        duals_dynamics[t][:] = duals[np.asarray(idx, dtype=np.int64) - 1]
    for t, idx in enumerate(con_idx):
        duals_constraints[t][:] = duals[np.asarray(idx, dtype=np.int64) - 1]
    duals_general[:] = duals[np.asarray(gen_idx, dtype=np.int64) - 1]


class NLPData:
    """src/data.jl:106-121,150-220 plus the evaluator methods of src/moi.jl."""

    def __init__(self, trajopt: TrajectoryOptimizationData, evaluate_hessian: bool = False,
                 general_constraint: Optional[E.GeneralConstraint] = None):
        general = general_constraint if general_constraint is not None else E.GeneralConstraint()
        sd, ad = trajopt.state_dimensions, trajopt.action_dimensions
        total_variables = sum(sd) + sum(ad)
        num_dynamics = E.num_constraint_dynamics(trajopt.dynamics)
        num_stage = E.num_constraint_stage(trajopt.constraints)
        num_general = general.num_constraint
        total_constraints = num_dynamics + num_stage + num_general
        total_jacobians = (E.num_jacobian_dynamics(trajopt.dynamics) + E.num_jacobian_stage(trajopt.constraints)
                           + general.num_jacobian)
        sp_dyn = E.sparsity_jacobian_dynamics(trajopt.dynamics, sd, ad, row_shift=0)
        sp_con = E.sparsity_jacobian_stage(trajopt.constraints, sd, ad, row_shift=num_dynamics)
        sp_gen = E.sparsity_jacobian_general(general, total_variables, row_shift=num_dynamics + num_stage)
        jacobian_sparsity = sp_dyn + sp_con + sp_gen
        hl = (E.sparsity_hessian_objective(trajopt.objective, sd, ad)
              + E.sparsity_hessian_dynamics(trajopt.dynamics, sd, ad)
              + E.sparsity_hessian_stage(trajopt.constraints, sd, ad)
              + E.sparsity_hessian_general(general, total_variables))
        key = sorted(set(hl))  # sort(unique(...)): lexicographic on (row, col)
        self.trajopt = trajopt
        self.num_variables = total_variables
        self.num_constraint = total_constraints
        self.num_jacobian = total_jacobians
        self.num_hessian_lagrangian = len(hl)  # NON-unique length (Q4)
        self.indices = TrajectoryOptimizationIndices(trajopt.objective, trajopt.dynamics, trajopt.constraints,
                                                     general, key, sd, ad, total_variables)
        self.variable_bounds = list(primal_bounds(trajopt.bounds, total_variables, self.indices.states,
                                                  self.indices.actions))
        self.constraint_bounds = list(constraint_bounds(trajopt.constraints, general, num_dynamics, num_stage,
                                                        self.indices))
        self.jacobian_sparsity = jacobian_sparsity
        self.hessian_lagrangian_sparsity = key
        self.hessian_lagrangian = evaluate_hessian
        self.general_constraint = general
        self.parameters = (np.concatenate([np.asarray(p, float).reshape(-1) for p in trajopt.parameters])
                           if len(trajopt.parameters) else np.zeros(0))
        self.duals_general = np.zeros(general.num_constraint)

    # ---- MOI callbacks (src/moi.jl) ----
    def _unpack(self, z):
        trajectory_(self.trajopt.states, self.trajopt.actions, z, self.indices.states, self.indices.actions)

    def eval_objective(self, z) -> float:
        """src/moi.jl:1-13"""
        self._unpack(z)
        t = self.trajopt
        return E.cost(t.objective, t.states, t.actions, t.parameters)

    def eval_objective_gradient(self, gradient, z) -> None:
        """src/moi.jl:15-30"""
        gradient[:] = 0.0
        self._unpack(z)
        t = self.trajopt
        E.gradient_(gradient, self.indices.state_action, t.objective, t.states, t.actions, t.parameters)

    def eval_constraint(self, violations, z) -> None:
        """src/moi.jl:32-50"""
        violations[:] = 0.0
        self._unpack(z)
        t = self.trajopt
        E.constraints_dynamics(violations, self.indices.dynamics_constraints, t.dynamics, t.states, t.actions,
                               t.parameters)
        if len(self.indices.stage_constraints) > 0:
            E.constraints_stage(violations, self.indices.stage_constraints, t.constraints, t.states, t.actions,
                                t.parameters)
        if self.general_constraint.num_constraint != 0:
            E.constraints_general(violations, self.indices.general_constraint, self.general_constraint, z,
                                  self.parameters)

    def eval_constraint_jacobian(self, jacobian, z) -> None:
        """src/moi.jl:52-70"""
        jacobian[:] = 0.0
        self._unpack(z)
        t = self.trajopt
        E.jacobian_dynamics(jacobian, self.indices.dynamics_jacobians, t.dynamics, t.states, t.actions, t.parameters)
        if len(self.indices.stage_jacobians) > 0:
            E.jacobian_stage(jacobian, self.indices.stage_jacobians, t.constraints, t.states, t.actions,
                             t.parameters)
        if self.general_constraint.num_constraint != 0:
            E.jacobian_general(jacobian, self.indices.general_jacobian, self.general_constraint, z, self.parameters)

    def eval_hessian_lagrangian(self, hessian, z, scaling, duals) -> None:
        """src/moi.jl:72-120: zero; cost (scaled); dynamics; stage; general -- in that order."""
        hessian[:] = 0.0
        self._unpack(z)
        t = self.trajopt
        duals_(t.duals_dynamics, t.duals_constraints, self.duals_general, duals, self.indices.dynamics_constraints,
               self.indices.stage_constraints, self.indices.general_constraint)
        E.hessian_(hessian, self.indices.objective_hessians, t.objective, t.states, t.actions, t.parameters, scaling)
        E.hessian_lagrangian_dynamics(hessian, self.indices.dynamics_hessians, t.dynamics, t.states, t.actions,
                                      t.parameters, t.duals_dynamics)
        E.hessian_lagrangian_stage(hessian, self.indices.stage_hessians, t.constraints, t.states, t.actions,
                                   t.parameters, t.duals_constraints)
        E.hessian_lagrangian_general(hessian, self.indices.general_hessian, self.general_constraint, z,
                                     self.parameters, self.duals_general)

    def features_available(self):
        """src/moi.jl:122"""
        return ["Grad", "Jac", "Hess"] if self.hessian_lagrangian else ["Grad", "Jac"]

    def jacobian_structure(self):
        """src/moi.jl:124"""
        return self.jacobian_sparsity

    def hessian_lagrangian_structure(self):
        """src/moi.jl:125"""
        return self.hessian_lagrangian_sparsity


class Solver:
    """src/solver.jl:1-21 -- constructor chain only (Ipopt is absent here). `parameters`
    defaults to T empty vectors plus one extra (Q10)."""

    def __init__(self, dynamics, objective, constraints, bounds, evaluate_hessian=False,
                 general_constraint=None, options=None, parameters=None):
        if parameters is None:
            parameters = [np.zeros(n) for n in E.dimensions(dynamics)[2]] + [np.zeros(0)]
        trajopt = TrajectoryOptimizationData(objective, dynamics, constraints, bounds, parameters=parameters)
        self.nlp = NLPData(trajopt, evaluate_hessian=evaluate_hessian, general_constraint=general_constraint)

    def set_parameters(self, parameters):
        """Batched use: re-point the per-knot parameter vectors (and the flat copy the
        general constraint sees, src/data.jl:218) at another problem's data."""
        self.nlp.trajopt.parameters = [np.asarray(p, float) for p in parameters]
        self.nlp.parameters = (np.concatenate([p.reshape(-1) for p in self.nlp.trajopt.parameters])
                               if len(parameters) else np.zeros(0))

    def get_trajectory(self):
        """src/solver.jl:41-43"""
        return self.nlp.trajopt.states, self.nlp.trajopt.actions[:-1]


def linear_interpolation(initial_state, final_state, horizon):
    """src/utils.jl:1-10"""
    n = len(initial_state)
    X = [np.array(initial_state, dtype=float).copy() for _ in range(horizon)]
    for t in range(horizon):
        for i in range(n):
            X[t][i] = (final_state[i] - initial_state[i]) / (horizon - 1) * t + initial_state[i]
    return X
