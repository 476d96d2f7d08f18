"""GPU test of solve! through the lock-step driver: the pendulum swing-up example
(/root/reference/examples/pendulum/pendulum.jl:1-98) solved for a small batch with the CUDA callbacks
underneath, against the same solver with the CPU oracle's callbacks underneath. The solver above is
SciPy's trust-constr in both arms (Ipopt is absent), so what is compared is what the north star asks
of the callbacks: identical iterates (1e-8) and final objectives (1e-8)."""
import math

import numpy as np
import pytest

import dto_b200 as D
from dto_b200 import driver
from examples import models as M
from oracle import api as O

from util import OracleBatch

pytestmark = pytest.mark.gpu


def test_pendulum_solve_matches_oracle_driven_solve():
    B = 4
    mo, mp = M.BUILDERS["pendulum"](O), M.BUILDERS["pendulum"](D)
    osolver = O.solver_from(mo)
    psolver = D.solver_from(mp, batch=B)
    T = mp["T"]
    # initialize_states! / initialize_controls! (src/solver.jl:23-39), one guess per problem
    for b in range(B):
        psolver.initialize_states(D.linear_interpolation(mp["x1"], mp["xT"] * (1.0 + 0.05 * b), T), problem=b)
        psolver.initialize_controls([np.array([0.01 * (b + 1)]) for _ in range(T - 1)], problem=b)
    z0 = psolver._initial.copy()
    opts = {"maxiter": 300}
    res = psolver.solve(options=opts, record_iterates=True, method="broker")
    n0 = psolver.nlp.launch_count()
    assert n0 > 0, "no CUDA kernel was launched by solve!"
    Zo, reso, _, its_o = driver.solve_batch(OracleBatch(osolver, B), z0, options=opts, record_iterates=True)
    c = np.zeros(osolver.nlp.num_constraint)
    for b in range(B):
        assert res[b].status in (1, 2), res[b].message
        assert res[b].nit == reso[b].nit
        its = psolver.iterates[b]
        assert len(its) == len(its_o[b])
        for k, (x, y) in enumerate(zip(its, its_o[b])):
            assert np.max(np.abs(x - y)) <= 1e-8 * max(1.0, np.max(np.abs(y))), f"problem {b}: iterate {k} differs"
        assert abs(res[b].fun - reso[b].fun) <= 1e-8 * max(1.0, abs(reso[b].fun))
        states, actions = psolver.get_trajectory(b)          # src/solver.jl:41-43
        assert len(states) == T and len(actions) == T - 1
        assert np.allclose(states[0], mp["x1"], atol=1e-6) and np.allclose(states[-1], [math.pi, 0.0], atol=1e-6)
        z = np.concatenate([np.concatenate([states[t], actions[t]]) for t in range(T - 1)] + [states[-1]])
        osolver.nlp.eval_constraint(c, z)
        assert np.max(np.abs(c)) < 1e-6
    # batching: far fewer batched GPU calls than per-problem callbacks
    br = psolver.broker
    assert sum(br.batched_calls.values()) < 0.5 * sum(br.requests.values())
    psolver.nlp.close()


def test_reference_solve_test_with_user_provided_dynamics_gradients():
    """/root/reference/test/solve.jl:140-225 verbatim: double integrator whose dynamics Jacobian is supplied by the user
    (second Dynamics constructor), no Hessians (the reference's Ipopt then runs in limited-memory mode), end points pinned
    by bounds, states interpolated and controls ~ N(0, 1) as the guess. Solver.solve() routes it to the lock-step driver
    (no :Hess feature => quasi-Newton stand-in solver, bounds handled); the test's own acceptance must hold for every
    problem of the batch and the Hessian callback must never be asked for."""
    B = 3
    mp = M.build_user_jacobian(D)
    s = D.solver_from(mp, batch=B)
    assert s.nlp.hessian_lagrangian is False
    T = mp["T"]
    rng = np.random.default_rng(11)
    s.initialize_states(D.linear_interpolation(mp["x1"], mp["xT"], T))
    for b in range(B):
        s.initialize_controls([rng.normal(size=1) for _ in range(T - 1)], problem=b)
    res = s.solve(options={"maxiter": 500})
    assert s.broker is not None and s.broker.requests["H"] == 0 and s.nlp.launch_count() > 0
    for b in range(B):
        xs, us = s.get_trajectory(b)
        assert np.linalg.norm(xs[0] - mp["x1"]) < 1.0e-3 and np.linalg.norm(xs[-1] - mp["xT"]) < 1.0e-3, (b, res[b].message)
    s.nlp.close()
