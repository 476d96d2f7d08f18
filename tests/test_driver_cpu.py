"""CPU tests of the lock-step batched solve driver (reference wiring src/data.jl:222-255,
src/solver.jl:45-47): the rendezvous logic is exercised with the CPU oracle as the batched evaluator
(test infrastructure) under SciPy's trust-constr -- Ipopt is not available in this image."""
import math

import numpy as np
import pytest

import dto_b200 as D  # noqa: F401  (the driver lives in the product package)
from dto_b200 import driver
from examples import models as M
from oracle import api as O

from util import OracleBatch


def _pendulum_guesses(mo, nlp, B):
    T, n, m = mo["T"], mo["n"], mo["m"]
    z0 = np.zeros((B, nlp.num_variables))
    for b in range(B):
        xs = O.linear_interpolation(mo["x1"], mo["xT"] * (1.0 + 0.05 * b), T)
        for t in range(T):
            o = t * (n + m)
            z0[b, o:o + n] = xs[t]
            if t < T - 1:
                z0[b, o + n:o + n + m] = 0.01 * (b + 1)
    return z0


def test_lockstep_equals_solo_and_batches_calls():
    mo = M.BUILDERS["pendulum"](O)
    osolver = O.solver_from(mo)
    B = 3
    nlp = OracleBatch(osolver, B)
    z0 = _pendulum_guesses(mo, nlp, B)
    opts = {"maxiter": 300}
    Z, res, broker, its = driver.solve_batch(nlp, z0, options=opts, record_iterates=True)
    assert Z.shape == (B, nlp.num_variables)
    c = np.zeros(nlp.num_constraint)
    for b in range(B):
        assert res[b].status in (1, 2), res[b].message          # gtol or xtol termination
        osolver.nlp.eval_constraint(c, Z[b])
        assert np.max(np.abs(c)) < 1e-6                          # dynamics + endpoint constraints hold
        # swing-up reached: last state = [pi, 0]
        assert abs(Z[b, -2] - math.pi) < 1e-6 and abs(Z[b, -1]) < 1e-6
    # batching happened: each flush serves every live problem with one call per waiting kind
    total_req = sum(broker.requests.values())
    total_calls = sum(broker.batched_calls.values())
    assert total_calls < total_req and broker.flushes <= total_calls
    assert not broker.live
    # lock step does not change any problem's iterates: solve problem b alone and compare bit for bit
    for b in range(B):
        solo = OracleBatch(osolver, 1)
        Zs, rs, _, its_s = driver.solve_batch(solo, z0[b:b + 1], options=opts, record_iterates=True)
        assert rs[0].nit == res[b].nit
        assert np.array_equal(Zs[0], Z[b])
        assert len(its_s[0]) == len(its[b]) and all(np.array_equal(x, y) for x, y in zip(its_s[0], its[b]))


def test_error_in_batched_call_reaches_every_task():
    mo = M.BUILDERS["pendulum"](O)
    nlp = OracleBatch(O.solver_from(mo), 2)

    def boom(G, Z):
        raise FloatingPointError("evaluator failed")

    nlp.eval_objective_gradient = boom
    with pytest.raises(Exception) as e:
        driver.solve_batch(nlp, _pendulum_guesses(mo, nlp, 2), options={"maxiter": 5})
    assert "failed" in str(e.value)


def test_solve_without_hessian_uses_quasi_newton():
    """evaluate_hessian=False: features_available has no :Hess (src/moi.jl:122) and the Hessian callback must never
    be requested; the stand-in solver runs on BFGS updates, like Ipopt's limited-memory mode."""
    mo = M.BUILDERS["pendulum"](O, evaluate_hessian=False)
    osolver = O.solver_from(mo)
    assert ":Hess" not in [str(f) for f in osolver.nlp.features_available()] and "Hess" not in str(osolver.nlp.features_available())
    nlp = OracleBatch(osolver, 2)
    assert nlp.hessian_lagrangian is False
    Z, res, broker, _ = driver.solve_batch(nlp, _pendulum_guesses(mo, nlp, 2), options={"maxiter": 500})
    assert broker.requests["H"] == 0
    c = np.zeros(nlp.num_constraint)
    for b in range(2):
        osolver.nlp.eval_constraint(c, Z[b])
        assert np.max(np.abs(c)) < 1e-5
