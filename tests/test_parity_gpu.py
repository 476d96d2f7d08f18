"""GPU parity: every callback of the CUDA path, called through the C ABI (libdto.so), against
the CPU oracle on identical seeded (z, lambda, sigma, w). Tolerance: 1e-12 relative or 1e-14
absolute (BASELINE.json north_star); structures bit-exact."""
import numpy as np
import pytest

import dto_b200 as D
from examples import models as M
from oracle import api as O

from util import assert_close, make_inputs, oracle_eval_all

pytestmark = pytest.mark.gpu

CASES = [
    ("pendulum", dict(), 5, 1),
    ("cartpole", dict(T=101), 3, 2),
    ("cartpole", dict(T=51, parameterized=False), 2, 5),
    ("acrobot", dict(T=101), 3, 3),
    ("car", dict(T=201, obstacle="general"), 2, 4),
    ("car", dict(T=51, obstacle="stage"), 3, 4),
    ("acrobot_hessian_test", dict(), 40, 6),
    ("linear_general", dict(), 3, 7),
    ("piecewise", dict(), 64, 8),
]


@pytest.mark.parametrize("name,kw,B,config", CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(CASES)])
def test_callbacks_match_oracle(name, kw, B, config):
    mo = M.BUILDERS[name](O, **kw)
    mp = M.BUILDERS[name](D, **kw)
    osolver = O.solver_from(mo)
    psolver = D.solver_from(mp, batch=B)
    on, pn = osolver.nlp, psolver.nlp
    assert pn.jacobian_structure() == on.jacobian_structure()
    assert pn.hessian_lagrangian_structure() == on.hessian_lagrangian_structure()
    z, lam, sigma, w = make_inputs(name, mp, pn.num_variables, pn.num_constraint, pn.num_parameter, B, config)
    ref = oracle_eval_all(osolver, mo, z, lam, sigma, w)
    if pn.num_parameter:
        pn.set_parameters(w)
    f = pn.eval_objective(z)
    g = np.full((B, pn.num_variables), np.nan)
    c = np.full((B, pn.num_constraint), np.nan)
    J = np.full((B, pn.num_jacobian), np.nan)
    H = np.full((B, pn.num_hessian), np.nan)
    pn.eval_objective_gradient(g, z)
    pn.eval_constraint(c, z)
    pn.eval_constraint_jacobian(J, z)
    pn.eval_hessian_lagrangian(H, z, sigma, lam)
    assert_close("objective", f, ref["f"])
    assert_close("gradient", g, ref["g"])
    assert_close("constraint", c, ref["c"])
    assert_close("jacobian", J, ref["J"])
    assert_close("hessian", H, ref["H"])
    # fused pass must reproduce the separate passes (to rounding: separately compiled programs)
    J2 = np.full_like(J, np.nan)
    H2 = np.full_like(H, np.nan)
    pn.eval_jacobian_hessian(J2, H2)
    assert np.array_equal(J2, J) or np.allclose(J2, J, rtol=1e-13, atol=1e-15)
    assert_close("fused jacobian", J2, ref["J"])
    assert_close("fused hessian", H2, ref["H"])
    pn.close()


@pytest.mark.parametrize("nonlinear", [False, True])
def test_user_jacobian_dynamics_matches_oracle(nonlinear):
    """Second Dynamics constructor (/root/reference/src/dynamics.jl:59-101, test/solve.jl:140-225): the USER's
    Jacobian expressions are what gets evaluated (dense column-major slots, no Hessian). With nonlinear=True the
    supplied Jacobian is deliberately not the derivative of f, so "differentiated f instead" would fail here."""
    kw = dict(nonlinear=nonlinear)
    mo, mp = M.build_user_jacobian(O, **kw), M.build_user_jacobian(D, **kw)
    B = 33
    osolver, psolver = O.solver_from(mo), D.solver_from(mp, batch=B)
    on, pn = osolver.nlp, psolver.nlp
    assert pn.jacobian_structure() == on.jacobian_structure() and pn.num_jacobian == 100
    assert pn.features_available() == ["Grad", "Jac"]  # src/moi.jl:122 without :Hess
    z, lam, sigma, w = make_inputs("user_jacobian", mp, pn.num_variables, pn.num_constraint, 0, B, 9)
    ref = oracle_eval_all(osolver, mo, z, lam, sigma, w, hessian=False)
    g = np.full((B, pn.num_variables), np.nan)
    c = np.full((B, pn.num_constraint), np.nan)
    J = np.full((B, pn.num_jacobian), np.nan)
    assert_close("objective", pn.eval_objective(z), ref["f"])
    pn.eval_objective_gradient(g, z)
    pn.eval_constraint(c, z)
    pn.eval_constraint_jacobian(J, z)
    assert_close("gradient", g, ref["g"])
    assert_close("constraint", c, ref["c"])
    assert_close("jacobian", J, ref["J"])
    if nonlinear:  # the scaled entry really is the user's, not d f / d x
        exact = -0.1 * np.cos(z[:, 0]) * z[:, 2]
        assert np.allclose(J[:, 1], 3.0 * exact, rtol=1e-12) and not np.allclose(J[:, 1], exact, rtol=1e-3)
    pn.close()


def test_heterogeneous_problem_matches_oracle():
    """Dims, element kinds and parameter lengths change from knot to knot, stage constraints exist on
    some knots only, and a nonlinear GeneralConstraint couples distant knots (reference data model:
    SURVEY Q1, Q7, Q10, Q11). Parameters use the reference's vcat(parameters...) layout."""
    mo, mp = M.build_heterogeneous(O), M.build_heterogeneous(D)
    B = 37
    r = np.random.default_rng(99)
    nw = mo["nw"]
    psolver = D.Solver(mp["dynamics"], mp["objective"], mp["constraints"], mp["bounds"], evaluate_hessian=True,
                       general_constraint=mp["general"], batch=B, name="heterogeneous")
    pn = psolver.nlp
    assert pn.num_parameter == sum(nw)
    z = r.uniform(-1, 1, (B, pn.num_variables))
    lam = r.normal(size=(B, pn.num_constraint))
    sigma = r.uniform(0.2, 2.0, B)
    w = r.uniform(-1, 1, (B, pn.num_parameter))
    pn.set_parameters(w)
    f = pn.eval_objective(z)
    g = np.full((B, pn.num_variables), np.nan)
    c = np.full((B, pn.num_constraint), np.nan)
    J = np.full((B, pn.num_jacobian), np.nan)
    H = np.full((B, pn.num_hessian), np.nan)
    pn.eval_objective_gradient(g)
    pn.eval_constraint(c)
    pn.eval_constraint_jacobian(J)
    pn.eval_hessian_lagrangian(H, None, sigma, lam)
    J2, H2 = np.full_like(J, np.nan), np.full_like(H, np.nan)
    pn.eval_jacobian_hessian(J2, H2, z, sigma, lam)
    osolver = O.solver_from(mo, parameters=[np.zeros(n) for n in nw] + [np.zeros(0)])
    on = osolver.nlp
    assert pn.jacobian_structure() == on.jacobian_structure()
    assert pn.hessian_lagrangian_structure() == on.hessian_lagrangian_structure()
    off = np.concatenate([[0], np.cumsum(nw)])
    for b in range(B):
        osolver.set_parameters([w[b, off[t]:off[t + 1]] for t in range(len(nw))] + [np.zeros(0)])
        assert_close("f", f[b], on.eval_objective(z[b]))
        ref = np.zeros(pn.num_variables); on.eval_objective_gradient(ref, z[b]); assert_close("g", g[b], ref)
        ref = np.zeros(pn.num_constraint); on.eval_constraint(ref, z[b]); assert_close("c", c[b], ref)
        ref = np.zeros(pn.num_jacobian); on.eval_constraint_jacobian(ref, z[b]); assert_close("J", J[b], ref)
        assert_close("J fused", J2[b], ref)
        ref = np.zeros(pn.num_hessian); on.eval_hessian_lagrangian(ref, z[b], float(sigma[b]), lam[b])
        assert_close("H", H[b], ref)
        assert_close("H fused", H2[b], ref)
    pn.close()
