"""The reference's own tests (test/objective.jl, test/dynamics.jl, test/constraints.jl,
test/hessian_lagrangian.jl) restated against the CPU oracle. These are the known-answer checks
that pin the oracle; tolerances are the reference's (1e-8) unless a closed form allows exactness."""
import math

import numpy as np
import pytest
import sympy as sp

from examples import models as M
from oracle import api as O
from oracle import elements as E
from oracle import nlp as N
from oracle import symbolics as S


def test_objective_known_answers():
    """test/objective.jl:1-38"""
    T, n, m = 3, 2, 1
    ot = lambda x, u, w: M.dot(x, x) + 0.1 * M.dot(u, u)
    oT = lambda x, u, w: 10.0 * M.dot(x, x)
    ct, cT = O.Cost(ot, n, m), O.Cost(oT, n, 0)
    objective = [ct] * (T - 1) + [cT]
    x1, u1, w1 = np.ones(n), np.ones(m), np.zeros(0)
    X, U, W = [x1] * T, [u1] * (T - 1) + [np.zeros(0)], [w1] * T
    ct.evaluate(ct.evaluate_cache, x1, u1, w1)
    ct.gradient(ct.gradient_cache, x1, u1, w1)
    assert ct.evaluate_cache[0] == ot(x1, u1, w1)
    assert np.linalg.norm(ct.gradient_cache - np.concatenate([2.0 * x1, 0.2 * u1])) < 1e-8
    cT.evaluate(cT.evaluate_cache, x1, u1, w1)
    cT.gradient(cT.gradient_cache, x1, u1, w1)
    assert cT.evaluate_cache[0] == oT(x1, u1, w1)
    assert np.linalg.norm(cT.gradient_cache - 20.0 * x1) < 1e-8
    J = E.cost(objective, X, U, W)
    assert abs(J - sum(ot(X[t], U[t], W[t]) for t in range(T - 1)) - oT(X[T - 1], U[T - 1], W[T - 1])) < 1e-12
    idx_xu = [[(t * (n + m)) + i for i in range(1, n + (0 if t == T - 1 else m) + 1)] for t in range(T)]
    g = np.zeros((T - 1) * (n + m) + n)
    E.gradient_(g, idx_xu, objective, X, U, W)
    assert np.linalg.norm(g - np.concatenate([np.concatenate([2.0 * x1, 0.2 * u1])] * (T - 1) + [20.0 * x1])) < 1e-8


def test_dynamics_jacobian_and_layout():
    """test/dynamics.jl:1-84 (ForwardDiff replaced by exact sympy derivatives)."""
    T, n, m = 3, 2, 1
    dt = O.Dynamics(M.test_euler_implicit, n, n, m)
    dynamics = [dt] * (T - 1)
    x1, u1, w1 = np.ones(n), np.ones(m), np.zeros(0)
    X, U, W = [x1] * T, [u1] * T, [w1] * T
    dt.evaluate(dt.evaluate_cache, x1, x1, u1, w1)
    ref = np.array([float(v) for v in M.test_euler_implicit(x1, x1, u1, w1)])
    assert np.linalg.norm(dt.evaluate_cache - ref) < 1e-8
    dt.jacobian(dt.jacobian_cache, x1, x1, u1, w1)
    jac_dense = np.zeros((n, n + m + n))
    for i, ji in enumerate(dt.jacobian_cache):
        jac_dense[dt.jacobian_sparsity[0][i] - 1, dt.jacobian_sparsity[1][i] - 1] = ji
    s = dt.sym
    vars_ = list(s["x"]) + list(s["u"]) + list(s["y"])
    sub = {v: 1.0 for v in vars_}
    exact = np.array([[float(sp.diff(e, v).evalf(30, subs=sub)) for v in vars_] for e in s["evaluate"]])
    assert np.linalg.norm(jac_dense - exact) < 1e-8
    idx_dyn, idx_jac = E.constraint_indices_dynamics(dynamics), E.jacobian_indices_dynamics(dynamics)
    d = np.zeros(E.num_constraint_dynamics(dynamics))
    j = np.zeros(E.num_jacobian_dynamics(dynamics))
    E.constraints_dynamics(d, idx_dyn, dynamics, X, U, W)
    assert np.linalg.norm(d - np.concatenate([ref] * (T - 1))) < 1e-8
    E.jacobian_dynamics(j, idx_jac, dynamics, X, U, W)
    sd, ad, _ = E.dimensions(dynamics)
    sp_ = E.sparsity_jacobian_dynamics(dynamics, sd, ad)
    dense = np.zeros((E.num_constraint_dynamics(dynamics), E.num_state_action_next_state(dynamics)))
    for i, ji in enumerate(j):
        dense[sp_[i][0] - 1, sp_[i][1] - 1] = ji
    blk = np.zeros_like(dense)  # [J 0; 0 J] with knot-to-knot column shift n+m
    blk[0:n, 0:2 * n + m] = exact
    blk[n:2 * n, n + m:n + m + 2 * n + m] = exact
    assert np.linalg.norm(dense - blk) < 1e-8
    # trajectory! round trip through state/action indices (test/dynamics.jl:62-81)
    x_idx, u_idx = E.state_indices(dynamics), E.action_indices(dynamics)
    nz = T * n + (T - 1) * m
    z = np.random.default_rng(0).uniform(size=nz)
    x = [np.zeros(n) for _ in range(T)]
    u = [np.zeros(m) for _ in range(T - 1)] + [np.zeros(0)]
    N.trajectory_(x, u, z, x_idx, u_idx)
    zb = np.zeros(nz)
    for t, idx in enumerate(x_idx):
        zb[np.asarray(idx) - 1] = x[t]
    for t, idx in enumerate(u_idx):
        zb[np.asarray(idx) - 1] = u[t]
    assert np.array_equal(z, zb)


def test_stage_constraints_known_jacobian():
    """test/constraints.jl:1-45"""
    T, n, m = 5, 2, 1
    r = np.random.default_rng(1)
    x = [r.uniform(size=n) for _ in range(T)]
    u = [r.uniform(size=m) for _ in range(T - 1)] + [np.zeros(0)]
    w = [np.zeros(0) for _ in range(T)]
    ct = lambda x, u, w: M.cat(-np.ones(n) - x, x - np.ones(n))
    cT = lambda x, u, w: x
    cont = O.Constraint(ct, n, m, indices_inequality=list(range(1, 2 * n + 1)))
    conT = O.Constraint(cT, n, 0)
    cons = [cont] * (T - 1) + [conT]
    nc, nj = E.num_constraint_stage(cons), E.num_jacobian_stage(cons)
    idx_c, idx_j = E.constraint_indices_stage(cons), E.jacobian_indices_stage(cons)
    c, j = np.zeros(nc), np.zeros(nj)
    E.constraints_stage(c, idx_c, cons, x, u, w)
    ref = np.concatenate([np.concatenate([-1 - x[t], x[t] - 1]) for t in range(T - 1)] + [x[T - 1]])
    assert np.linalg.norm(c - ref) < 1e-8
    E.jacobian_stage(j, idx_j, cons, x, u, w)
    dct = np.block([[-np.eye(n), np.zeros((n, m))], [np.eye(n), np.zeros((n, m))]])
    dc = np.zeros((nc, T * n + (T - 1) * m))
    for t in range(T - 1):
        dc[t * 2 * n:(t + 1) * 2 * n, t * (n + m):(t + 1) * (n + m)] = dct
    dc[(T - 1) * 2 * n:, (T - 1) * (n + m):] = np.eye(n)
    sp_ = E.sparsity_jacobian_stage(cons, [n] * T, [m] * (T - 1) + [0])
    dense = np.zeros_like(dc)
    for i, v in enumerate(sp_):
        dense[v[0] - 1, v[1] - 1] = j[i]
    assert np.linalg.norm(dense - dc) < 1e-8


def test_hessian_of_lagrangian_acrobot():
    """test/hessian_lagrangian.jl:1-210 -- the single most important behavioural test: index maps,
    assembled Hessian vs the dense symbolic Hessian of the hand-built Lagrangian (BOTH triangles),
    element-for-element in sorted-key order, and the full eval_hessian_lagrangian callback."""
    m_ = M.build_acrobot_hessian_test(O)
    solver = O.solver_from(m_)
    nlp = solver.nlp
    n, m = 4, 1
    np_ = 3 * n + 2 * m
    nd = 2 * n + 2 * (m + n) + n
    zs = [sp.Symbol(f"q{i}") for i in range(np_ + nd)]
    z = np.array(zs, dtype=object)
    x1, u1 = z[0:n], z[n:n + m]
    x2, u2 = z[n + m:2 * n + m], z[2 * n + m:2 * n + 2 * m]
    x3 = z[2 * n + 2 * m:3 * n + 2 * m]
    o = np_
    l1d, l2d = z[o:o + n], z[o + n:o + 2 * n]
    o += 2 * n
    l1s, l2s, l3s = z[o:o + m + n], z[o + m + n:o + 2 * (m + n)], z[o + 2 * (m + n):o + 2 * (m + n) + n]
    w0 = np.zeros(0)
    L = (m_["ot"](x1, u1, w0) + m_["ot"](x2, u2, w0) + m_["oT"](x3, np.zeros(0), w0)
         + M.dot(l1d, M.acrobot_midpoint(x2, x1, u1, w0)) + M.dot(l2d, M.acrobot_midpoint(x3, x2, u2, w0))
         + M.dot(l1s, m_["ct"](x1, u1, w0)) + M.dot(l2s, m_["ct"](x2, u2, w0)) + M.dot(l3s, m_["cT"](x3, np.zeros(0), w0)))
    prim = list(zs[:np_])
    rows, cols, nz = S.sparsehessian(L, prim)
    key = nlp.hessian_lagrangian_sparsity
    # sorted-key order == CSC order of the symmetric pattern (test/hessian_lagrangian.jl:201)
    assert sorted(zip(rows, cols)) == key
    Lxx = S.build_function(nz, zs)
    z0 = np.random.default_rng(2).uniform(size=np_ + nd)
    ref_csc = np.zeros(len(nz))
    Lxx(ref_csc, z0)
    ref = {rc: v for rc, v in zip(zip(rows, cols), ref_csc)}
    # index maps (test/hessian_lagrangian.jl:191-193)
    sd, ad = nlp.trajopt.state_dimensions, nlp.trajopt.action_dimensions
    t = nlp.trajopt
    for idx, sp_ in ((nlp.indices.objective_hessians, E.sparsity_hessian_objective(t.objective, sd, ad)),
                     (nlp.indices.dynamics_hessians, E.sparsity_hessian_dynamics(t.dynamics, sd, ad)),
                     (nlp.indices.stage_hessians, E.sparsity_hessian_stage(t.constraints, sd, ad))):
        flat = [i for v in idx for i in v]
        assert [key[i - 1] for i in flat] == sp_
    h0 = np.zeros(len(key))
    nlp.eval_hessian_lagrangian(h0, z0[:np_], 1.0, z0[np_:])
    got = {rc: v for rc, v in zip(key, h0)}
    err = max(abs(got[rc] - ref[rc]) for rc in key)
    assert err < 1e-8
    # both triangles present and symmetric
    for (r, c) in key:
        assert (c, r) in got and got[(c, r)] == got[(r, c)]
    # sigma scales only the objective part
    h1 = np.zeros(len(key))
    nlp.eval_hessian_lagrangian(h1, z0[:np_], 0.0, z0[np_:])
    h2 = np.zeros(len(key))
    nlp.eval_hessian_lagrangian(h2, z0[:np_], 2.0, z0[np_:])
    assert np.allclose(h2 - h1, 2.0 * (h0 - h1), rtol=1e-12, atol=1e-14)


SHAPES = [("pendulum", dict(T=4)), ("cartpole", dict(T=4)), ("cartpole", dict(T=3, parameterized=False)), ("acrobot", dict(T=3)),
          ("car", dict(T=4, obstacle="general")), ("car", dict(T=4, obstacle="stage"))]


@pytest.mark.parametrize("name,kw", SHAPES, ids=[f"{n}-{'-'.join(str(v) for v in k.values())}" for n, k in SHAPES])
def test_hessian_of_lagrangian_all_baseline_shapes(name, kw):
    """The check of test/hessian_lagrangian.jl:131-205 -- assembled Hessian of the Lagrangian against an
    INDEPENDENT dense symbolic Hessian of the hand-built Lagrangian, both triangles, sorted-key order -- for all
    four BASELINE shapes (small T), including the car with its obstacle as a nonlinear GeneralConstraint (Q7: the
    intended (z, w, lambda) semantics). The reference Hessian is plain `sympy.hessian` of
        L = sigma * sum_t cost_t + sum_t lambda_dyn_t . d(x_{t+1}, x_t, u_t, w_t) + sum_t lambda_stage_t . c_t + lambda_gen . g(z, w)
    built from the model FUNCTIONS, not from the element objects: it shares neither the structural-pattern rules
    (oracle/symbolics.py) nor the index builders (oracle/elements.py) with what it checks."""
    m_ = M.BUILDERS[name](O, **kw)
    solver = O.solver_from(m_)
    nlp = solver.nlp
    T, n, m = m_["T"], m_["n"], m_["m"]
    fns = m_["fns"]
    nw = 8 if m_.get("shared_parameters") else 0
    N_z, N_c = nlp.num_variables, nlp.num_constraint
    zs = [sp.Symbol(f"q{i}") for i in range(N_z)]
    ls = [sp.Symbol(f"l{i}") for i in range(N_c)]
    ws = np.array([sp.Symbol(f"p{i}") for i in range(nw)], dtype=object)
    sigma = sp.Symbol("sg")
    z = np.array(zs, dtype=object)
    X = [z[t * (n + m):t * (n + m) + n] for t in range(T)]
    U = [z[t * (n + m) + n:(t + 1) * (n + m)] for t in range(T - 1)] + [np.zeros(0, dtype=object)]
    L = 0
    for t in range(T):
        L = L + sigma * fns["cost"][t](X[t], U[t], ws)
    row = 0
    for t in range(T - 1):                       # dynamics rows first (src/data.jl:64-75)
        d = M.cat(fns["dyn"](X[t + 1], X[t], U[t], ws))
        L = L + M.dot(ls[row:row + len(d)], d)
        row += len(d)
    for t in range(T):                           # then stage rows
        if fns["con"][t] is not None:
            c = M.cat(fns["con"][t](X[t], U[t], ws))
            L = L + M.dot(ls[row:row + len(c)], c)
            row += len(c)
    if fns["general"] is not None:               # then the general block
        gv = M.cat(fns["general"](z, ws))
        L = L + M.dot(ls[row:row + len(gv)], gv)
        row += len(gv)
    assert row == N_c
    # reference Hessian: central second differences of L itself in 60-digit arithmetic (truncation ~1e-30): no
    # symbolic differentiation at all, so it is independent of sympy's diff as well as of the oracle's rules
    import mpmath
    mpmath.mp.dps = 60
    rng = np.random.default_rng(12)
    z0, l0, w0 = rng.uniform(-1, 1, N_z), rng.normal(size=N_c), rng.normal(size=nw)
    sg = 0.37
    Lf = sp.lambdify(zs, L.subs({**dict(zip(ls, [sp.Float(v, 60) for v in l0])), **dict(zip(ws, [sp.Float(v, 60) for v in w0])),
                                 sigma: sp.Float(sg, 60)}), modules="mpmath")
    h = mpmath.mpf(10) ** -15
    x0 = [mpmath.mpf(float(v)) for v in z0]

    def at(shifts):
        x = list(x0)
        for i, sgn in shifts:
            x[i] = x[i] + sgn * h
        return Lf(*x)

    f0 = at(())
    fp = [at(((i, 1),)) for i in range(N_z)]
    fm = [at(((i, -1),)) for i in range(N_z)]
    dense = np.zeros((N_z, N_z))
    band = 2 * (n + m)                            # knots couple only with their neighbour (and the general block per knot)
    for i in range(N_z):
        dense[i, i] = float((fp[i] - 2 * f0 + fm[i]) / h ** 2)
        for j in range(i + 1, min(N_z, i + band)):
            v = (at(((i, 1), (j, 1))) - at(((i, 1), (j, -1))) - at(((i, -1), (j, 1))) + at(((i, -1), (j, -1)))) / (4 * h ** 2)
            dense[i, j] = dense[j, i] = float(v)
    if nw:
        solver.set_parameters([w0.copy() for _ in range(T)] + [np.zeros(0)])
    key = nlp.hessian_lagrangian_structure()
    hv = np.zeros(len(key))
    nlp.eval_hessian_lagrangian(hv, z0, sg, l0)
    assert key == sorted(set(key))                                    # sorted unique (row, col) keys
    got = np.zeros((N_z, N_z))
    for (r, c), v in zip(key, hv):
        got[r - 1, c - 1] = v
        assert (c, r) in set(key)                                     # both triangles handed to Ipopt (Q8)
    assert np.allclose(got, got.T, rtol=0, atol=0)
    # every non-zero of the exact Hessian has a slot, and the values agree
    missing = [(i + 1, j + 1) for i in range(N_z) for j in range(N_z) if abs(dense[i, j]) > 1e-20 and (i + 1, j + 1) not in set(key)]
    assert missing == []
    assert np.max(np.abs(got - dense)) <= 1e-12 * max(1.0, np.max(np.abs(dense)))


def test_structure_sizes_match_survey_tables():
    """SURVEY App. B global sizes (computed there with an independent throw-away restatement)."""
    expect = {
        ("pendulum", ()): (11, 32, 24, 94, 72, 52),
        ("cartpole", (("T", 51), ("parameterized", False))): (51, 254, 208, 958, 704, 554),
        ("cartpole", (("T", 101),)): (101, 504, 408, 1908, 1404, 1104),
        ("acrobot", (("T", 101),)): (101, 504, 408, 2608, 5502, 4112),
        ("acrobot", (("T", 101), ("stage_endpoint_constraints", False))): (101, 504, 400, 2600, 5502, 4112),
        ("car", (("T", 51), ("obstacle", "stage"))): (51, 253, 201, 752, 602, 553),
    }
    for (name, kw), (T, nz, nc, nj, nhl, nh) in expect.items():
        model = M.BUILDERS[name](O, **dict(kw))
        nlp = O.solver_from(model).nlp
        assert (model["T"], nlp.num_variables, nlp.num_constraint, nlp.num_jacobian, nlp.num_hessian_lagrangian,
                len(nlp.hessian_lagrangian_sparsity)) == (T, nz, nc, nj, nhl, nh), name


def test_local_patterns_match_survey_tables():
    """SURVEY App. B local element patterns, 1-based, CSC order."""
    pend = O.Dynamics(M.pendulum_midpoint, 2, 2, 1, evaluate_hessian=True)
    assert list(zip(*pend.jacobian_sparsity)) == [(1, 1), (2, 1), (1, 2), (2, 2), (2, 3), (1, 4), (2, 4), (1, 5), (2, 5)]
    assert list(zip(*pend.hessian_sparsity)) == [(1, 1), (4, 1), (1, 4), (4, 4)]
    car = O.Dynamics(M.car_midpoint, 3, 3, 2, evaluate_hessian=True)
    assert list(zip(*car.jacobian_sparsity)) == [(1, 1), (2, 2), (1, 3), (2, 3), (3, 3), (1, 4), (2, 4), (3, 5), (1, 6),
                                                  (2, 7), (1, 8), (2, 8), (3, 8)]
    assert sorted(zip(*car.hessian_sparsity)) == sorted([(3, 3), (4, 3), (8, 3), (3, 4), (8, 4), (3, 8), (4, 8), (8, 8)])
    cp = O.Dynamics(M.cartpole_rk3_implicit, 4, 4, 1, evaluate_hessian=True)
    assert list(zip(*cp.jacobian_sparsity)) == [(1, 1), (1, 2), (2, 2), (3, 2), (4, 2), (1, 3), (3, 3), (1, 4), (2, 4),
                                                 (3, 4), (4, 4), (1, 5), (2, 5), (3, 5), (4, 5), (1, 6), (2, 7), (3, 8),
                                                 (4, 9)]
    assert sorted(set(r for r in cp.hessian_sparsity[0])) == [2, 4, 5] and cp.num_hessian == 9
    ac = O.Dynamics(M.acrobot_midpoint, 4, 4, 1, evaluate_hessian=True)
    assert ac.num_jacobian == 26 and ac.num_hessian == 52


def test_bounds_and_constraint_bounds():
    """src/data.jl:123-148: primal bounds placement; inequality rows get lower = -Inf (c <= 0)."""
    model = M.build_car(O, T=6, obstacle="stage")
    nlp = O.solver_from(model).nlp
    lo, hi = nlp.constraint_bounds
    nd = 5 * 3
    assert np.all(lo[:nd] == 0) and np.all(hi == 0)
    assert np.all(np.isneginf(lo[nd:])) and len(lo) == nd + 6
    plo, phi = nlp.variable_bounds
    assert np.array_equal(plo[0:3], model["x1"]) and np.array_equal(phi[0:3], model["x1"])
    assert np.array_equal(plo[3:5], [-0.5, -0.5]) and np.array_equal(phi[3:5], [0.5, 0.5])
    assert np.array_equal(plo[-3:], model["xT"]) and np.isneginf(plo[5])


def test_general_constraint_linear_and_quirks():
    """test/solve.jl:227-296 flavour + Q4 (non-unique count) + Q11 (empty elements)."""
    model = M.build_linear_general(O, T=6)
    nlp = O.solver_from(model).nlp
    assert nlp.general_constraint.num_hessian == 0  # linear -> empty Hessian, guarded
    z = np.random.default_rng(3).uniform(size=nlp.num_variables)
    c = np.zeros(nlp.num_constraint)
    nlp.eval_constraint(c, z)
    assert np.allclose(c[-4:], np.concatenate([z[0:2] - model["x1"], z[-2:] - model["xT"]]))
    H = np.zeros(len(nlp.hessian_lagrangian_sparsity))
    nlp.eval_hessian_lagrangian(H, z, 1.0, np.ones(nlp.num_constraint))
    assert np.allclose(H, 0.2)
    assert nlp.num_hessian_lagrangian >= len(nlp.hessian_lagrangian_sparsity)
