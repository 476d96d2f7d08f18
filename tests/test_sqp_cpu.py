"""The batched Newton-KKT algorithm (dto_b200/sqp.py) driven by the CPU oracle (tests/sqp_oracle.py): the
host-side logic -- inertia control, Levenberg-Marquardt damping, merit line search with second-order
correction, per-problem masks -- checked without a GPU. The device arm runs the same function
(tests/test_sqp_gpu.py compares the two iterate by iterate)."""
import numpy as np

import dto_b200 as D
from dto_b200 import kkt as PK
from dto_b200 import sqp
from examples import models as M
from oracle import api as O

from sqp_oracle import OracleBackend, band_ldl_solve


def _guess(model, B, seed):
    T, n, m = model["T"], model["n"], model["m"]
    rng = np.random.default_rng(seed)
    z = np.zeros((B, T * n + (T - 1) * m))
    for t in range(T):
        o = t * (n + m)
        z[:, o:o + n] = model["x1"] + (model["xT"] - model["x1"]) * t / (T - 1)
        if t < T - 1:
            z[:, o + n:o + n + m] = rng.normal(size=(B, m))
    return z


def test_oracle_driven_solver_converges_and_meets_reference_acceptance():
    """pendulum swing-up (examples/pendulum/pendulum.jl), 3 problems in lock step: converged KKT point, end points
    within the reference's 1e-3 (test/solve.jl:134-137), problems that converge early stop moving."""
    mo, mp = M.build_pendulum(O), M.build_pendulum(D)
    osolver = O.solver_from(mo)
    perm, bw = PK.analyze(D.solver_from(mp, batch=1).nlp)     # host-only symbolic phase of the product
    B = 3
    be = OracleBackend(osolver, B, perm=perm - 1, bw=bw, linear="band")
    res = sqp.solve(be, _guess(mo, B, 1), options=sqp.SQPOptions(max_iter=40), record=True)
    assert res.converged.all() and res.constraint_violation.max() < 1e-8 and res.dual_residual.max() < 1e-6
    n = mo["n"]
    assert np.all(np.linalg.norm(res.z[:, :n] - mo["x1"], axis=1) < 1e-3)
    assert np.all(np.linalg.norm(res.z[:, -n:] - mo["xT"], axis=1) < 1e-3)
    # a converged problem is frozen: its z does not change in later iterations
    for b in range(B):
        k = int(res.iterations[b])
        for h in res.history[k + 1:]:
            assert np.array_equal(h["z"][b], res.history[k]["z"][b])


def test_inertia_control_regularises_an_indefinite_hessian():
    """A Hessian with negative curvature in the null space of J gives the wrong inertia; the loop must raise the
    per-problem regularisation until D has exactly N_c negative pivots (Ipopt's IC rule)."""
    class Toy:
        xp = sqp._XP(np)
        B, N_z, N_c = 2, 3, 1
        free = np.ones(3)

        def __init__(self):
            self.deltas = []

        def callbacks(self, z, lam, lam_h, delta):
            self.z = z
            f = 0.5 * (-z[:, 0] ** 2 + z[:, 1] ** 2 + z[:, 2] ** 2) + 0.25 * z[:, 0] ** 4
            g = np.stack([-z[:, 0] + z[:, 0] ** 3, z[:, 1], z[:, 2]], 1)
            c = (z[:, 1] + z[:, 2] - 1.0)[:, None]
            self.g, self.c, self.lam = g, c, lam
            return f, g, c

        def _K(self, b, d):
            H = np.diag([-1.0 + 3 * self.z[b, 0] ** 2, 1.0, 1.0]) + d * np.eye(3)
            J = np.array([[0.0, 1.0, 1.0]])
            return np.block([[H, J.T], [J, -1e-9 * np.eye(1)]]), J

        def newton(self, delta, mask=None):
            self.deltas.append(delta.copy())
            sol, nneg, rz = np.zeros((2, 4)), np.zeros(2), np.zeros((2, 3))
            for b in range(2):
                K, J = self._K(b, delta[b])
                rz[b] = self.g[b] + J.T @ self.lam[b]
                sol[b], nneg[b] = band_ldl_solve(K, np.concatenate([rz[b], self.c[b]]), np.arange(4), 3)
            return sol, nneg, rz

        def newton_soc(self, c_soc, delta, mask=None):
            saved, self.c = self.c, c_soc
            try:
                return self.newton(delta)[0]
            finally:
                self.c = saved

        def objective_constraint(self, z):
            return 0.5 * (-z[:, 0] ** 2 + z[:, 1] ** 2 + z[:, 2] ** 2) + 0.25 * z[:, 0] ** 4, (z[:, 1] + z[:, 2] - 1.0)[:, None]

    be = Toy()
    res = sqp.solve(be, np.array([[0.1, 0.0, 0.0], [1.5, 0.3, 0.3]]), options=sqp.SQPOptions(max_iter=60, lm_first=0.0))
    assert res.converged.all()
    assert np.allclose(np.abs(res.z[:, 0]), 1.0, atol=1e-6) and np.allclose(res.z[:, 1:], 0.5, atol=1e-6)
    # problem 0 started where d2f/dx0^2 < 0: its regularisation was raised above 1 - 3 x0^2 at least once, problem 1's was not
    assert max(d[0] for d in be.deltas) > 0.9 and max(d[1] for d in be.deltas[:3]) == 0.0


def test_repacking_the_batch_changes_no_result():
    """Converged problems leave the working batch once they are half of it; every remaining problem must walk through
    exactly the same iterates as without re-packing (bit for bit: problems are independent, updates are row-wise)."""
    mo, mp = M.build_pendulum(O), M.build_pendulum(D)
    osolver = O.solver_from(mo)
    perm, bw = PK.analyze(D.solver_from(mp, batch=1).nlp)
    B = 6
    first = sqp.solve(OracleBackend(osolver, 2, perm=perm - 1, bw=bw, linear="band"), _guess(mo, 2, 1), options=sqp.SQPOptions(max_iter=40))
    z0 = _guess(mo, B, 3)
    z0[[0, 2, 5]] = first.z[0]                       # three problems start at a solution ...
    lam0 = np.zeros((B, osolver.nlp.num_constraint))
    lam0[[0, 2, 5]] = first.lam[0]                   # ... with its multipliers: they converge in the first iteration
    res = {}
    for repack in (False, True):
        be = OracleBackend(osolver, B, perm=perm - 1, bw=bw, linear="band")
        res[repack] = sqp.solve(be, z0, lam0, options=sqp.SQPOptions(max_iter=40, repack=repack, repack_min=1))
    assert res[True].backend.B < B and res[False].backend.B == B       # the batch really shrank
    assert res[True].converged.all() and np.array_equal(res[True].iterations, res[False].iterations)
    assert res[True].iterations[0] == 0 and res[True].iterations[1] > 0
    for k in ("z", "lam", "objective", "constraint_violation"):
        assert np.array_equal(getattr(res[True], k), getattr(res[False], k)), k


def test_interior_point_mode_handles_active_bounds():
    """Inequality bounds on variables (Bound(action_lower, action_upper), src/bounds.jl; Ipopt's barrier treatment of the
    bounds MOI hands it, src/data.jl:240-248) in `sqp.solve`: pendulum swing-up with |u| <= 15, whose unconstrained
    optimum peaks at |u| = 18.6. The result must be a KKT point of the BOUNDED problem, checked from scratch with the
    oracle: feasibility, strict interior, and stationarity g + J'lam = z_L - z_U with z_L, z_U >= 0 supported only on
    controls at their bound (complementarity) -- the multipliers are recovered from the residual, not taken from the
    solver. Bounds that are never active must give back the unconstrained optimum."""
    mo = M.build_pendulum(O, u_bnd=15.0)
    osolver = O.solver_from(mo)
    B = 2
    opts = sqp.SQPOptions(max_iter=60)
    be = OracleBackend(osolver, B, linear="dense", options=opts)
    assert be.bounds is not None and not be.fixed.any()
    res = sqp.solve(be, _guess(mo, B, 1) * 0.3, options=opts)
    assert res.converged.all()
    T, n, m = mo["T"], mo["n"], mo["m"]
    isu = np.zeros(be.N_z, dtype=bool)
    isu[[t * (n + m) + n for t in range(T - 1)]] = True
    nlp = osolver.nlp
    jr = np.array([r for r, _ in nlp.jacobian_structure()]) - 1
    jc = np.array([c for _, c in nlp.jacobian_structure()]) - 1
    for b in range(B):
        z, lam = res.z[b], res.lam[b]
        u = z[isu]
        assert np.all(np.abs(u) < 15.0) and (np.abs(u) > 15.0 - 1e-5).sum() >= 2          # strictly inside, some AT the bound
        c, g, J = np.zeros(be.N_c), np.zeros(be.N_z), np.zeros(len(jr))
        nlp.eval_constraint(c, z); nlp.eval_objective_gradient(g, z); nlp.eval_constraint_jacobian(J, z)
        assert np.max(np.abs(c)) < 1e-8
        Jd = np.zeros((be.N_c, be.N_z)); Jd[jr, jc] = J
        r = g + Jd.T @ lam                                    # = z_L - z_U at a KKT point
        assert np.max(np.abs(r[~isu])) < 1e-5                 # states carry no bound: plain stationarity
        at_lower, at_upper = u < -15.0 + 1e-4, u > 15.0 - 1e-4
        ru = r[isu]
        assert np.all(ru[at_lower] > -1e-5) and np.all(ru[at_upper] < 1e-5)               # signs of the bound multipliers
        assert np.max(np.abs(ru[~(at_lower | at_upper)])) < 1e-4                          # complementarity: inactive => zero
    loose = M.build_pendulum(O, u_bnd=1.0e3)
    ol = O.solver_from(loose)
    r_loose = sqp.solve(OracleBackend(ol, 1, linear="dense", options=opts), _guess(loose, 1, 1) * 0.3, options=opts)
    r_free = sqp.solve(OracleBackend(O.solver_from(M.build_pendulum(O)), 1, linear="dense"), _guess(loose, 1, 1) * 0.3, options=opts)
    assert r_loose.converged.all() and r_free.converged.all()
    assert abs(r_loose.objective[0] - r_free.objective[0]) < 1e-5 * r_free.objective[0] < res.objective.min() - r_free.objective[0]


def test_reference_car_example_inequality_rows_and_bound_continuation():
    """/root/reference/examples/car/car.jl as published (T = 51, |u| <= 0.5 as Bounds, both end states pinned by Bounds, the
    circular obstacle as one INEQUALITY row per knot -- Constraint(obs, ...; indices_inequality = [1]) -- and the example's
    guess: states interpolated, controls 0.001 randn), solved by the oracle-driven arm of the solver. Inequality rows
    c_i(z) <= 0 get slacks that are eliminated into a -t_i/lam_i diagonal of K's (2,2) block; the direct solve runs into
    the control bounds while the dynamics are still violated (no restoration phase), `solve_bounded` then widens the bounds
    tenfold, solves, and tightens back with a warm start. The result must be a KKT point of the TRUE problem, checked from
    scratch: equalities, obstacle rows <= 0 with the active ones at 0, controls within +-0.5 and some at the bound,
    multipliers of the inequality rows >= 0 and zero on inactive rows, stationarity in the free coordinates."""
    T = 51
    mo = M.build_car(O, T=T, obstacle="stage")
    osolver = O.solver_from(mo)
    n, m = mo["n"], mo["m"]
    rng = np.random.default_rng(3)
    z0 = np.zeros((1, T * n + (T - 1) * m))
    for t in range(T):
        o = t * (n + m)
        z0[:, o:o + n] = mo["x1"] + (mo["xT"] - mo["x1"]) * t / (T - 1)
        if t < T - 1:
            z0[:, o + n:o + n + m] = 0.001 * rng.normal(size=(1, m))
    opts = sqp.SQPOptions(max_iter=90)
    be = OracleBackend(osolver, 1, linear="dense", options=opts)
    hasI = be.bounds["hasI"] > 0
    assert hasI.sum() == T and be.fixed.sum() == 2 * n and be.bounds["hasL"].sum() == (T - 1) * m
    direct = sqp.solve(be, z0, options=opts)
    assert not direct.converged.any()                                  # the situation the fallback exists for
    res = sqp.solve_bounded(be, z0, options=opts)
    assert res.converged.all() and res.staged.all()
    z, lam = res.z[0], res.lam[0]
    nlp = osolver.nlp
    c, g = np.zeros(be.N_c), np.zeros(be.N_z)
    nlp.eval_constraint(c, z); nlp.eval_objective_gradient(g, z)
    js = nlp.jacobian_structure()
    J = np.zeros(len(js)); nlp.eval_constraint_jacobian(J, z)
    Jd = np.zeros((be.N_c, be.N_z)); Jd[[r - 1 for r, _ in js], [cc - 1 for _, cc in js]] = J
    assert np.max(np.abs(c[~hasI])) < 1e-8 and np.max(c[hasI]) < 1e-8
    active = c[hasI] > -1e-6
    assert 1 <= active.sum() <= 6                                      # the path touches the obstacle at a few knots
    lamI = lam[hasI]
    assert np.all(lamI > 0.0) and np.max(lamI[~active]) < 1e-5 and np.min(lamI[active]) > 1e-3
    isu = np.zeros(be.N_z, dtype=bool)
    for t in range(T - 1):
        isu[t * (n + m) + n: t * (n + m) + n + m] = True
    u = z[isu]
    assert np.all(np.abs(u) < 0.5) and (np.abs(u) > 0.5 - 1e-5).sum() >= 3
    assert np.allclose(z[:n], mo["x1"]) and np.allclose(z[-n:], mo["xT"])   # pinned exactly
    r = g + Jd.T @ lam                                                 # = z_L - z_U on bounded variables, anything on pinned ones
    free_state = ~isu & ~be.fixed
    assert np.max(np.abs(r[free_state])) < 1e-5
    ru = r[isu]
    inner = np.abs(u) < 0.5 - 1e-4
    assert np.max(np.abs(ru[inner])) < 1e-4 and np.all(ru[u > 0.5 - 1e-5] < 1e-5) and np.all(ru[u < -0.5 + 1e-5] > -1e-5)


def test_bound_arrays_classify_and_scale_bounds():
    """primal_bounds (src/data.jl:123-133) -> solver masks: lower == upper pins a variable, any other finite bound is an
    inequality; a continuation stage widens TWO-SIDED bounds about their midpoint and leaves one-sided and pinned ones
    alone; constraint rows with bounds (-Inf, 0] are inequality rows, anything else but equalities is refused."""
    import pytest
    inf = np.inf
    lo = np.array([-inf, 0.0, 1.0, -2.0, -inf, 3.0])
    up = np.array([inf, 0.0, 5.0, inf, 4.0, 3.0])
    fixed, b = sqp.bound_arrays(lo, up, clo=np.array([0.0, -inf, 0.0]), cup=np.array([0.0, 0.0, 0.0]))
    assert fixed.tolist() == [False, True, False, False, False, True]
    assert b["hasL"].tolist() == [0, 0, 1, 1, 0, 0] and b["hasU"].tolist() == [0, 0, 1, 0, 1, 0] and b["hasI"].tolist() == [0, 1, 0]
    assert b["lo"][2] == 1.0 and b["up"][2] == 5.0 and 0.0 < b["pushL"][2] <= 0.01 * 4.0
    _, b10 = sqp.bound_arrays(lo, up, scale=10.0)
    assert b10["lo"][2] == 3.0 - 20.0 and b10["up"][2] == 3.0 + 20.0            # two-sided: widened about the midpoint 3
    assert b10["lo"][3] == -2.0 and b10["up"][4] == 4.0                          # one-sided: unchanged
    assert sqp.bound_arrays(np.full(3, -inf), np.full(3, inf))[1] is None       # nothing to do: equality-only mode
    with pytest.raises(NotImplementedError):
        sqp.bound_arrays(lo, up, clo=np.array([-1.0]), cup=np.array([1.0]))     # a ranged row is outside the scope
    with pytest.raises(ValueError):
        sqp.bound_arrays(np.array([2.0]), np.array([1.0]))
