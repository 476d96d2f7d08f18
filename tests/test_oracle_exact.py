"""Pins the oracle's float64 VALUES (which the reference's own tests do not: they use unseeded rand
at 1e-8): every element closure is compared with the exact symbolic derivative evaluated in
50-digit mpmath arithmetic, at the tolerance the GPU path is later held to."""
import mpmath
import numpy as np
import pytest
import sympy as sp

from examples import models as M
from oracle import api as O

RTOL, ATOL = 1e-12, 1e-14


def _mp_eval(exprs, syms, vals):
    mpmath.mp.dps = 50
    f = sp.lambdify(syms, list(exprs), modules="mpmath")
    return [mpmath.mpf(v) for v in f(*[mpmath.mpf(float(v)) for v in vals])]


def _check(name, got, exact):
    for k, (g, e) in enumerate(zip(got, exact)):
        err = abs(mpmath.mpf(float(g)) - e)
        assert err <= ATOL or err <= RTOL * abs(e), f"{name}[{k}]: float64 {g!r} vs exact {mpmath.nstr(e, 20)}"


CASES = [
    ("pendulum", M.pendulum_midpoint, 2, 2, 1),
    ("car", M.car_midpoint, 3, 3, 2),
    ("acrobot", M.acrobot_midpoint, 4, 4, 1),
    ("cartpole", M.cartpole_rk3_implicit, 4, 4, 1),
]


@pytest.mark.parametrize("name,f,ny,nx,nu", CASES, ids=[c[0] for c in CASES])
def test_dynamics_closures_vs_50_digit(name, f, ny, nx, nu):
    d = O.Dynamics(f, ny, nx, nu, evaluate_hessian=True)
    s = d.sym
    r = np.random.default_rng(7)
    for trial in range(2):
        y, x, u = r.uniform(-1.5, 1.5, ny), r.uniform(-1.5, 1.5, nx), r.uniform(-1, 1, nu)
        lam = r.normal(size=ny)
        w = np.zeros(0)
        syms = list(s["y"]) + list(s["x"]) + list(s["u"]) + list(s["lam"])
        vals = list(y) + list(x) + list(u) + list(lam)
        d.evaluate(d.evaluate_cache, y, x, u, w)
        _check(name + ".evaluate", d.evaluate_cache, _mp_eval(s["evaluate"], syms, vals))
        d.jacobian(d.jacobian_cache, y, x, u, w)
        _check(name + ".jacobian", d.jacobian_cache, _mp_eval(s["jacobian"], syms, vals))
        d.hessian(d.hessian_cache, y, x, u, w, lam)
        _check(name + ".hessian", d.hessian_cache, _mp_eval(s["hessian"], syms, vals))


def test_structural_equals_exact_nonzero_pattern():
    """for the four example models the structural (Symbolics-rule) Hessian pattern
    coincides with the exact non-zero pattern of the symbolic Hessian."""
    for name, f, ny, nx, nu in CASES:
        d = O.Dynamics(f, ny, nx, nu, evaluate_hessian=True)
        s = d.sym
        vars_ = list(s["x"]) + list(s["u"]) + list(s["y"])
        lag = sum(l * e for l, e in zip(s["lam"], s["evaluate"]))
        exact = set()
        for j, vj in enumerate(vars_):
            gj = sp.diff(lag, vj)
            for i, vi in enumerate(vars_):
                if sp.diff(gj, vi) != 0:
                    exact.add((i + 1, j + 1))
        assert exact == set(zip(*d.hessian_sparsity)), name
