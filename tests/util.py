"""Shared test helpers: seeded inputs, oracle batch evaluation, comparisons."""
from __future__ import annotations

import math

import numpy as np

from examples import models as M
from oracle import api as O

RTOL, ATOL = 1e-12, 1e-14   # BASELINE.json north_star: 1e-12 relative or 1e-14 absolute


def rng_for(config: int, shard: int = 0):
    return np.random.default_rng(20261017 + 1000 * config + shard)


def make_inputs(name: str, model: dict, nz: int, nc: int, nw: int, B: int, config: int = 0, shard: int = 0):
    """z, lambda, sigma, w per SURVEY 8(d)."""
    r = rng_for(config, shard)
    T, n, m = model["T"], model["n"], model["m"]
    z = np.empty((B, nz))
    if name == "cartpole":
        X = np.concatenate([r.uniform(-math.pi, math.pi, (B, T, 2)), r.normal(0, 1, (B, T, 2))], axis=2)
        U = r.uniform(-3, 3, (B, T, m))
    elif name == "acrobot":
        X = np.concatenate([r.uniform(-math.pi, math.pi, (B, T, 2)), r.normal(0, 1, (B, T, 2))], axis=2)
        U = r.uniform(-3, 3, (B, T, m))
    elif name == "car":
        X = np.concatenate([r.uniform(0, 1, (B, T, 2)), r.uniform(-math.pi, math.pi, (B, T, 1))], axis=2)
        U = r.uniform(-0.5, 0.5, (B, T, m))
    else:
        X = r.uniform(0, 1, (B, T, n))
        U = r.uniform(0, 1, (B, T, m))
    for t in range(T):
        o = t * (n + m)
        z[:, o:o + n] = X[:, t]
        if t < T - 1:
            z[:, o + n:o + n + m] = U[:, t]
    lam = r.normal(0, 1, (B, nc)) if name in ("cartpole", "acrobot", "car") else r.uniform(0, 1, (B, nc))
    sigma = np.ones(B)
    if B > 1:
        sigma[1::2] = 0.37
    w = np.zeros((B, nw))
    if nw == 8:  # cartpole: w = [x1; xT]
        w[:, 0:4] = r.normal(0, 0.1, (B, 4))
        w[:, 4:8] = np.array([0.0, math.pi, 0.0, 0.0]) + r.normal(0, 0.1, (B, 4))
    elif nw > 0:
        w[:] = r.normal(0, 1, (B, nw))
    return z, lam, sigma, w


def oracle_parameters(model: dict, w_row: np.ndarray):
    """Per-knot parameter list for ONE problem from its flat row (reference: T+1 vectors, Q10)."""
    T = model["T"]
    if model.get("shared_parameters"):
        return [w_row.copy() for _ in range(T)] + [np.zeros(0)]
    return None


def oracle_eval_all(osolver, model, z, lam, sigma, w, hessian=True):
    """Run the five callbacks of the CPU oracle problem by problem."""
    nlp = osolver.nlp
    B = z.shape[0]
    out = dict(f=np.zeros(B), g=np.zeros((B, nlp.num_variables)), c=np.zeros((B, nlp.num_constraint)),
               J=np.zeros((B, nlp.num_jacobian)), H=np.zeros((B, len(nlp.hessian_lagrangian_sparsity))))
    for b in range(B):
        p = oracle_parameters(model, w[b])
        if p is not None:
            osolver.set_parameters(p)
        out["f"][b] = nlp.eval_objective(z[b])
        nlp.eval_objective_gradient(out["g"][b], z[b])
        nlp.eval_constraint(out["c"][b], z[b])
        nlp.eval_constraint_jacobian(out["J"][b], z[b])
        if hessian:
            nlp.eval_hessian_lagrangian(out["H"][b], z[b], float(sigma[b]), lam[b])
    return out


def assert_close(name, got, ref, rtol=RTOL, atol=ATOL):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f"{name}: shape {got.shape} vs {ref.shape}"
    err = np.abs(got - ref)
    ok = (err <= atol) | (err <= rtol * np.abs(ref))
    if not ok.all():
        idx = np.argwhere(~ok)
        i = tuple(idx[0])
        worst = np.nanmax(np.where(np.abs(ref) > 0, err / np.maximum(np.abs(ref), 1e-300), 0))
        raise AssertionError(f"{name}: {len(idx)} of {got.size} values differ beyond rtol={rtol}/atol={atol}; "
                             f"first at {i}: got {got[i]!r} ref {ref[i]!r}; worst rel {worst:.3e}")
