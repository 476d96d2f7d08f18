"""Generates tests/golden/*.npz: seeded inputs and the CPU oracle's outputs for the five
callbacks + both structures. The reference itself cannot run here (no Julia), and it ships no
golden vectors, so these fixtures freeze the ORACLE (which is pinned by the reference's restated
known-answer tests and by 50-digit exact derivatives, tests/test_oracle_*.py).

    python tests/golden/make_golden.py        # rewrites the fixtures
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from examples import models as M  # noqa: E402
from oracle import api as O  # noqa: E402
from util import make_inputs, oracle_eval_all  # noqa: E402

# name, builder kwargs, B, config id
GOLDEN = [
    ("pendulum", dict(), 4, 1),
    ("cartpole", dict(T=11), 3, 2),
    ("acrobot", dict(T=9), 3, 3),
    ("car", dict(T=12, obstacle="general"), 3, 4),
    ("car", dict(T=7, obstacle="stage"), 2, 4),
    ("acrobot_hessian_test", dict(), 5, 6),
    ("linear_general", dict(T=6), 2, 7),
    ("piecewise", dict(T=9), 6, 8),
    ("user_jacobian", dict(T=5, nonlinear=True), 3, 9),
]


def tag(name, kw):
    return name + "".join(f"_{k}{v}" for k, v in sorted(kw.items()))


def main():
    for name, kw, B, config in GOLDEN:
        model = M.BUILDERS[name](O, **kw)
        solver = O.solver_from(model)
        nlp = solver.nlp
        nw = 8 if model.get("shared_parameters") else 0
        z, lam, sigma, w = make_inputs(name, model, nlp.num_variables, nlp.num_constraint, nw, B, config)
        out = oracle_eval_all(solver, model, z, lam, sigma, w, hessian=bool(model["evaluate_hessian"]))
        js = np.array(nlp.jacobian_structure(), dtype=np.int64).reshape(-1, 2)
        hs = np.array(nlp.hessian_lagrangian_structure(), dtype=np.int64).reshape(-1, 2)
        lo, hi = nlp.constraint_bounds
        path = os.path.join(HERE, tag(name, kw) + ".npz")
        np.savez_compressed(path, z=z, lam=lam, sigma=sigma, w=w, jac_structure=js, hess_structure=hs, c_lower=lo,
                            c_upper=hi, num_hessian_nonunique=nlp.num_hessian_lagrangian, **out)
        print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
