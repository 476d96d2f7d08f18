"""Oracle regression against the committed golden fixtures (tests/golden/*.npz), for both the
python restatement and its C twin (no-CSE "reference-like" and CSE'd variants)."""
import os

import numpy as np
import pytest

from examples import models as M
from golden.make_golden import GOLDEN, tag
from oracle import api as O
from oracle import cgen
from util import assert_close, oracle_eval_all

HERE = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,kw,B,config", GOLDEN, ids=[tag(g[0], g[1]) for g in GOLDEN])
def test_oracle_reproduces_golden(name, kw, B, config):
    fx = np.load(os.path.join(HERE, tag(name, kw) + ".npz"))
    model = M.BUILDERS[name](O, **kw)
    solver = O.solver_from(model)
    nlp = solver.nlp
    assert np.array_equal(np.array(nlp.jacobian_structure()).reshape(-1, 2), fx["jac_structure"])
    assert np.array_equal(np.array(nlp.hessian_lagrangian_structure()).reshape(-1, 2), fx["hess_structure"])
    hess = bool(model["evaluate_hessian"])
    out = oracle_eval_all(solver, model, fx["z"], fx["lam"], fx["sigma"], fx["w"], hessian=hess)
    for k in ("f", "g", "c", "J", "H"):
        assert np.array_equal(out[k], fx[k]), k  # same code, same inputs: bit-identical
    if any(d.sym is None for d in model["dynamics"]):
        return  # user-closure Dynamics (src/dynamics.jl:59-101) has no expressions for the C twin to print
    shared = bool(model.get("shared_parameters"))
    if shared:
        solver.set_parameters([np.zeros(8) for _ in range(model["T"])] + [np.zeros(0)])
    for cse in (False, True):
        co = cgen.build_c_oracle(solver, tag(name, kw), shared_parameters=shared, cse=cse)
        got = co.eval(31, fx["z"], fx["lam"], fx["sigma"], fx["w"])
        for k in ("f", "g", "c", "J", "H"):
            assert_close(f"C twin (cse={cse}) {k}", got[k], fx[k], rtol=1e-12, atol=1e-14)
