"""CPU tests of the KKT consumer: the oracle restatement of
/root/reference/examples/pendulum/pendulum.jl:109-211 is pinned by the script's own check
(`norm(sol - H \\ h, Inf)` small), and the product's host-only symbolic phase (ordering, bandwidth)
is checked against the structures. No GPU, no product numerics."""
import numpy as np
import pytest

import dto_b200 as D
from dto_b200 import kkt as PK
from examples import models as M
from oracle import api as O
from oracle import kkt as OK

SHAPES = [
    ("pendulum", dict()),
    ("cartpole", dict(T=11)),
    ("cartpole", dict(T=101)),
    ("acrobot", dict(T=9)),
    ("acrobot", dict(T=101)),
    ("car", dict(T=12, obstacle="general")),
    ("car", dict(T=201, obstacle="general")),
    ("car", dict(T=7, obstacle="stage")),
    ("acrobot_hessian_test", dict()),
    ("heterogeneous", dict()),
]


def test_oracle_pendulum_script_check():
    """pendulum.jl:120-211 on the example problem itself: z, y ~ rand (seeded here), regularisation
    1e-5 / 1e-5, and the script's final check norm(sol - H \\ h, Inf)."""
    mo = M.BUILDERS["pendulum"](O)
    nlp = O.solver_from(mo).nlp
    r = np.random.default_rng(20261017)
    z, y = r.uniform(0, 1, nlp.num_variables), r.uniform(0, 1, nlp.num_constraint)
    out = OK.kkt_solve(nlp, z, y)
    K, h, sol = out["K"], out["h"], out["sol"]
    n = nlp.num_variables + nlp.num_constraint
    assert K.shape == (n, n) and np.array_equal(K, K.T)
    # blocks as the script builds them
    nz = nlp.num_variables
    assert np.all(np.diag(K)[nz:] == -1.0e-5)
    for k, (i, j) in enumerate(nlp.jacobian_structure()):
        assert K[nz + i - 1, j - 1] == out["J"][k] == K[j - 1, nz + i - 1]
    ref = np.linalg.solve(K, h)
    assert np.max(np.abs(sol - ref)) <= 1e-7 * max(1.0, np.max(np.abs(ref)))
    # L D L' reconstructs K; inertia of a quasi-definite matrix: nz positive, ny negative pivots
    Lm, Dv = out["L"], out["D"]
    assert np.max(np.abs(Lm @ np.diag(Dv) @ Lm.T - K)) <= 1e-12 * np.max(np.abs(K))
    assert int((Dv > 0).sum()) == nz and int((Dv < 0).sum()) == nlp.num_constraint
    # C'y by hand
    Jd = np.zeros((nlp.num_constraint, nz))
    for k, (i, j) in enumerate(nlp.jacobian_structure()):
        Jd[i - 1, j - 1] = out["J"][k]
    assert np.allclose(h[:nz], out["g"] + Jd.T @ y, rtol=1e-14, atol=1e-15)
    assert np.array_equal(h[nz:], out["c"])


def test_oracle_reproduces_kkt_golden():
    """tests/golden/kkt_*.npz (make_golden_kkt.py): same code, same inputs => bit-identical; and the frozen
    solutions satisfy the script's own check against a dense LU solve."""
    import os
    from golden.make_golden import tag
    from golden.make_golden_kkt import KKT_GOLDEN, REG
    from util import oracle_parameters
    here = os.path.join(os.path.dirname(__file__), "golden")
    for name, kw, _ in KKT_GOLDEN:
        fx = np.load(os.path.join(here, tag(name, kw) + ".npz"))
        gk = np.load(os.path.join(here, "kkt_" + tag(name, kw) + ".npz"))
        model = M.BUILDERS[name](O, **kw)
        solver = O.solver_from(model)
        for b in range(fx["z"].shape[0]):
            p = oracle_parameters(model, fx["w"][b])
            if p is not None:
                solver.set_parameters(p)
            r = OK.kkt_solve(solver.nlp, fx["z"][b], fx["lam"][b], REG, REG)
            assert np.array_equal(r["K"], gk["K"][b]) and np.array_equal(r["h"], gk["h"][b])
            assert np.array_equal(r["sol"], gk["sol"][b])
            ref = np.linalg.solve(gk["K"][b], gk["h"][b])
            assert np.max(np.abs(gk["sol"][b] - ref)) <= 1e-6 * max(1.0, np.max(np.abs(ref)))


def test_oracle_qdldl_with_permutation():
    r = np.random.default_rng(7)
    n1, n2 = 9, 5
    A = r.normal(size=(n1, n1))
    Jm = r.normal(size=(n2, n1)) * (r.uniform(size=(n2, n1)) < 0.5)
    K = np.block([[A @ A.T + np.eye(n1), Jm.T], [Jm, -0.1 * np.eye(n2)]])
    b = r.normal(size=n1 + n2)
    perm = r.permutation(n1 + n2)
    for p in (None, perm):
        Lm, Dv, Dinv, pos = OK.qdldl_factor(K, p)
        Kp = K if p is None else K[np.ix_(p, p)]
        assert np.max(np.abs(Lm @ np.diag(Dv) @ Lm.T - Kp)) < 1e-12
        assert pos == n1
        x = OK.qdldl_solve(Lm, Dinv, b, p)
        assert np.max(np.abs(K @ x - b)) < 1e-11


@pytest.mark.parametrize("name,kw", SHAPES, ids=[f"{n}-{i}" for i, (n, _) in enumerate(SHAPES)])
def test_analyze_permutation_and_bandwidth(name, kw):
    pn = D.solver_from(M.BUILDERS[name](D, **kw), batch=1).nlp
    perm, bw = PK.analyze(pn)
    n = pn.num_variables + pn.num_constraint
    assert sorted(perm.tolist()) == list(range(1, n + 1))
    pos = np.empty(n + 1, dtype=np.int64)
    pos[perm] = np.arange(n)
    nz = pn.num_variables
    worst = 0
    for (i, j) in pn.hessian_lagrangian_structure():
        worst = max(worst, abs(int(pos[i]) - int(pos[j])))
    for (i, j) in pn.jacobian_structure():
        worst = max(worst, abs(int(pos[nz + i]) - int(pos[j])))
    assert worst == bw
    assert bw <= 31, f"{name}: half bandwidth {bw} is outside the banded kernels"


def test_kkt_create_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pn = D.solver_from(M.BUILDERS["pendulum"](D), batch=2).nlp
    with pytest.raises(Exception) as e:
        PK.KKTSystem(pn)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_kkt_abi_error_codes_without_gpu():
    import ctypes as C
    from dto_b200 import _lib
    L = _lib.lib()
    assert L.dto_status_string(-7) == b"unsupported shape"
    assert L.dto_kkt_analyze(None, None, None) == -1            # DTO_ERR_BAD_ARG
    assert b"null shape" in L.dto_last_error()
    h = C.c_void_p()
    assert L.dto_kkt_create(None, 1e-5, 1e-5, C.byref(h)) == -1 and not h.value
    assert L.dto_kkt_solve(None, None) == -1
    assert L.dto_kkt_solve_host(None, None, None, None, None, 0) == -1
    assert L.dto_kkt_dim(None) == -1 and L.dto_kkt_bandwidth(None) == -1
    L.dto_kkt_destroy(None)                                     # no-op


def test_kkt_ordering_of_a_cyclic_coupling():
    """A general constraint tying the first knot to the last one (a cycle in the KKT graph) still orders into a
    narrow band: reverse Cuthill-McKee interleaves the two arms of the cycle."""
    pn = D.solver_from(M.BUILDERS["heterogeneous"](D), batch=1).nlp
    perm, bw = PK.analyze(pn)
    assert 15 < bw <= 31      # the 32-lane kernel's range (tests/test_kkt_gpu.py::test_kkt_heterogeneous_shape)
