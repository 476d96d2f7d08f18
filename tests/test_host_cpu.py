"""CPU-side tests of the product: the C-ABI library loads and exports every declared symbol, the
shape assembly (structures, sizes, bounds) is bit-exact against the oracle and the golden
fixtures, error paths return codes, and the codegen emits what it claims. No GPU, no compute."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import dto_b200 as D
from dto_b200 import _lib, codegen, sharding
from examples import models as M
from oracle import api as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

SHAPES = [
    ("pendulum", dict()),
    ("cartpole", dict(T=11)),
    ("cartpole", dict(T=51, parameterized=False)),
    ("acrobot", dict(T=9)),
    ("acrobot", dict(T=12, stage_endpoint_constraints=False)),
    ("car", dict(T=12, obstacle="general")),
    ("car", dict(T=7, obstacle="stage")),
    ("acrobot_hessian_test", dict()),
    ("linear_general", dict(T=6)),
    ("heterogeneous", dict()),
]


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(L, n), f"libdto.so does not export {n} declared in include/dto.h"
    assert L.dto_abi_version() == 1
    assert L.dto_status_string(0) == b"ok" and L.dto_status_string(-5) == b"Hessian not available"
    assert L.dto_device_count() >= 0


@pytest.mark.parametrize("name,kw", SHAPES, ids=[f"{n}-{i}" for i, (n, _) in enumerate(SHAPES)])
def test_structures_bit_exact_vs_oracle(name, kw):
    mo = M.BUILDERS[name](O, **kw)
    mp = M.BUILDERS[name](D, **kw)
    on = O.solver_from(mo).nlp
    pn = D.solver_from(mp, batch=3).nlp
    assert pn.num_variables == on.num_variables
    assert pn.num_constraint == on.num_constraint
    assert pn.num_jacobian == on.num_jacobian
    assert pn.num_hessian == len(on.hessian_lagrangian_sparsity)
    assert pn.num_hessian_lagrangian == on.num_hessian_lagrangian  # non-unique count (Q4)
    assert pn.jacobian_structure() == on.jacobian_structure()
    assert pn.hessian_lagrangian_structure() == on.hessian_lagrangian_structure()
    assert pn.features_available() == on.features_available()
    lo, hi = pn.constraint_bounds
    assert np.array_equal(lo, on.constraint_bounds[0]) and np.array_equal(hi, on.constraint_bounds[1])
    plo, phi = pn.variable_bounds
    assert np.array_equal(plo, on.variable_bounds[0]) and np.array_equal(phi, on.variable_bounds[1])
    # element-level local patterns
    for ep, eo in zip(mp["dynamics"][:1] + mp["objective"][-1:], mo["dynamics"][:1] + mo["objective"][-1:]):
        if hasattr(ep, "jacobian_sparsity"):
            assert ep.jacobian_sparsity == eo.jacobian_sparsity
            assert ep.hessian_sparsity == eo.hessian_sparsity
        else:
            assert ep.sparsity == eo.sparsity


def test_structures_match_golden_fixtures():
    from golden.make_golden import GOLDEN as G, tag
    for name, kw, B, config in G:
        fx = np.load(os.path.join(GOLDEN, tag(name, kw) + ".npz"))
        pn = D.solver_from(M.BUILDERS[name](D, **kw), batch=B).nlp
        r, c = pn.jacobian_structure_arrays()
        assert np.array_equal(np.stack([r, c], 1), fx["jac_structure"])
        r, c = pn.hessian_lagrangian_structure_arrays()
        assert np.array_equal(np.stack([r, c], 1).reshape(-1, 2), fx["hess_structure"])
        assert pn.num_hessian_lagrangian == int(fx["num_hessian_nonunique"])


def test_large_horizon_assembly_is_fast():
    """N4: O(nnz log nnz) structure builder; the reference's is quadratic (minutes at T=1001)."""
    import time
    mp = M.build_cartpole(D, T=1001)
    t = time.time()
    pn = D.solver_from(mp, batch=1).nlp
    assert (pn.num_variables, pn.num_constraint, pn.num_jacobian, pn.num_hessian) == (5004, 4008, 19008, 11004)
    r, c = pn.hessian_lagrangian_structure_arrays()
    assert np.all(np.diff(r * 10**6 + c) > 0)  # sorted, unique
    assert time.time() - t < 30


def test_error_paths_return_codes():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.dto_model_load(b"/nonexistent/model.so", C.byref(h)) == -4
    assert b"dlopen" in L.dto_last_error()
    assert L.dto_model_load(None, C.byref(h)) == -1
    solver = D.solver_from(M.build_pendulum(D), batch=2)
    nlp = solver.nlp
    ip = C.POINTER(C.c_int32)
    T = 11
    kd = np.zeros(T - 1, np.int32)
    kc = np.array([0] * (T - 1) + [1], np.int32)
    ks = np.array([0] + [-1] * (T - 2) + [1], np.int32)
    pd = np.zeros(T, np.int32)

    def desc(**over):
        d = dict(T=T, dynamics_kind=kd.ctypes.data_as(ip), cost_kind=kc.ctypes.data_as(ip), stage_kind=ks.ctypes.data_as(ip),
                 use_general=0, parameter_dim=pd.ctypes.data_as(ip), parameter_offset=None, num_parameter=0)
        d.update(over)
        return _lib.ShapeDesc(**d)

    out = C.c_void_p()
    assert L.dto_shape_create(nlp.model.handle, C.byref(desc()), C.byref(out)) == 0
    assert L.dto_num_variables(out) == 32
    L.dto_shape_destroy(out)
    assert L.dto_shape_create(nlp.model.handle, C.byref(desc(T=1)), C.byref(out)) == -1
    bad = kc.copy()
    bad[3] = 7
    assert L.dto_shape_create(nlp.model.handle, C.byref(desc(cost_kind=bad.ctypes.data_as(ip))), C.byref(out)) == -1
    assert b"cost_kind[3]" in L.dto_last_error()
    # terminal cost (num_action=0) placed on an interior knot: gradient slice would not fit (src/costs.jl:61)
    bad = kc.copy()
    bad[2] = 1
    assert L.dto_shape_create(nlp.model.handle, C.byref(desc(cost_kind=bad.ctypes.data_as(ip))), C.byref(out)) == -1
    assert L.dto_shape_create(nlp.model.handle, C.byref(desc(use_general=1)), C.byref(out)) == -4
    if L.dto_device_count() == 0:  # no GPU: batch creation must fail loudly, never fall back
        b = C.c_void_p()
        assert L.dto_batch_create(nlp.shape, 2, None, 0, C.byref(b)) == -2
        assert b"no CPU fallback" in L.dto_last_error()
        with pytest.raises(_lib.DtoError):
            nlp.eval_objective(np.zeros((2, 32)))
        with pytest.raises(RuntimeError):
            solver.solve()


def test_hessian_switch_semantics():
    """Q9: a Cost without evaluate_hessian makes the Hessian callback unavailable."""
    mp = M.build_cartpole(D, T=5, evaluate_hessian=False)
    nlp = D.solver_from(mp, batch=1).nlp
    assert _lib.lib().dto_hessian_available(nlp.shape) == 0 and nlp.num_hessian == 0
    assert nlp.features_available() == ["Grad", "Jac"]
    mp = M.build_cartpole(D, T=5, evaluate_hessian=True)
    nlp = D.solver_from(mp, batch=1).nlp
    assert _lib.lib().dto_hessian_available(nlp.shape) == 1 and nlp.features_available() == ["Grad", "Jac", "Hess"]


def test_codegen_outputs_and_modes(monkeypatch):
    """Both derivative modes generate a compilable sm_100a library; the DAG mode needs far fewer
    operations than tree-CSE of the expanded derivatives (the whole point of ir.py)."""
    mp = M.build_cartpole(D, T=4)
    s1 = D.solver_from(mp, batch=1)
    log = open(s1.model.path.replace(".so", ".log")).read()
    assert "arch=compute_100a,code=sm_100a" in log and "-lineinfo" in log
    stats_dag = eval(log.strip().splitlines()[-1])
    monkeypatch.setenv("DTO_DERIV", "sympy")
    s2 = D.solver_from(mp, batch=1)
    assert s2.model.path != s1.model.path
    stats_sym = eval(open(s2.model.path.replace(".so", ".log")).read().strip().splitlines()[-1])
    assert stats_dag["dyn0_jac_hess"] < 0.75 * stats_sym["dyn0_jac_hess"]
    src = open(s1.model.path.replace(".so", ".cu")).read()
    assert "sincos(" in src and "__constant__ double dto_k[]" in src and "dto_model_entry" in src
    # general-constraint outputs are rolled into index-shifted templates
    car = D.solver_from(M.build_car(D, T=12, obstacle="general"), batch=1)
    st = eval(open(car.model.path.replace(".so", ".log")).read().strip().splitlines()[-1])
    assert st["gen_templates"] == [1, 1, 1]  # T rows, 2T Jacobian and 2T Hessian entries: one template each


def test_user_jacobian_dynamics_is_traced():
    """Second Dynamics constructor (src/dynamics.jl:59-101, Q14): dense column-major pattern, no Hessian."""
    def f(y, x, u, w):
        return y - (x + 0.1 * M.arr(x[1], u[0] - M.sin(x[0])))

    def fj(J, y, x, u, w):
        J[0, 0], J[0, 1], J[0, 3] = -1.0, -0.1, 1.0
        J[1, 0], J[1, 1], J[1, 2], J[1, 4] = 0.1 * M.cos(x[0]), -1.0, -0.1, 1.0

    d = D.Dynamics(f, fj, 2, 2, 1)
    assert d.num_jacobian == 10 and d.num_hessian == 0
    assert d.jacobian_sparsity == [[1, 2] * 5, [1, 1, 2, 2, 3, 3, 4, 4, 5, 5]]
    with pytest.raises(TypeError):
        D.Dynamics(lambda y, x, u, w: np.linalg.solve(np.eye(2), x.astype(float)), 2, 2, 1)


def test_partition_rule():
    assert sharding.partition(4096, 8) == [(512 * i, 512) for i in range(8)]
    assert sharding.partition(10, 4) == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert sharding.partition(2, 4) == [(0, 1), (1, 1), (2, 0), (2, 0)]
    cover = [i for b, n in sharding.partition(1001, 7) for i in range(b, b + n)]
    assert cover == list(range(1001))


def test_json_spec_round_trip_builds_identical_model(tmp_path):
    """The language-neutral JSON spec (what the Julia glue writes) reproduces the same model library
    content hash as the python front end: same expressions, patterns and gather classes."""
    import json
    from dto_b200 import spec_io
    for name, kw in (("pendulum", dict()), ("cartpole", dict(T=11)), ("car", dict(T=12, obstacle="general"))):
        mp = M.BUILDERS[name](D, **kw)
        s = D.solver_from(mp, batch=1)
        T = mp["T"]
        kd, kc, ks = [k.tolist() for k in s.nlp._keep[:3]]
        doc = spec_io.dump_spec(s.model.spec, dict(T=T, dynamics_kind=kd, cost_kind=kc, stage_kind=ks))
        p = tmp_path / f"{name}.json"
        p.write_text(json.dumps(doc))
        spec2 = spec_io.load_spec(json.loads(p.read_text()))
        assert codegen.spec_hash(spec2) == codegen.spec_hash(s.model.spec), name
        assert spec_io.main([str(p)]) == 0
    # a front-end pattern that disagrees with the structural detection is rejected
    doc["dynamics"][0]["jacobian_sparsity"][0][0] = 2
    with pytest.raises(ValueError):
        spec_io.load_spec(doc)


def test_stage_constraint_using_a_missing_action_is_rejected():
    """a Constraint at the terminal knot (no action there) whose function USES u would put Jacobian
    columns beyond that problem's z; the reference raises a BoundsError, the shape assembly must refuse too. Merely
    declaring num_action > 0 without using u (examples/pendulum's terminal constraint) stays legal."""
    n, m, T = 2, 1, 4
    dt = D.Dynamics(M.pendulum_midpoint, n, n, m)
    ct, cT = D.Cost(lambda x, u, w: M.dot(x, x) + M.dot(u, u), n, m), D.Cost(lambda x, u, w: M.dot(x, x), n, 0)
    ok_T = D.Constraint(lambda x, u, w: x - np.array([1.0, 0.0]), n, m)          # declares an action, does not use it
    bad_T = D.Constraint(lambda x, u, w: x - np.array([1.0, 0.0]) * u[0], n, m)  # uses the action the terminal knot lacks
    bounds = [D.Bound(n, m)] * (T - 1) + [D.Bound(n, 0)]
    s = D.Solver([dt] * (T - 1), [ct] * (T - 1) + [cT], [D.Constraint()] * (T - 1) + [ok_T], bounds, batch=1)
    assert s.nlp.num_constraint == (T - 1) * n + n
    with pytest.raises(_lib.DtoError) as e:
        D.Solver([dt] * (T - 1), [ct] * (T - 1) + [cT], [D.Constraint()] * (T - 1) + [bad_T], bounds, batch=1).nlp.num_constraint
    assert "uses variable 3" in str(e.value)


def test_lone_scaling_is_an_error():
    """eval_hessian_lagrangian(H, z, scaling=2.0) without duals must not silently keep an old sigma."""
    s = D.solver_from(M.build_pendulum(D), batch=1)
    H = np.empty((1, s.nlp.num_hessian))
    with pytest.raises(ValueError):
        s.nlp.eval_hessian_lagrangian(H, None, scaling=2.0)
    J = np.empty((1, s.nlp.num_jacobian))
    with pytest.raises(ValueError):
        s.nlp.eval_jacobian_hessian(J, H, None, scaling=2.0)


def test_native_solver_options_agree_across_the_three_languages():
    """dto_sqp_options (include/dto.h), SQPOptions (sqp.py) and SQPOptions (julia/DTOB200.jl) describe the same algorithm:
    the C defaults (dto_sqp_default_options runs without a GPU) equal the python defaults field by field, the ctypes
    mirror has the C struct's field order, and so has the Julia struct."""
    import ctypes as C
    import re
    from dto_b200 import _lib, sqp
    HERE = os.path.dirname(os.path.abspath(__file__))
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "dto.h")).read()
    body = re.search(r"typedef struct dto_sqp_options \{(.*?)\} dto_sqp_options;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for decl in body.split(";"):
        m = re.match(r"\s*(int32_t|double)\s+(.*)", decl.strip(), flags=re.S)
        if m:
            c_fields += [(n.strip(), m.group(1)) for n in m.group(2).split(",")]
    assert [n for n, _ in c_fields] == [n for n, _ in _lib.SqpOptions._fields_]
    assert all((t is C.c_int32) == (ct == "int32_t") for (_, ct), (_, t) in zip(c_fields, _lib.SqpOptions._fields_))
    co = _lib.SqpOptions()
    _lib.lib().dto_sqp_default_options(C.byref(co))
    po = sqp.SQPOptions()
    for name, _ in c_fields:
        assert float(getattr(co, name)) == float(getattr(po, name)), name
    jl = open(os.path.join(os.path.dirname(HERE), "julia", "DTOB200.jl")).read()
    jbody = re.search(r"Base\.@kwdef struct SQPOptions(.*?)\nend", jl, flags=re.S).group(1)
    j_fields = re.findall(r"(\w+)::(Int32|Float64)\s*=\s*([^;\n]+)", jbody)
    assert [n for n, _, _ in j_fields] == [n for n, _ in c_fields]
    assert all((jt == "Int32") == (ct == "int32_t") for (_, jt, _), (_, ct) in zip(j_fields, c_fields))
    for name, _, val in j_fields:
        assert abs(float(eval(val.strip())) - float(getattr(po, name))) <= 1e-15 * abs(float(getattr(po, name))), name
