"""The text front end a non-python caller uses (julia/DTOB200.jl -> JSON spec -> spec_io -> model library):
Symbolics-C-target functions and Julia-printed expressions are parsed by a real parser (exprparse.py, no eval)
into exactly the expressions the python front end traces -- same content hash, same generated library
(/root/reference/src/dynamics.jl:23-35 builds the expressions; SURVEY App. C records the C-target format)."""
import json
import os
import re

import pytest
import sympy as sp

import dto_b200 as D
from dto_b200 import codegen, spec_io
from dto_b200.exprparse import ParseError, parse_c_function, parse_expr
from examples import models as M

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "fixtures_ctarget")
CASES = [("pendulum", dict()), ("cartpole", dict(T=11)), ("acrobot", dict(T=9)), ("car", dict(T=12, obstacle="general")),
         ("piecewise", dict())]


def tag(name, kw):
    return name + "".join(f"_{k}{v}" for k, v in sorted(kw.items()))


@pytest.mark.parametrize("name,kw", CASES, ids=[tag(*c) for c in CASES])
def test_ctarget_fixture_builds_the_same_model(name, kw):
    """parser(fixture) == python front end: equal content hash => the very same generated CUDA library."""
    doc = json.load(open(os.path.join(FIX, tag(name, kw) + ".json")))
    assert all("evaluate_c" in d and "void" in d["evaluate_c"] for d in doc["dynamics"] + doc["costs"] + doc["constraints"])
    spec2 = spec_io.load_spec(doc)
    s = D.solver_from(M.BUILDERS[name](D, **kw), batch=1)
    assert codegen.spec_hash(spec2) == codegen.spec_hash(s.model.spec)
    for a, b in zip(spec2.dyn, s.model.spec.dyn):
        assert [sp.srepr(e) for e in a.evaluate] == [sp.srepr(e) for e in b.evaluate]
        assert (a.jac_rows, a.jac_cols, a.hess_rows, a.hess_cols) == (b.jac_rows, b.jac_cols, b.hess_rows, b.hess_cols)
    assert os.path.exists(codegen.build_model(spec2))  # cached: no second nvcc run


X = {n: sp.Symbol(n) for n in ("x1", "x2", "y1", "u1", "lam1")}
x1, x2, y1, u1 = X["x1"], X["x2"], X["y1"], X["u1"]
JULIA_TEXT = [
    ("0.5x1 + 2x2", 0.5 * x1 + 2 * x2),                       # juxtaposed coefficients
    ("-0.5(x1 + y1)", -0.5 * (x1 + y1)),                      # coefficient times parenthesis
    ("x1^2 + 1//2*x2", x1 ** 2 + sp.Rational(1, 2) * x2),     # rational: r01 read this as floor division = 0
    ("2sin(x₂)", 2 * sp.sin(x2)),                             # function call after a coefficient, unicode subscript
    ("x₁*(x₂^-1)", x1 / x2),
    ("inv(x1) + abs2(x2)", 1 / x1 + x2 ** 2),
    ("2x1^2", 2 * x1 ** 2),                                   # power binds tighter than juxtaposition
    ("1/2x1", 1 / (2 * x1)),                                  # Julia: juxtaposition binds tighter than /
    ("-x1^2", -(x1 ** 2)),
    ("2^-x1", sp.Integer(2) ** (-x1)),
    ("x1^3.0", x1 ** 3),                                      # the reference's models write x.^3.0
    ("2.0e-5x1 - 3e2", sp.Float(2.0e-5) * x1 - sp.Float(300.0)),
    ("pow(x1, 2) * 1", x1 ** 2),                              # C target: pow and the `* 1` hack
    ("ifelse(x1 > 0, x1, -x1)", sp.Piecewise((x1, x1 > 0), (-x1, True))),
    ("(x1 < 0.5) ? sin(x1) : cos(x2)", sp.Piecewise((sp.sin(x1), x1 < 0.5), (sp.cos(x2), True))),
    ("3.141592653589793u1 + π", sp.Float(3.141592653589793) * u1 + sp.Float(3.141592653589793)),
]


@pytest.mark.parametrize("text,expected", JULIA_TEXT, ids=[t for t, _ in JULIA_TEXT])
def test_julia_printed_expressions(text, expected):
    got = spec_io.text_to_expr(text, X)
    assert sp.srepr(got) == sp.srepr(sp.sympify(expected)), (got, expected)


@pytest.mark.parametrize("text", ["().__class__.__base__.__subclasses__()", "__import__('os').system('true')", "x1.real",
                                  "x9 + 1", "2 x1", "x1 +", "x1 ? x2 : y1", "max(x1, x2)", "x1[0]", "lambda: 1", "x1; x2"])
def test_parser_rejects_what_it_does_not_know(text):
    """nothing is eval-ed: attribute access, calls of unknown functions, unknown names, a space where
    Julia needs juxtaposition, a non-boolean condition -- all are errors, never code."""
    with pytest.raises(ParseError):
        parse_expr(text, X)


def test_c_function_parser():
    y, x, u = [sp.Symbol("y1")], [sp.Symbol("x1"), sp.Symbol("x2")], [sp.Symbol("u1")]
    src = """#include <math.h>
    void diffeqf(double* du, const double* RHS1, const double* RHS2, const double* RHS3) {
      du[0] = RHS1[0] - (RHS2[0] + 0.05 * pow(RHS2[1], 2) * 1);
      du[1] = 2RHS3[0] * sin(RHS2[0]) /* inline */ - 1//3 * RHS2[1];
    }"""
    name, out, ev = parse_c_function(src, {"RHS1": y, "RHS2": x, "RHS3": u})
    assert (name, out) == ("diffeqf", "du")
    assert sp.srepr(ev[0]) == sp.srepr(y[0] - (x[0] + sp.Float(0.05) * x[1] ** 2))
    assert sp.srepr(ev[1]) == sp.srepr(2 * u[0] * sp.sin(x[0]) - sp.Rational(1, 3) * x[1])
    for bad in (src.replace("RHS2[1], 2", "RHS2[2], 2"),      # index out of range
                src.replace("du[1]", "du[3]"),                # outputs not 0..n-1
                src.replace("du[1] =", "dv[1] ="),            # assignment to something else
                src.replace("const double* RHS3", "const double* q")):  # argument without a symbol vector
        with pytest.raises(ParseError):
            parse_c_function(bad, {"RHS1": y, "RHS2": x, "RHS3": u})


def test_julia_glue_defines_what_it_uses():
    """julia/DTOB200.jl cannot be executed here; at least every function it CALLS is defined in the file or
    belongs to Julia Base / the packages it imports (r01 shipped a file that advertised export_spec without
    defining it)."""
    src = open(os.path.join(os.path.dirname(HERE), "julia", "DTOB200.jl")).read()
    code = "\n".join(line.split("#")[0] if not line.lstrip().startswith("#") else "" for line in src.splitlines())
    code = re.sub(r'"(?:[^"\\]|\\.)*"', '""', code)  # drop string literals
    defined = set(re.findall(r"function\s+(?:\w+\.)*(\w+!?)", code)) | set(re.findall(r"^\s*(?:\w+\.)*(\w+!?)\([^)\n]*\)\s*=", code, flags=re.M))
    defined |= set(re.findall(r"(?:mutable\s+)?struct\s+(\w+)", code)) | set(re.findall(r"const\s+(\w+)\s*=", code))
    called = set(re.findall(r"(?<![\w.:@])([A-Za-z_]\w*!?)\(", code))
    base = {"get", "normpath", "joinpath", "error", "unsafe_string", "ccall", "Ref", "pointer", "length", "collect", "zip", "ones",
            "zeros", "fill", "fill!", "first", "eachindex", "Symbol", "findnz", "dot", "string", "replace", "join", "tempname", "write",
            "strip", "read", "setenv", "haskey", "push!", "enumerate", "vcat", "repeat", "fieldnames", "typeof", "getfield", "String",
            "Int", "Int32", "Int64", "Cint", "Float64", "Vector", "Dict", "IdDict", "Ptr", "isempty", "in", "if", "for", "while",
            "sizeof", "undef", "Tuple", "Matrix", "Cdouble", "Val", "convert", "throw", "size", "min", "max",
            "f", "empty"}  # (function-valued arguments)
    unknown = sorted(n for n in called - defined - base if not n[0].isupper() or n in ("Solver",))
    assert unknown == [] or unknown == ["Solver"], unknown
    for name in ("export_spec", "dynamics_spec", "cost_spec", "constraint_spec", "general_spec", "Solver", "build_model",
                 "Dynamics", "Cost", "Constraint", "GeneralConstraint", "BatchedNLPData", "BatchedKKT"):
        assert name in defined, name
    # every ccall names an entry point that include/dto.h declares
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "dto.h")).read()
    for sym in set(re.findall(r"ccall\(\(:(\w+)", src)) | set(re.findall(r"structure\(:(\w+)", src)):
        assert re.search(r"\b" + sym + r"\(", hdr), sym
