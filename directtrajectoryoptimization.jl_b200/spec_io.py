"""Language-neutral model specification (JSON) <-> ModelSpec.

This is the interchange a non-python front end uses -- in particular the Julia glue
(julia/DTOB200.jl), which prints the Symbolics.jl expressions the reference's constructors already
build (/root/reference/src/dynamics.jl:24-35 etc.) as plain infix text and hands them to the code
generator:

    python -m dto_b200.spec_io spec.json           # -> prints the path of the built model library

Format (all indices 1-based as in Julia; expressions are infix text over the variable names
x1.. u1.. y1.. w1.. lam1.. z1.., operators + - * / ^ (or **), functions sin cos tan exp log sqrt
atan sinh cosh tanh, decimal literals with up to 17 significant digits):

{ "name": "cartpole",
  "dynamics": [ {"num_next_state": 4, "num_state": 4, "num_action": 1, "num_parameter": 0,
                 "evaluate": ["y1 - (x1 + ...)", ...],
                 "jacobian_sparsity": [[rows...], [cols...]],
                 "evaluate_hessian": true, "hessian_sparsity": [[rows...], [cols...]]} ],
  "costs":      [ {"num_state":..,"num_action":..,"num_parameter":..,"evaluate":"...", "evaluate_hessian":..,
                   "hessian_sparsity": [[..],[..]]} ],
  "constraints":[ {... "evaluate": [...], "indices_inequality": [...], ...} ],
  "general":    null | {"num_variables":..,"num_parameter":..,"evaluate":[...], "jacobian_sparsity":..,
                        "evaluate_hessian":.., "hessian_sparsity":.., "indices_inequality": [...]},
  "shape": {"T": 101, "dynamics_kind": [0,...], "cost_kind": [...], "stage_kind": [..-1..]} }

The structural sparsity patterns are taken from the front end (they are the slots Ipopt was told
about); only derivative VALUES are synthesised here. Patterns are cross-checked against this
package's own structural detection and a mismatch is reported.
"""
from __future__ import annotations

import json
import sys
from typing import Dict, List

import sympy as sp

from . import symbolic as S
from .codegen import ElementSpec, GeneralSpec, ModelSpec, build_model

_FUNCS = {n: getattr(sp, n) for n in ("sin", "cos", "tan", "exp", "log", "sqrt", "atan", "sinh", "cosh", "tanh")}


def expr_to_text(e: sp.Expr) -> str:
    """Infix text with full-precision literals (repr of the double)."""
    e = sp.sympify(e)
    if e.is_Symbol:
        return e.name
    if e.is_Integer:
        return str(int(e)) if int(e) >= 0 else f"({int(e)})"
    if e.is_Rational:
        return f"({int(e.p)}/{int(e.q)})"
    if e.is_Number or e.is_NumberSymbol:
        v = float(e)
        return repr(v) if v >= 0 else f"({v!r})"
    if e.is_Add:
        return "(" + " + ".join(expr_to_text(a) for a in e.args) + ")"
    if e.is_Mul:
        return "(" + "*".join(expr_to_text(a) for a in e.args) + ")"
    if e.is_Pow:
        return f"({expr_to_text(e.args[0])})^({expr_to_text(e.args[1])})"
    if e.is_Function:
        return e.func.__name__ + "(" + ", ".join(expr_to_text(a) for a in e.args) + ")"
    raise NotImplementedError(type(e))


def text_to_expr(text: str, symbols: Dict[str, sp.Symbol]) -> sp.Expr:
    """Julia-printed (or C) expression text -> sympy, through the recursive-descent parser of exprparse.py:
    juxtaposed coefficients (0.5x1, 2sin(x2), -0.5(x1 + y1)), `//` rationals, unicode subscripts, inv, abs2,
    ifelse and comparisons are understood; nothing is `eval`-ed."""
    from .exprparse import parse_expr
    return parse_expr(text, symbols)


def expr_to_c(e: sp.Expr, ref: Dict[sp.Symbol, str]) -> str:
    """One right-hand side the way the Symbolics C target prints it: Julia's Expr printer
    (binary operators with spaces, shortest-repr literals, `1//2` rationals) with `^` rewritten to pow(a, b)
    and variables as zero-based references into the C arguments (`x[1]`)."""
    e = sp.sympify(e)
    if e.is_Symbol:
        return ref[e]
    if e.is_Integer:
        return str(int(e))
    if e.is_Rational:
        return f"{int(e.p)}//{int(e.q)}"
    if e.is_Number or e.is_NumberSymbol:
        return repr(float(e))
    if e.is_Add:
        out = ""
        for k, a in enumerate(e.args):
            neg = a.could_extract_minus_sign() and k > 0
            t = expr_to_c(-a if neg else a, ref)
            if (-a if neg else a).is_Add:
                t = f"({t})"
            out = t if k == 0 else f"{out} {'-' if neg else '+'} {t}"
        return out
    if e.is_Mul:
        parts = []
        for a in e.args:
            t = expr_to_c(a, ref)
            if a.is_Add or (a.is_Number and a < 0) or (a.is_Rational and not a.is_Integer):
                t = f"({t})"
            parts.append(t)
        return " * ".join(parts)
    if e.is_Pow:
        return f"pow({expr_to_c(e.args[0], ref)}, {expr_to_c(e.args[1], ref)})"
    if isinstance(e, sp.Piecewise):
        (a, c), (b, _) = e.args[0], e.args[1]
        return f"ifelse({expr_to_c(c, ref)}, {expr_to_c(a, ref)}, {expr_to_c(b, ref)})"
    if e.is_Relational:
        return f"{expr_to_c(e.lhs, ref)} {e.rel_op} {expr_to_c(e.rhs, ref)}"
    if e.is_Function:
        name = {"Abs": "abs"}.get(e.func.__name__, e.func.__name__)
        return name + "(" + ", ".join(expr_to_c(a, ref) for a in e.args) + ")"
    raise NotImplementedError(type(e))


def c_function(fname: str, exprs, args: Dict[str, List[sp.Symbol]], lhsname: str = "out") -> str:
    """`build_function(exprs, args...; target = Symbolics.CTarget(), fname, lhsname, rhsnames)` as text."""
    ref = {s: f"{n}[{i}]" for n, syms in args.items() for i, s in enumerate(syms)}
    sig = ", ".join([f"double* {lhsname}"] + [f"const double* {n}" for n, syms in args.items() if len(syms)])  # no empty arrays in C
    body = "\n".join(f"  {lhsname}[{i}] = {expr_to_c(e, ref)};" for i, e in enumerate(exprs))
    return f"#include <math.h>\nvoid {fname}({sig}) {{\n{body}\n}}\n"


def _syms(prefix: str, n: int) -> List[sp.Symbol]:
    return [sp.Symbol(f"{prefix}{i + 1}") for i in range(n)]


def _evaluate(d: dict, table: Dict[str, sp.Symbol], cargs: Dict[str, List[sp.Symbol]]) -> List[sp.Expr]:
    """The element's expressions: "evaluate_c" = one Symbolics C-target function over the arguments `cargs`
    (what julia/DTOB200.jl writes), or "evaluate" = Julia-printed text per output over x1.. u1.. names."""
    if "evaluate_c" in d:
        from .exprparse import parse_c_function
        return parse_c_function(d["evaluate_c"], cargs)[2]
    ev = d["evaluate"]
    return [text_to_expr(t, table) for t in ([ev] if isinstance(ev, str) else ev)]


def _element(role: str, d: dict) -> ElementSpec:
    nx, nu, nw = d["num_state"], d["num_action"], d.get("num_parameter", 0)
    x, u, w = _syms("x", nx), _syms("u", nu), _syms("w", nw)
    table = {s.name: s for s in x + u + w}
    if role == "dyn":
        ny = d["num_next_state"]
        y = _syms("y", ny)
        table.update({s.name: s for s in y})
        ev = _evaluate(d, table, {"y": y, "x": x, "u": u, "w": w})
        vars_ = x + u + y
        lam = _syms("lam", ny)
        args = {"y": y, "x": x, "u": u, "w": w, "lam": lam}
        n_out = ny
    elif role == "cost":
        ev = _evaluate(d, table, {"x": x, "u": u, "w": w})[:1]
        vars_, lam, n_out = x + u, [], 1
        args = {"x": x, "u": u, "w": w}
    else:
        ev = _evaluate(d, table, {"x": x, "u": u, "w": w})
        vars_ = x + u
        lam = _syms("lam", len(ev))
        args = {"x": x, "u": u, "w": w, "lam": lam}
        n_out = len(ev)
    has_h = bool(d.get("evaluate_hessian", False))
    if role == "cost":
        jr, jc = [1] * len(vars_), list(range(1, len(vars_) + 1))
    else:
        jr, jc = [list(v) for v in d["jacobian_sparsity"]]
        mine = S.jacobian_pattern(ev, vars_)
        if [list(mine[0]), list(mine[1])] != [jr, jc]:
            raise ValueError(f"{role}: front-end Jacobian pattern differs from the structural pattern detected here")
    hr, hc = ([list(v) for v in d["hessian_sparsity"]] if has_h else ([], []))
    el = ElementSpec(role=role, n_out=n_out, nx=nx, nu=nu, nw=nw, args=args, evaluate=ev, jac_rows=jr, jac_cols=jc,
                     has_hess=has_h, hess_rows=hr, hess_cols=hc, ineq=list(d.get("indices_inequality", [])), vars=vars_,
                     lam=lam)
    if has_h:
        mine = S.hessian_pattern(el.lagrangian, vars_)
        if [list(mine[0]), list(mine[1])] != [hr, hc]:
            raise ValueError(f"{role}: front-end Hessian pattern differs from the structural pattern detected here")
    return el


def _general(d: dict) -> GeneralSpec:
    nz, nw = d["num_variables"], d.get("num_parameter", 0)
    z, w = _syms("z", nz), _syms("w", nw)
    table = {s.name: s for s in z + w}
    ev = _evaluate(d, table, {"z": z, "w": w})
    lam = _syms("lam", len(ev))
    jr, jc = [list(v) for v in d["jacobian_sparsity"]]
    jv = S.jacobian_values(ev, z, jr, jc)
    has_h = bool(d.get("evaluate_hessian", False))
    hr, hc, hv = [], [], []
    if has_h:
        hr, hc = [list(v) for v in d["hessian_sparsity"]]
        hv = S.hessian_values(S.dot(lam, ev), z, hr, hc)
    return GeneralSpec(num_variables=nz, num_parameter=nw, args={"z": z, "w": w, "lam": lam}, evaluate=ev, jac_rows=jr,
                       jac_cols=jc, jac=jv, has_hess=has_h, hess_rows=hr, hess_cols=hc, hess=hv,
                       ineq=list(d.get("indices_inequality", [])))


def load_spec(doc: dict) -> ModelSpec:
    spec = ModelSpec(name=doc.get("name", "model"), dyn=[_element("dyn", d) for d in doc["dynamics"]],
                     cost=[_element("cost", d) for d in doc["costs"]],
                     stage=[_element("stage", d) for d in doc.get("constraints", [])],
                     general=_general(doc["general"]) if doc.get("general") else None)
    sh = doc.get("shape")
    if sh:
        from .recipes import classes_of, knot_meta, knot_recipes
        rec = knot_recipes(sh["T"], sh["dynamics_kind"], sh["cost_kind"], sh["stage_kind"], spec.dyn, spec.cost, spec.stage,
                           spec.general)
        if rec is not None:
            classes, metas, _ = classes_of(rec, knot_meta(sh["T"], sh["dynamics_kind"], sh["cost_kind"], sh["stage_kind"],
                                                          spec.dyn, spec.cost, spec.stage))
            if len(classes) <= 16 and max(len(c) for c in classes) <= 128:
                spec.hg_classes = classes
                spec.hg_meta = metas
    return spec


def dump_spec(spec: ModelSpec, shape: dict = None, style: str = "text") -> dict:
    """ModelSpec -> JSON document (what the Julia glue writes). style "text": one infix string per output;
    "ctarget": one Symbolics-C-target function per element ("evaluate_c")."""
    def put(d, name, exprs, cargs, single=False):
        if style == "ctarget":
            d["evaluate_c"] = c_function(name, exprs, cargs)
        else:
            d["evaluate"] = expr_to_text(exprs[0]) if single else [expr_to_text(x) for x in exprs]

    def el(e: ElementSpec) -> dict:
        d = {"num_state": e.nx, "num_action": e.nu, "num_parameter": e.nw, "evaluate_hessian": e.has_hess,
             "hessian_sparsity": [list(e.hess_rows), list(e.hess_cols)]}
        if e.role == "dyn":
            d["num_next_state"] = e.n_out
        cargs = {k: list(v) for k, v in e.args.items() if k != "lam"}
        if e.role == "cost":
            put(d, "cost_evaluate", e.evaluate, cargs, single=True)
        else:
            put(d, f"{e.role}_evaluate", e.evaluate, cargs)
            d["jacobian_sparsity"] = [list(e.jac_rows), list(e.jac_cols)]
            d["indices_inequality"] = list(e.ineq)
        return d

    doc = {"name": spec.name, "dynamics": [el(e) for e in spec.dyn], "costs": [el(e) for e in spec.cost],
           "constraints": [el(e) for e in spec.stage], "general": None, "shape": shape}
    if spec.general is not None:
        g = spec.general
        doc["general"] = {"num_variables": g.num_variables, "num_parameter": g.num_parameter,
                          "jacobian_sparsity": [list(g.jac_rows), list(g.jac_cols)], "evaluate_hessian": g.has_hess,
                          "hessian_sparsity": [list(g.hess_rows), list(g.hess_cols)], "indices_inequality": list(g.ineq)}
        put(doc["general"], "general_evaluate", g.evaluate, {"z": list(g.args["z"]), "w": list(g.args["w"])})
    return doc


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print(__doc__)
        return 2
    with open(argv[0]) as f:
        doc = json.load(f)
    print(build_model(load_spec(doc), verbose=False))
    return 0


if __name__ == "__main__":
    sys.exit(main())
