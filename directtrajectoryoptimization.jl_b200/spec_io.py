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
import re
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
    t = text.replace("^", "**")
    def lit(m):
        tok = m.group(1)
        return f"_F({tok!r})" if any(c in tok for c in ".eE") else f"_I({tok})"

    t = re.sub(r"(?<![\w.])(\d+\.?\d*(?:[eE][+-]?\d+)?)", lit, t)
    env = dict(_FUNCS)
    env.update(symbols)
    env["_F"] = lambda tok: sp.Float(float(tok))   # the double nearest to the literal, 53-bit
    env["_I"] = sp.Integer
    return sp.sympify(eval(t, {"__builtins__": {}}, env))  # noqa: S307 - restricted namespace, numeric grammar only


def _syms(prefix: str, n: int) -> List[sp.Symbol]:
    return [sp.Symbol(f"{prefix}{i + 1}") for i in range(n)]


def _element(role: str, d: dict) -> ElementSpec:
    nx, nu, nw = d["num_state"], d["num_action"], d.get("num_parameter", 0)
    x, u, w = _syms("x", nx), _syms("u", nu), _syms("w", nw)
    table = {s.name: s for s in x + u + w}
    if role == "dyn":
        ny = d["num_next_state"]
        y = _syms("y", ny)
        table.update({s.name: s for s in y})
        ev = [text_to_expr(t, table) for t in d["evaluate"]]
        vars_ = x + u + y
        lam = _syms("lam", ny)
        args = {"y": y, "x": x, "u": u, "w": w, "lam": lam}
        n_out = ny
    elif role == "cost":
        ev = [text_to_expr(d["evaluate"] if isinstance(d["evaluate"], str) else d["evaluate"][0], table)]
        vars_, lam, n_out = x + u, [], 1
        args = {"x": x, "u": u, "w": w}
    else:
        ev = [text_to_expr(t, table) for t in d["evaluate"]]
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
    ev = [text_to_expr(t, table) for t in d["evaluate"]]
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


def dump_spec(spec: ModelSpec, shape: dict = None) -> dict:
    """ModelSpec -> JSON document (what the Julia glue writes)."""
    def el(e: ElementSpec) -> dict:
        d = {"num_state": e.nx, "num_action": e.nu, "num_parameter": e.nw, "evaluate_hessian": e.has_hess,
             "hessian_sparsity": [list(e.hess_rows), list(e.hess_cols)]}
        if e.role == "dyn":
            d["num_next_state"] = e.n_out
        if e.role == "cost":
            d["evaluate"] = expr_to_text(e.evaluate[0])
        else:
            d["evaluate"] = [expr_to_text(x) for x in e.evaluate]
            d["jacobian_sparsity"] = [list(e.jac_rows), list(e.jac_cols)]
            d["indices_inequality"] = list(e.ineq)
        return d

    doc = {"name": spec.name, "dynamics": [el(e) for e in spec.dyn], "costs": [el(e) for e in spec.cost],
           "constraints": [el(e) for e in spec.stage], "general": None, "shape": shape}
    if spec.general is not None:
        g = spec.general
        doc["general"] = {"num_variables": g.num_variables, "num_parameter": g.num_parameter,
                          "evaluate": [expr_to_text(x) for x in g.evaluate],
                          "jacobian_sparsity": [list(g.jac_rows), list(g.jac_cols)], "evaluate_hessian": g.has_hess,
                          "hessian_sparsity": [list(g.hess_rows), list(g.hess_cols)], "indices_inequality": list(g.ineq)}
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
