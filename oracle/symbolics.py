"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the slice of Symbolics.jl 0.1.29-0.1.32 that the reference
calls from its element constructors (/root/reference/src/dynamics.jl:23-35,
src/costs.jl:18-27, src/constraints.jl:27-40, src/general_constraint.jl:23-36):

    @variables, Symbolics.gradient, Symbolics.sparsejacobian (+jacobian_sparsity),
    Symbolics.sparsehessian (+hessian_sparsity), Symbolics.build_function(...)[2]

Symbolics.jl is an un-vendored dependency (Project.toml:14,22; tests pin 0.1.29 in
test/Project.toml:8-9) and cannot be run here (no Julia), so its *published
algorithm* is restated over sympy:

  * jacobian_sparsity: entry (i, j) iff vars[j] occurs in the canonical form of
    exprs[i]; sparsejacobian differentiates exactly those entries and keeps
    explicit symbolic zeros; storage order is CSC (column-major) like
    SparseMatrixCSC.nzval / findnz.
  * hessian_sparsity: linearity propagation over "term combinations"
    (Symbolics src/linearity.jl): `+` unions, `*` merges degrees pairwise
    (capped at 2), a combination times itself collapses to one term with all
    degrees 2, unary nonlinear f(t) -> t*t, x^p -> x if p == 1 else x*x,
    binary functions use a (linear11, linear22, linear12) triple. The pattern is
    (k,k) when degree>=2 plus every pair inside a term, symmetrised.
  * sparsehessian: lower triangle differentiated, mirrored to a full symmetric
    pattern (both triangles stored).
  * build_function(...)[2]: in-place closure f!(out, args...) that evaluates one
    independent expression per output (no CSE in that era).

PARITY STATUS: the reference holds no golden vectors (its tests use unseeded
rand at 1e-8), so values at 1e-12 are *unpinned by the reference*; this oracle is
pinned by the reference's known-answer tests restated in tests/test_oracle_*.py and
by exact-derivative cross checks in 50-digit mpmath.
"""
from __future__ import annotations

import math
import sys
from typing import Callable, Dict, FrozenSet, Iterable, List, Sequence, Tuple

import numpy as np
import sympy as sp

sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))


# --------------------------------------------------------------------------
# @variables
# --------------------------------------------------------------------------
def variables(name: str, n: int) -> np.ndarray:
    """`@variables name[1:n]` -> vector of n scalar symbols (Symbolics 0.1.x
    gives a plain Vector{Num}; /root/reference/src/dynamics.jl:23)."""
    out = np.empty(n, dtype=object)
    for i in range(n):
        out[i] = sp.Symbol(f"{name}{i + 1}")
    return out


def _as_expr(e) -> sp.Expr:
    return sp.sympify(e)


# --------------------------------------------------------------------------
# Jacobian: occurrence-based structural sparsity, CSC order
# --------------------------------------------------------------------------
def jacobian_sparsity(exprs: Sequence, vars_: Sequence) -> List[Tuple[int, int]]:
    """1-based (row, col) pairs in CSC order (col-major, rows ascending)."""
    index = {v: j for j, v in enumerate(vars_)}
    pairs = set()
    for i, e in enumerate(exprs):
        for s in _as_expr(e).free_symbols:
            j = index.get(s)
            if j is not None:
                pairs.add((i + 1, j + 1))
    return sorted(pairs, key=lambda rc: (rc[1], rc[0]))


def sparsejacobian(exprs: Sequence, vars_: Sequence):
    """-> (rows, cols, nzval) with 1-based indices in CSC order; explicit zeros kept."""
    pat = jacobian_sparsity(exprs, vars_)
    rows = [r for r, _ in pat]
    cols = [c for _, c in pat]
    nz = [sp.diff(_as_expr(exprs[r - 1]), vars_[c - 1]) for r, c in pat]
    return rows, cols, nz


def gradient(expr, vars_: Sequence) -> List[sp.Expr]:
    """Dense gradient including structural zeros (src/costs.jl:20)."""
    e = _as_expr(expr)
    return [sp.diff(e, v) for v in vars_]


# --------------------------------------------------------------------------
# Hessian: linearity propagation (TermCombination)
# --------------------------------------------------------------------------
Term = FrozenSet[Tuple[int, int]]  # frozenset of (var index, degree)


class TermCombination:
    """Set of {var -> degree} maps. `one` = {{}} (a pure scalar), `zero` = {}."""

    __slots__ = ("terms",)

    def __init__(self, terms: Iterable[Term]):
        self.terms = frozenset(terms)

    @staticmethod
    def one() -> "TermCombination":
        return TermCombination([frozenset()])

    @staticmethod
    def zero() -> "TermCombination":
        return TermCombination([])

    @staticmethod
    def idx(i: int) -> "TermCombination":
        return TermCombination([frozenset({(i, 1)})])

    def iszero(self) -> bool:
        return len(self.terms) == 0

    def isone(self) -> bool:
        return all(len(t) == 0 for t in self.terms)  # vacuously true for zero, as in Julia

    def __add__(self, other: "TermCombination") -> "TermCombination":
        if self.isone() and not other.iszero():
            return other
        if other.isone() and not self.iszero():
            return self
        if self is other:
            return self
        return TermCombination(self.terms | other.terms)

    def __mul__(self, other: "TermCombination") -> "TermCombination":
        if self.isone():
            return other
        if other.isone():
            return self
        if self is other:  # squaring: every variable of every term at degree 2, one term
            t = frozenset((k, 2) for term in self.terms for (k, _) in term)
            return TermCombination([t])
        out = set()
        for d1 in self.terms:
            m1 = dict(d1)
            for d2 in other.terms:
                d = dict(m1)
                for k, v in d2:
                    d[k] = min(2, m1.get(k, 0) + v)
                out.add(frozenset(d.items()))
        return TermCombination(out)


# linearity triples (linear11, linear22, linear12) of binary functions
_LINEARITY_2 = {
    "default": (False, False, False),  # ^, atan2, hypot, max, min ... fully nonlinear
}


def _combine_terms_2(lin, t1: TermCombination, t2: TermCombination) -> TermCombination:
    l11, l22, l12 = lin
    term = TermCombination.zero()
    if l11:
        if not l12:
            term = term + t1
    else:
        term = term + t1 * t1
    if l22:
        if not l12:
            term = term + t2
    else:
        term = term + t2 * t2
    if l12:
        term = term + (t1 + t2)
    else:
        term = term + t1 * t2
    return term


def _propagate(e: sp.Expr, index: Dict[sp.Symbol, int], memo: dict):
    """Returns a TermCombination, or None for a sub-expression that is not an
    `idx` (number / non-differentiation symbol) -- those are filtered in + and *."""
    key = e
    if key in memo:
        return memo[key]
    if e.is_Symbol:
        r = TermCombination.idx(index[e]) if e in index else None
    elif e.is_Number or e.is_NumberSymbol:
        r = None
    elif e.is_Add:
        acc = TermCombination.one()
        for a in e.args:
            c = _propagate(a, index, memo)
            if c is not None:
                acc = acc + c
        r = acc
    elif e.is_Mul:
        acc = TermCombination.one()
        for a in e.args:
            c = _propagate(a, index, memo)
            if c is not None:
                acc = acc * c
        r = acc
    elif e.is_Pow:
        base = _propagate(e.args[0], index, memo)
        expo = _propagate(e.args[1], index, memo)
        if base is not None and expo is None:
            p = e.args[1]
            r = base if (p.is_Number and p == 1) else base * base
        else:
            a = base if base is not None else TermCombination.one()
            b = expo if expo is not None else TermCombination.one()
            r = _combine_terms_2(_LINEARITY_2["default"], a, b)
    elif isinstance(e, sp.Piecewise):
        # IfElse.ifelse(c, a, b) (imported at /root/reference/src/DirectTrajectoryOptimization.jl:5, used by no
        # reference model): RECOLLECTION of the Symbolics rule `ifelse(c, x, y) => x + y` -- the branches
        # combine linearly, the condition contributes nothing. Unverifiable here.
        acc = TermCombination.one()
        for (val, _cond) in e.args:
            c = _propagate(val, index, memo)
            if c is not None:
                acc = acc + c
        r = acc
    elif isinstance(e, sp.Function) or e.is_Function:
        args = [_propagate(a, index, memo) for a in e.args]
        if len(args) == 1:
            a = args[0]
            # all unary functions the models use (sin, cos, tan, exp, log, ...) are nonlinear
            r = TermCombination.one() if a is None else a * a
        elif len(args) == 2:
            a = args[0] if args[0] is not None else TermCombination.one()
            b = args[1] if args[1] is not None else TermCombination.one()
            r = _combine_terms_2(_LINEARITY_2["default"], a, b)
        else:
            raise NotImplementedError(f"function of unknown linearity: {e.func}")
    else:
        raise NotImplementedError(f"unsupported node {type(e)}")
    memo[key] = r
    return r


def hessian_sparsity(expr, vars_: Sequence) -> List[Tuple[int, int]]:
    """1-based full-symmetric (row, col) pattern in CSC order."""
    index = {v: j for j, v in enumerate(vars_)}
    lp = _propagate(_as_expr(expr), index, {})
    pairs = set()
    if lp is not None:
        for term in lp.terms:
            kv = list(term)
            for a in range(len(kv)):
                k, v = kv[a]
                if v >= 2:
                    pairs.add((k + 1, k + 1))
                for b in range(a + 1, len(kv)):
                    l = kv[b][0]
                    pairs.add((k + 1, l + 1))
                    pairs.add((l + 1, k + 1))
    return sorted(pairs, key=lambda rc: (rc[1], rc[0]))


def sparsehessian(expr, vars_: Sequence):
    """-> (rows, cols, nzval), full symmetric, CSC order. Lower triangle is
    differentiated (d/dvars[i] of d/dvars[j], i >= j) and mirrored."""
    e = _as_expr(expr)
    pat = hessian_sparsity(e, vars_)
    lower: Dict[Tuple[int, int], sp.Expr] = {}
    dcache: Dict[int, sp.Expr] = {}
    for (i, j) in pat:
        if j > i:
            continue
        if j not in dcache:
            dcache[j] = sp.diff(e, vars_[j - 1])
        lower[(i, j)] = sp.diff(dcache[j], vars_[i - 1])
    rows = [r for r, _ in pat]
    cols = [c for _, c in pat]
    nz = [lower[(i, j)] if i >= j else lower[(j, i)] for (i, j) in pat]
    return rows, cols, nz


# --------------------------------------------------------------------------
# build_function: one independent expression per output, no CSE
# --------------------------------------------------------------------------
_FUNCS = {"sin": "_sin", "cos": "_cos", "tan": "_tan", "exp": "_exp", "log": "_log",
          "sqrt": "_sqrt", "atan": "_atan", "asin": "_asin", "acos": "_acos",
          "sinh": "_sinh", "cosh": "_cosh", "tanh": "_tanh", "Abs": "abs", "atan2": "_atan2"}


def _pystr(e: sp.Expr, names: Dict[sp.Symbol, str]) -> str:
    """Full-precision (repr) python source for one expression tree."""
    if e.is_Symbol:
        return names[e]
    if e.is_Integer:
        return repr(float(int(e)))
    if e.is_Rational:
        return f"({float(e.p)!r}/{float(e.q)!r})"
    if e.is_Float:
        return repr(float(e))
    if e.is_NumberSymbol:
        return repr(float(e))
    if e.is_Add:
        return "(" + " + ".join(_pystr(a, names) for a in e.args) + ")"
    if e.is_Mul:
        return "(" + "*".join(_pystr(a, names) for a in e.args) + ")"
    if e.is_Pow:
        b, p = e.args
        if p.is_Integer:
            k = int(p)
            bs = _pystr(b, names)
            if k == -1:
                return f"(1.0/{bs})"
            if k < 0:
                return f"(1.0/({bs}**{-k}))"
            return f"({bs}**{k})"
        return f"({_pystr(b, names)}**{_pystr(p, names)})"
    if isinstance(e, sp.Piecewise):
        out = "0.0"
        for val, cond in reversed(e.args):
            out = _pystr(val, names) if cond is sp.true else f"({_pystr(val, names)} if {_pystr(cond, names)} else {out})"
        return out
    if e.is_Relational:
        return f"({_pystr(e.lhs, names)} {e.rel_op} {_pystr(e.rhs, names)})"
    if e.is_Function:
        fn = _FUNCS.get(e.func.__name__)
        if fn is None:
            raise NotImplementedError(e.func)
        return fn + "(" + ", ".join(_pystr(a, names) for a in e.args) + ")"
    raise NotImplementedError(type(e))


_ENV = {"_sin": math.sin, "_cos": math.cos, "_tan": math.tan, "_exp": math.exp, "_log": math.log,
        "_sqrt": math.sqrt, "_atan": math.atan, "_asin": math.asin, "_acos": math.acos,
        "_sinh": math.sinh, "_cosh": math.cosh, "_tanh": math.tanh, "_atan2": math.atan2}


def build_function(exprs: Sequence, *args: Sequence) -> Callable:
    """Symbolics.build_function(exprs, args...)[2]: f!(out, args...) in place.
    Every output is evaluated from its own expression tree (no CSE)."""
    names: Dict[sp.Symbol, str] = {}
    for a, arr in enumerate(args):
        for i, s in enumerate(arr):
            names[s] = f"a{a}[{i}]"
    lines = [f"def _f(out, {', '.join(f'a{a}' for a in range(len(args)))}):"]
    for k, e in enumerate(exprs):
        lines.append(f"    out[{k}] = {_pystr(_as_expr(e), names)}")
    lines.append("    return None")
    env = dict(_ENV)
    exec(compile("\n".join(lines), "<oracle build_function>", "exec"), env)
    return env["_f"]
