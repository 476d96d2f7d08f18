"""Symbolic front-end of the B200 code generator.

Plays the role Symbolics.jl plays for the reference's element constructors
(/root/reference/src/dynamics.jl:23-35, src/costs.jl:18-27, src/constraints.jl:27-40,
src/general_constraint.jl:23-36): trace a user function on symbols, find the STRUCTURAL
sparsity of its Jacobian / Hessian (these patterns define the fixed value slots that
`jacobian_structure` / `hessian_lagrangian_structure` hand to Ipopt, so they must be
bit-exact), and produce the derivative expressions the code generator lowers to CUDA.

Patterns are structural, not "derivative != 0":
  * Jacobian (i, j) iff variable j occurs in expression i;
  * Hessian by degree propagation: every sub-expression is summarised as a set of
    monomial "shapes" {var: degree<=2}; sums union the shapes, products merge them,
    nonlinear functions square them; (k,k) is present when a degree reaches 2, (k,l) when
    k and l share a shape; the result is symmetrised.
Storage order is column-major (CSC), 1-based, like SparseMatrixCSC.nzval.
"""
from __future__ import annotations

import sys
from typing import Dict, List, Sequence, Tuple

import numpy as np
import sympy as sp

sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))

Shape = Tuple[Tuple[int, int], ...]  # sorted ((var, degree), ...)


def variables(name: str, n: int) -> np.ndarray:
    out = np.empty(n, dtype=object)
    for i in range(n):
        out[i] = sp.Symbol(f"{name}{i + 1}")
    return out


def flatten(value) -> List[sp.Expr]:
    if isinstance(value, np.ndarray):
        return [sp.sympify(v) for v in value.reshape(-1)]
    if isinstance(value, (list, tuple)):
        out: List[sp.Expr] = []
        for v in value:
            out.extend(flatten(v))
        return out
    return [sp.sympify(value)]


def dot(a, b):
    s = sp.Integer(0)
    for p, q in zip(a, b):
        s = s + p * q
    return s


# ----------------------------------------------------------------------------- Jacobian
def jacobian_pattern(exprs: Sequence[sp.Expr], vars_: Sequence[sp.Symbol]) -> Tuple[List[int], List[int]]:
    col_of = {v: j + 1 for j, v in enumerate(vars_)}
    per_col: Dict[int, List[int]] = {}
    for i, e in enumerate(exprs):
        for s in e.free_symbols:
            j = col_of.get(s)
            if j is not None:
                per_col.setdefault(j, []).append(i + 1)
    rows: List[int] = []
    cols: List[int] = []
    for j in sorted(per_col):
        for i in sorted(per_col[j]):
            rows.append(i)
            cols.append(j)
    return rows, cols


# ----------------------------------------------------------------------------- Hessian
class _Shapes:
    """A set of monomial shapes; `scalar` ({()}) carries no variable."""

    __slots__ = ("s",)

    def __init__(self, shapes):
        self.s = frozenset(shapes)

    @property
    def is_scalar(self) -> bool:  # also true for the empty set, as in the rule being mirrored
        return all(len(t) == 0 for t in self.s)

    @property
    def is_empty(self) -> bool:
        return len(self.s) == 0


_SCALAR = _Shapes([()])


def _plus(a: _Shapes, b: _Shapes) -> _Shapes:
    if a.is_scalar and not b.is_empty:
        return b
    if b.is_scalar and not a.is_empty:
        return a
    if a is b:
        return a
    return _Shapes(a.s | b.s)


def _square(a: _Shapes) -> _Shapes:
    if a.is_scalar:
        return a
    vs = sorted({k for t in a.s for (k, _) in t})
    return _Shapes([tuple((k, 2) for k in vs)])


def _times(a: _Shapes, b: _Shapes) -> _Shapes:
    if a.is_scalar:
        return b
    if b.is_scalar:
        return a
    if a is b:
        return _square(a)
    out = set()
    for t1 in a.s:
        d1 = dict(t1)
        for t2 in b.s:
            d = dict(d1)
            for k, v in t2:
                d[k] = min(2, d1.get(k, 0) + v)
            out.add(tuple(sorted(d.items())))
    return _Shapes(out)


def _binary_nonlinear(a: _Shapes, b: _Shapes) -> _Shapes:
    # f(a, b) nonlinear in both arguments and in their interaction
    r = _Shapes([])
    r = _plus(r, _square(a) if not a.is_scalar else _times(a, a))
    r = _plus(r, _square(b) if not b.is_scalar else _times(b, b))
    r = _plus(r, _times(a, b))
    return r


def _shapes(e: sp.Expr, var_index: Dict[sp.Symbol, int], memo: Dict[sp.Expr, object]):
    """None = not a tracked quantity (number / parameter / multiplier)."""
    got = memo.get(e, 0)
    if got != 0:
        return got
    r = None
    if e.is_Symbol:
        if e in var_index:
            r = _Shapes([((var_index[e], 1),)])
    elif e.is_Number or e.is_NumberSymbol:
        r = None
    elif e.is_Add:
        r = _SCALAR
        for arg in e.args:
            c = _shapes(arg, var_index, memo)
            if c is not None:
                r = _plus(r, c)
    elif e.is_Mul:
        r = _SCALAR
        for arg in e.args:
            c = _shapes(arg, var_index, memo)
            if c is not None:
                r = _times(r, c)
    elif e.is_Pow:
        base = _shapes(e.args[0], var_index, memo)
        expo = _shapes(e.args[1], var_index, memo)
        if base is not None and expo is None:
            r = base if e.args[1] == 1 else _times(base, base)
        else:
            r = _binary_nonlinear(base if base is not None else _SCALAR, expo if expo is not None else _SCALAR)
    elif isinstance(e, sp.Piecewise):
        # ifelse(c, a, b): the branches combine like a sum, the condition carries no curvature (the rule
        # Symbolics applies to IfElse.ifelse as far as recorded in SURVEY App. C; no reference model uses it)
        r = _SCALAR
        for val, _cond in e.args:
            c = _shapes(val, var_index, memo)
            if c is not None:
                r = _plus(r, c)
    elif e.is_Function:
        cs = [_shapes(arg, var_index, memo) for arg in e.args]
        if len(cs) == 1:
            r = _SCALAR if cs[0] is None else _times(cs[0], cs[0])
        elif len(cs) == 2:
            r = _binary_nonlinear(cs[0] if cs[0] is not None else _SCALAR, cs[1] if cs[1] is not None else _SCALAR)
        else:
            raise NotImplementedError(f"function of unknown linearity: {e.func}")
    else:
        raise NotImplementedError(f"unsupported expression node {type(e)}")
    memo[e] = r
    return r


def hessian_pattern(expr: sp.Expr, vars_: Sequence[sp.Symbol]) -> Tuple[List[int], List[int]]:
    var_index = {v: j for j, v in enumerate(vars_)}
    sh = _shapes(sp.sympify(expr), var_index, {})
    per_col: Dict[int, set] = {}
    if sh is not None:
        for t in sh.s:
            for a in range(len(t)):
                k, v = t[a]
                if v >= 2:
                    per_col.setdefault(k, set()).add(k)
                for b in range(a + 1, len(t)):
                    l = t[b][0]
                    per_col.setdefault(k, set()).add(l)
                    per_col.setdefault(l, set()).add(k)
    rows: List[int] = []
    cols: List[int] = []
    for j in sorted(per_col):
        for i in sorted(per_col[j]):
            rows.append(i + 1)
            cols.append(j + 1)
    return rows, cols


# ----------------------------------------------------------------------------- derivatives
def jacobian_values(exprs, vars_, rows, cols) -> List[sp.Expr]:
    return [sp.diff(exprs[r - 1], vars_[c - 1]) for r, c in zip(rows, cols)]


def hessian_values(expr, vars_, rows, cols) -> List[sp.Expr]:
    """Full-symmetric value list in pattern order; each unordered pair differentiated once."""
    first: Dict[int, sp.Expr] = {}
    second: Dict[Tuple[int, int], sp.Expr] = {}
    out = []
    for r, c in zip(rows, cols):
        i, j = (r, c) if r >= c else (c, r)
        if (i, j) not in second:
            if j not in first:
                first[j] = sp.diff(expr, vars_[j - 1])
            second[(i, j)] = sp.diff(first[j], vars_[i - 1])
        out.append(second[(i, j)])
    return out
