"""Expression text -> sympy, by a small recursive-descent parser (no `eval`).

Two producers feed it:
  * the Symbolics C target (`build_function(...; target = Symbolics.CTarget())`): a C
    function `void f(double* out, const double* y, const double* x, ...) { out[0] = ...; }` whose
    right-hand sides are Julia-printed expressions with `^` rewritten to `pow` and a `* 1` appended to
    literals -- `parse_c_function` reads the whole function;
  * plain Julia `string(expr)` output (the JSON interchange of spec_io.py) -- `parse_expr`.
Because the C target prints the expressions with Julia's printer, one grammar covers both: the Julia
superset the reference's models and Symbolics can produce (/root/reference/src/dynamics.jl:23-35 builds
the expressions; examples/*.jl and test/*.jl use + - * / ^ sin cos tan dot):

    expr    := cmp [ '?' expr ':' expr ]                      (C ternary)
    cmp     := sum [ ('<' | '<=' | '>' | '>=' | '==' | '!=') sum ]
    sum     := term { ('+' | '-') term }
    term    := rat { ('*' | '/' | '÷'-free) rat }
    rat     := unary { '//' unary }                           (Julia rational: exact, never floor division)
    unary   := ('-' | '+') unary | juxt
    juxt    := NUMBER power-operand                            (Julia coefficient: 2x1, 0.5sin(x2), -0.5(x1 + y1))
             | power
    power   := atom [ ('^' | '**') unary ]                     (right associative)
    atom    := NUMBER | NAME | NAME '[' INT ']' | NAME '(' expr {',' expr} ')' | '(' expr ')'

Names may carry unicode subscripts (x₁ -> x1) and `π`. Functions: sin cos tan exp log sqrt atan sinh cosh
tanh abs/fabs, pow(a, b), inv(a), abs2(a), ifelse(c, a, b), min/max/fmin/fmax are NOT accepted (no rule
for their sparsity). Integer literals stay exact (sympy Integer), decimal literals become the nearest
double (sympy Float, 53 bits), `a//b` a sympy Rational.
"""
from __future__ import annotations

import re
from typing import Callable, Dict, List, Optional, Tuple

import sympy as sp

_SUB = str.maketrans("₀₁₂₃₄₅₆₇₈₉", "0123456789")
_NUM = re.compile(r"(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?")
_NAME = re.compile(r"[A-Za-z_α-ωΑ-Ω][A-Za-z_0-9α-ωΑ-Ω]*")
_OPS = ("<=", ">=", "==", "!=", "//", "**", "+", "-", "*", "/", "^", "(", ")", "[", "]", ",", "<", ">", "?", ":", ";", "=")

_UNARY = {n: getattr(sp, n) for n in ("sin", "cos", "tan", "exp", "log", "sqrt", "atan", "sinh", "cosh", "tanh", "asin", "acos")}
_UNARY["abs"] = sp.Abs
_UNARY["fabs"] = sp.Abs


class ParseError(ValueError):
    pass


def tokenize(text: str) -> List[Tuple[str, str, bool]]:
    """-> [(kind, text, glued)] with kind in num / name / op; `glued` = no whitespace before the token
    (needed for Julia's juxtaposition rule: `2x` is a product, `2 x` is an error)."""
    text = text.translate(_SUB)
    out: List[Tuple[str, str, bool]] = []
    i, n = 0, len(text)
    glued = False
    while i < n:
        c = text[i]
        if c.isspace():
            i += 1
            glued = False
            continue
        m = _NUM.match(text, i)
        if m and (c.isdigit() or c == "."):
            tok = m.group(0)
            # "2e" followed by a name is the coefficient 2 times a variable starting with e (not an exponent)
            out.append(("num", tok, glued))
            i = m.end()
            glued = True
            continue
        m = _NAME.match(text, i)
        if m:
            out.append(("name", m.group(0), glued))
            i = m.end()
            glued = True
            continue
        for op in _OPS:
            if text.startswith(op, i):
                out.append(("op", op, glued))
                i += len(op)
                glued = True
                break
        else:
            raise ParseError(f"unexpected character {c!r} at {i} in {text[max(0, i - 20):i + 20]!r}")
    return out


Resolver = Callable[[str, Optional[int]], sp.Expr]


class _Parser:
    def __init__(self, toks, resolve: Resolver):
        self.t = toks
        self.i = 0
        self.resolve = resolve

    # ---- token helpers
    def peek(self, k=0):
        j = self.i + k
        return self.t[j] if j < len(self.t) else ("end", "", False)

    def take(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, *ops) -> Optional[str]:
        k, s, _ = self.peek()
        if k == "op" and s in ops:
            self.i += 1
            return s
        return None

    def expect(self, op):
        if not self.accept(op):
            raise ParseError(f"expected {op!r}, found {self.peek()[1]!r}")

    # ---- grammar
    def expr(self):
        c = self.cmp()
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            return _select(c, a, b)
        return c

    def cmp(self):
        a = self.sum()
        op = self.accept("<=", ">=", "==", "!=", "<", ">")
        if op:
            b = self.sum()
            return {"<": sp.Lt, "<=": sp.Le, ">": sp.Gt, ">=": sp.Ge, "==": sp.Eq, "!=": sp.Ne}[op](a, b)
        return a

    def sum(self):
        # all terms of a chain go into ONE Add / all factors of a chain into ONE Mul: sympy's two-argument
        # shortcuts (a Number times an Add is distributed) would otherwise rebuild 1.0*(a + b)*c differently
        # from the traced expression, and the content hash of the model would change
        terms = [self.term()]
        while True:
            op = self.accept("+", "-")
            if not op:
                return terms[0] if len(terms) == 1 else sp.Add(*terms)
            b = self.term()
            terms.append(b if op == "+" else -b)

    def term(self):
        factors = [self.rat()]
        while True:
            op = self.accept("*", "/")
            if not op:
                return factors[0] if len(factors) == 1 else sp.Mul(*factors)
            b = self.rat()
            factors.append(b if op == "*" else sp.Pow(b, -1))

    def rat(self):
        a = self.unary()
        while self.accept("//"):
            b = self.unary()
            if a.is_Integer and b.is_Integer:
                a = sp.Rational(int(a), int(b))
            else:
                a = a / b
        return a

    def unary(self):
        op = self.accept("-", "+")
        if op:
            v = self.unary()
            return -v if op == "-" else v
        return self.juxt()

    def juxt(self):
        k, s, _ = self.peek()
        if k == "num":
            k2, s2, glued2 = self.peek(1)
            if glued2 and (k2 == "name" or (k2 == "op" and s2 == "(")):
                self.take()
                return _number(s) * self.power()   # 2x^2 = 2*(x^2): the power binds tighter
        return self.power()

    def power(self):
        a = self.atom()
        if self.accept("^", "**"):
            b = self.unary()   # right associative, and a^-b is allowed
            return _pow(a, b)
        return a

    def atom(self):
        k, s, _ = self.take()
        if k == "num":
            return _number(s)
        if k == "op" and s == "(":
            v = self.expr()
            self.expect(")")
            return v
        if k == "name":
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    args.append(self.expr())
                    while self.accept(","):
                        args.append(self.expr())
                    self.expect(")")
                return _call(s, args)
            if self.accept("["):
                kk, ss, _ = self.take()
                if kk != "num" or not ss.isdigit():
                    raise ParseError(f"index of {s} must be an integer literal")
                self.expect("]")
                return self.resolve(s, int(ss))
            if s in ("π", "pi", "M_PI"):
                return sp.Float(float(sp.pi))
            return self.resolve(s, None)
        raise ParseError(f"unexpected token {s!r}")


def _number(tok: str) -> sp.Expr:
    if re.fullmatch(r"\d+", tok):
        return sp.Integer(int(tok))
    return sp.Float(float(tok))  # the double nearest to the literal


def _pow(a, b):
    if b.is_Float and float(b) == int(float(b)):
        b = sp.Integer(int(float(b)))  # x^3.0 is the integer power (as the reference models write it)
    return sp.Pow(a, b)


def _select(c, a, b):
    if c is sp.true or c is sp.false:
        return a if c is sp.true else b
    if not getattr(c, "is_Relational", False):
        raise ParseError(f"condition of ifelse / ?: must be a comparison, got {c}")
    return sp.Piecewise((a, c), (b, True))


def _call(name: str, args):
    if name in _UNARY:
        if len(args) != 1:
            raise ParseError(f"{name} takes one argument")
        return _UNARY[name](args[0])
    if name == "pow" and len(args) == 2:
        return _pow(args[0], args[1])
    if name == "inv" and len(args) == 1:
        return 1 / args[0]
    if name == "abs2" and len(args) == 1:
        return args[0] ** 2
    if name == "ifelse" and len(args) == 3:
        return _select(args[0], args[1], args[2])
    if name == "atan2" and len(args) == 2:
        return sp.atan2(args[0], args[1])
    raise ParseError(f"no rule for function {name}/{len(args)}")


def parse_expr(text: str, symbols: Dict[str, sp.Symbol]) -> sp.Expr:
    """One Julia- or C-flavoured expression over named scalars (x1, u1, lam2, ... -- unicode subscripts accepted)."""
    def resolve(name, index):
        if index is not None:
            raise ParseError(f"indexed reference {name}[{index}] in a plain expression")
        if name not in symbols:
            raise ParseError(f"unknown name {name!r}")
        return symbols[name]

    p = _Parser(tokenize(text), resolve)
    v = p.expr()
    if p.peek()[0] != "end":
        raise ParseError(f"trailing input at token {p.peek()[1]!r} in {text[:60]!r}")
    return sp.sympify(v)


def parse_c_function(src: str, args: Dict[str, List[sp.Symbol]]):
    """A Symbolics C-target function -> (name, output name, [expr per output index]).
    `args` maps the C argument names (the `rhsnames` the emitter chose, e.g. y, x, u, w, lam) to their symbol
    vectors; `RHS1[0]`-style references index them zero-based. Statements other than `out[i] = expr;` are
    rejected; outputs that are never assigned are an error (build_function assigns every entry)."""
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)   # block comments (`//` is Julia's rational operator here, not a comment)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)             # #include <math.h>
    m = re.search(r"void\s+(\w+)\s*\(([^)]*)\)\s*\{(.*)\}", src, flags=re.S)
    if not m:
        raise ParseError("no `void name(double* out, const double* a, ...) { ... }` function found")
    fname, params, body = m.group(1), m.group(2), m.group(3)
    pnames = [re.sub(r".*[\s\*]", "", p.strip()) for p in params.split(",") if p.strip()]
    if not pnames:
        raise ParseError("function without parameters")
    out_name, in_names = pnames[0], pnames[1:]
    unknown = [n for n in in_names if n not in args]
    if unknown:
        raise ParseError(f"C arguments {unknown} have no symbol vector (known: {sorted(args)})")

    def resolve(name, index):
        if name not in args or name not in in_names:
            raise ParseError(f"unknown name {name!r} in {fname}")
        if index is None:
            raise ParseError(f"{name} used without an index in {fname}")
        if not 0 <= index < len(args[name]):
            raise ParseError(f"{name}[{index}] out of range (length {len(args[name])})")
        return args[name][index]

    outs: Dict[int, sp.Expr] = {}
    for stmt in body.split(";"):
        if not stmt.strip():
            continue
        toks = tokenize(stmt)
        if len(toks) < 6 or toks[0] != ("name", out_name, toks[0][2]) or toks[1][1] != "[" or toks[3][1] != "]" or toks[4][1] != "=":
            raise ParseError(f"statement is not `{out_name}[i] = expr`: {stmt.strip()[:60]!r}")
        idx = int(toks[2][1])
        p = _Parser(toks[5:], resolve)
        v = p.expr()
        if p.peek()[0] != "end":
            raise ParseError(f"trailing input in {stmt.strip()[:60]!r}")
        if idx in outs:
            raise ParseError(f"{out_name}[{idx}] assigned twice")
        outs[idx] = sp.sympify(v)
    if sorted(outs) != list(range(len(outs))):
        raise ParseError(f"{fname}: outputs {sorted(outs)} are not 0..n-1")
    return fname, out_name, [outs[i] for i in range(len(outs))]
