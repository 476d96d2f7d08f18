"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

Restates, problem by problem in plain Python/numpy, the consumer of the callbacks that the
reference sketches in /root/reference/examples/pendulum/pendulum.jl:109-211:

  * `callbacks`      pendulum.jl:124-136  the five MOI calls at (z, y), sigma = 1.0
  * `assemble`       pendulum.jl:138-199  K = [[H + primal_reg I, C'], [C, -dual_reg I]], h = [grad + C'y; c]
  * `qdldl_*`        pendulum.jl:206-210  F = qdldl(K); solve!(F, sol)

QDLDL.jl is a third-party dependency that is NOT under /root/reference and is not even listed in
examples/Project.toml (un-pinned; the example `using`s whatever is installed). Its published
algorithm (Stellato et al., OSQP, "QDLDL: a free LDL' factorisation routine for quasi-definite
linear systems"; qdldl.c: elimination tree + up-looking numeric factorisation on triu(K), no
pivoting, L_kc = y_c * Dinv_c, D_k -= y_c * L_kc) is restated below from the paper/source as I
recall it. QDLDL.jl applies an AMD ordering by default; an ordering changes rounding only, not the
mathematical factor, so the oracle takes the permutation as an argument (the tests pass the
product's, or the identity). PARITY UNPINNED: the reference holds no stored output of this script
(it uses `rand` inputs and only prints `norm(sol - H \\ h, Inf)`), so the pin is that same check:
the oracle solve against a dense LU solve.
"""
from __future__ import annotations

import numpy as np


def callbacks(nlp, z, y, sigma: float = 1.0):
    """pendulum.jl:124-136."""
    g = np.zeros(nlp.num_variables)
    c = np.zeros(nlp.num_constraint)
    J = np.zeros(nlp.num_jacobian)
    H = np.zeros(len(nlp.hessian_lagrangian_sparsity))
    nlp.eval_objective(z)
    nlp.eval_objective_gradient(g, z)
    nlp.eval_constraint(c, z)
    nlp.eval_constraint_jacobian(J, z)
    nlp.eval_hessian_lagrangian(H, z, sigma, y)
    return g, c, J, H


def assemble(nz, ny, jac_structure, hess_structure, g, c, J, H, y, primal_reg=1.0e-5, dual_reg=1.0e-5):
    """Dense K [nz+ny, nz+ny] and h [nz+ny] (pendulum.jl:138-199). Structures are lists of 1-based
    (row, col) tuples in the reference's order. `C[j, i]` in the script is the constraint Jacobian
    as a matrix: entry (j, i) of the COO triplets."""
    n = nz + ny
    K = np.zeros((n, n))
    h = np.zeros(n)
    Cm = {}
    for k, (r, col) in enumerate(jac_structure):
        Cm[(r, col)] = Cm.get((r, col), 0.0) + float(J[k])
    by_col = {}
    for (r, col) in sorted(Cm):
        by_col.setdefault(col, []).append(r)
    for i in range(1, nz + 1):
        h[i - 1] += g[i - 1]
        cy = 0.0
        for j in by_col.get(i, ()):  # for j = 1:ny if (j, i) in jacobian_sparsity
            cy += Cm[(j, i)] * float(y[j - 1])
        h[i - 1] += cy
    for j in range(ny):
        h[nz + j] += c[j]
    for k, (r, col) in enumerate(hess_structure):
        K[r - 1, col - 1] = H[k]
    for k, (r, col) in enumerate(jac_structure):
        K[nz + r - 1, col - 1] = J[k]
        K[col - 1, nz + r - 1] = J[k]
    for i in range(nz):
        K[i, i] += primal_reg
    for j in range(ny):
        K[nz + j, nz + j] -= dual_reg
    return K, h


# ---------------------------------------------------------------------------- QDLDL restated
def _triu_csc(K):
    n = K.shape[0]
    Ap, Ai, Ax = [0], [], []
    for j in range(n):
        for i in range(j + 1):
            if K[i, j] != 0.0 or i == j:
                Ai.append(i)
                Ax.append(float(K[i, j]))
        Ap.append(len(Ai))
    return Ap, Ai, Ax


def qdldl_factor(K, perm=None):
    """K[perm][:, perm] = L D L'. Returns (L dense unit-lower, D, Dinv, positive_count)."""
    K = np.asarray(K, dtype=np.float64)
    n = K.shape[0]
    if perm is not None:
        perm = np.asarray(perm)
        K = K[np.ix_(perm, perm)]
    Ap, Ai, Ax = _triu_csc(K)
    UNKNOWN = -1
    # elimination tree and column counts
    work = [0] * n
    Lnz = [0] * n
    etree = [UNKNOWN] * n
    for j in range(n):
        work[j] = j
        for p in range(Ap[j], Ap[j + 1]):
            i = Ai[p]
            while work[i] != j:
                if etree[i] == UNKNOWN:
                    etree[i] = j
                Lnz[i] += 1
                work[i] = j
                i = etree[i]
    Lp = [0] * (n + 1)
    for i in range(n):
        Lp[i + 1] = Lp[i] + Lnz[i]
    Li = [0] * Lp[n]
    Lx = [0.0] * Lp[n]
    D = [0.0] * n
    Dinv = [0.0] * n
    used = [False] * n
    yvals = [0.0] * n
    nxt = list(Lp[:n])
    pos = 0
    for k in range(n):
        yidx = []
        for p in range(Ap[k], Ap[k + 1]):
            b = Ai[p]
            if b == k:
                D[k] = Ax[p]
                continue
            yvals[b] = Ax[p]
            if not used[b]:
                used[b] = True
                elim = [b]
                nx = etree[b]
                while nx != UNKNOWN and nx < k:
                    if used[nx]:
                        break
                    used[nx] = True
                    elim.append(nx)
                    nx = etree[nx]
                while elim:
                    yidx.append(elim.pop())
        for c in reversed(yidx):
            yc = yvals[c]
            for q in range(Lp[c], nxt[c]):
                yvals[Li[q]] -= Lx[q] * yc
            Li[nxt[c]] = k
            Lx[nxt[c]] = yc * Dinv[c]
            D[k] -= yc * Lx[nxt[c]]
            nxt[c] += 1
            yvals[c] = 0.0
            used[c] = False
        if D[k] == 0.0:
            raise ZeroDivisionError(f"qdldl: zero pivot at column {k}")
        if D[k] > 0.0:
            pos += 1
        Dinv[k] = 1.0 / D[k]
    Lm = np.eye(n)
    for c in range(n):
        for q in range(Lp[c], Lp[c + 1]):
            Lm[Li[q], c] = Lx[q]
    return Lm, np.array(D), np.array(Dinv), pos


def qdldl_solve(Lm, Dinv, b, perm=None):
    """solve!(F, b): x = P' L^-T D^-1 L^-1 P b."""
    n = len(b)
    x = np.array(b, dtype=np.float64)
    if perm is not None:
        x = x[np.asarray(perm)]
    for i in range(n):
        val = x[i]
        rows = np.nonzero(Lm[i + 1:, i])[0] + i + 1
        for r in rows:
            x[r] -= Lm[r, i] * val
    x = x * Dinv
    for i in range(n - 1, -1, -1):
        val = x[i]
        rows = np.nonzero(Lm[i + 1:, i])[0] + i + 1
        for r in rows:
            val -= Lm[r, i] * x[r]
        x[i] = val
    if perm is not None:
        out = np.empty(n)
        out[np.asarray(perm)] = x
        return out
    return x


def kkt_solve(nlp, z, y, primal_reg=1.0e-5, dual_reg=1.0e-5, perm=None):
    """pendulum.jl:124-210 for one problem: returns dict(K, h, sol, g, c, J, H)."""
    g, c, J, H = callbacks(nlp, z, y)
    K, h = assemble(nlp.num_variables, nlp.num_constraint, nlp.jacobian_structure(), nlp.hessian_lagrangian_structure(),
                    g, c, J, H, y, primal_reg, dual_reg)
    Lm, D, Dinv, _ = qdldl_factor(K, perm)
    sol = qdldl_solve(Lm, Dinv, h, perm)
    return dict(K=K, h=h, sol=sol, g=g, c=c, J=J, H=H, L=Lm, D=D)
