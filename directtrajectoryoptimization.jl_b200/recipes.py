"""Compile-time Hessian gather recipes.

The Hessian-of-Lagrangian value of one slot is the sum, in the reference's += order
(/root/reference/src/moi.jl:88-111: cost_t, dynamics_{t-1} (its next-state rows), dynamics_t,
stage_t), of at most four element terms. Which terms feed which slot of knot t's rows depends
only on the local sparsity patterns of the elements around knot t, so the per-knot "recipe"
repeats along the horizon: a T=101 cartpole has three distinct recipes (first, interior, last
knot). The code generator assembles the shape's structure here exactly like the runtime does
(csrc/dto_runtime.cpp, restating /root/reference/src/data.jl:178-184), dedupes the per-knot
recipes into CLASSES and emits one straight-line gather function per class, so the kernel needs
no per-slot table loads. The runtime recomputes every knot's recipe from its own assembly and
matches it against the classes the model library carries (bit-exact), falling back to the
table-driven gather when a shape has recipes the library was not generated for.

Encoding of one source: k >= 0 -> term k of the knot's own term block ([cost][dynamics][stage]);
k <= -2 -> term (-k-2) of the PREVIOUS knot's term block; -1 -> none.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

Recipe = Tuple[Tuple[int, int, int, int], ...]  # per slot of the knot: 4 encoded sources


def knot_recipes(T: int, dyn_kind: Sequence[int], cost_kind: Sequence[int], stage_kind: Sequence[int], dyn, cost, stage,
                 general=None) -> Optional[List[Recipe]]:
    """dyn/cost/stage: lists of ElementSpec by kind; returns one Recipe per knot (None if a slot
    would need more than four knot terms or an element reaches outside its window)."""
    nx = [dyn[dyn_kind[t]].nx for t in range(T - 1)] + [dyn[dyn_kind[T - 2]].n_out]
    nu = [dyn[dyn_kind[t]].nu for t in range(T - 1)] + [0]
    zofs = [0]
    for t in range(T):
        zofs.append(zofs[-1] + nx[t] + nu[t])
    nh_c = [cost[cost_kind[t]].nnz_hess for t in range(T)]
    nh_d = [dyn[dyn_kind[t]].nnz_hess if t < T - 1 else 0 for t in range(T)]
    nh_s = [stage[stage_kind[t]].nnz_hess if stage_kind[t] >= 0 else 0 for t in range(T)]
    hterm = [0]
    for t in range(T):
        hterm.append(hterm[-1] + nh_c[t] + nh_d[t] + nh_s[t])
    # terms in the reference's concatenation order: objective, dynamics, stage (src/data.jl:178-182)
    terms: List[Tuple[int, int, int, int]] = []  # (row, col, knot, local term index)
    for t in range(T):
        e = cost[cost_kind[t]]
        for k in range(nh_c[t]):
            terms.append((zofs[t] + e.hess_rows[k], zofs[t] + e.hess_cols[k], t, k))
    for t in range(T - 1):
        e = dyn[dyn_kind[t]]
        for k in range(nh_d[t]):
            terms.append((zofs[t] + e.hess_rows[k], zofs[t] + e.hess_cols[k], t, nh_c[t] + k))
    for t in range(T):
        if stage_kind[t] >= 0:
            e = stage[stage_kind[t]]
            for k in range(nh_s[t]):
                terms.append((zofs[t] + e.hess_rows[k], zofs[t] + e.hess_cols[k], t, nh_c[t] + nh_d[t] + k))
    keys = {(r, c) for (r, c, _, _) in terms}
    if general is not None and general.has_hess:
        keys |= set(zip(general.hess_rows, general.hess_cols))
    key = sorted(keys)
    slot = {rc: i for i, rc in enumerate(key)}
    srcs: List[List[int]] = [[] for _ in key]
    owner_of_row = {}
    for t in range(T):
        for r in range(zofs[t] + 1, zofs[t + 1] + 1):
            owner_of_row[r] = t
    for (r, c, t, k) in terms:
        o = owner_of_row.get(r)
        if o is None:
            return None
        if o == t:
            enc = k
        elif o == t + 1:
            enc = -2 - k
        else:
            return None
        srcs[slot[(r, c)]].append(enc)
    out: List[List[Tuple[int, int, int, int]]] = [[] for _ in range(T)]
    for i, (r, c) in enumerate(key):
        s = srcs[i]
        if len(s) > 4:
            return None
        out[owner_of_row[r]].append(tuple(s + [-1] * (4 - len(s))))
    return [tuple(x) for x in out]


def knot_meta(T: int, dyn_kind: Sequence[int], cost_kind: Sequence[int], stage_kind: Sequence[int], dyn, cost, stage
              ) -> List[Tuple[int, int, int]]:
    """Per knot: (own cost terms, own dynamics terms, cost terms of the previous knot). With these a flat
    term index of a recipe splits into (role, index inside the element), which is what the
    register-resident gather of the warp-specialised kernel needs."""
    nh_c = [cost[cost_kind[t]].nnz_hess for t in range(T)]
    nh_d = [dyn[dyn_kind[t]].nnz_hess if t < T - 1 else 0 for t in range(T)]
    return [(nh_c[t], nh_d[t], nh_c[t - 1] if t > 0 else 0) for t in range(T)]


def classes_of(recipes: Sequence[Recipe], meta: Optional[Sequence[Tuple[int, int, int]]] = None):
    """Distinct (recipe, meta) pairs in order of first appearance, and the class id of every knot.
    Returns (classes, ids) without meta, (classes, metas, ids) with."""
    uniq: List[Recipe] = []
    metas: List[Tuple[int, int, int]] = []
    index: Dict[tuple, int] = {}
    ids = []
    for t, r in enumerate(recipes):
        key = (r, tuple(meta[t])) if meta is not None else (r,)
        if key not in index:
            index[key] = len(uniq)
            uniq.append(r)
            if meta is not None:
                metas.append(tuple(meta[t]))
        ids.append(index[key])
    if meta is None:
        return uniq, ids
    return uniq, metas, ids
