"""Interior-point mode of the lock-step solver on the reference's car example (examples/car/car.jl: T = 51, |u| <= 0.5, pinned end
states, one obstacle inequality row per knot; the example's guess: states interpolated, controls 0.001 randn):
    python tools/ip_car.py [B] [max_iter] [T] [stage|general]  -> JSON line
(`general` with T = 201 is BASELINE config 4's shape: the same inequalities as ONE GeneralConstraint over the whole z.)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dto_b200 as D  # noqa: E402


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)

from examples import models as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
T = int(sys.argv[3]) if len(sys.argv) > 3 else 51
obstacle = sys.argv[4] if len(sys.argv) > 4 else "stage"
model = M.build_car(D, T=T, obstacle=obstacle)
n, m, x1, xT = model["n"], model["m"], model["x1"], model["xT"]
s = D.solver_from(model, batch=B)
s.initialize_states(D.linear_interpolation(x1, xT, T))
rng = np.random.default_rng(2)
for b in range(B):
    s.initialize_controls([0.001 * rng.normal(size=m) for _ in range(T - 1)], problem=b)
so = dict(max_iter=iters)
so.update(json.loads(os.environ.get("DTO_SQP_OPTIONS", "{}")))
t0 = time.perf_counter()
res = s.solve(options=so)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
conv, it = _np(res.converged), _np(res.iterations)
Z = _np(res.z)
c = np.zeros((B, s.nlp.num_constraint))
s.nlp.eval_constraint(c, Z)
clo, cup = s.nlp.constraint_bounds
ineq = clo != cup
U = np.concatenate([Z[:, t * (n + m) + n: t * (n + m) + n + m] for t in range(T - 1)], axis=1)
print(json.dumps(dict(options=so, T=T, obstacle=obstacle, B=B, seconds=dt, converged=float(conv.mean()), staged=float(_np(res.staged).mean()) if hasattr(res, "staged") else 0.0,
                      it_median=float(np.median(it)), it_max=float(it.max()), eq_violation_max_converged=float(np.abs(c[conv][:, ~ineq]).max()) if conv.any() else None,
                      obstacle_max_converged=float(c[conv][:, ineq].max()) if conv.any() else None, u_max=float(np.abs(U).max()),
                      at_bound=float((np.abs(U) > 0.4999).mean()), f_median=float(np.median(_np(res.objective))), launches=int(s.sqp_launches))))
s.nlp.close()
