"""Small native solve that exercises every branch of dto_sqp_solve -- inertia ladder in candidate slots, second-order
correction through the stored factor, one-pass line search over trial slots -- for compute-sanitizer:
    compute-sanitizer --tool memcheck|initcheck|racecheck python tools/sanitize_sqp.py [B] [iterations]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import dto_b200 as D  # noqa: E402
from dto_b200 import sqp  # noqa: E402
from examples import models as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ma = M.build_acrobot(D, T=101, stage_endpoint_constraints=False)     # end points pinned by bounds (test/solve.jl:1-138)
pn = D.solver_from(ma, batch=B).nlp
T, n, m = ma["T"], ma["n"], ma["m"]
rng = np.random.default_rng(12)
z0 = np.zeros((B, T * n + (T - 1) * m))
for t in range(T):
    o = t * (n + m)
    z0[:, o:o + n] = ma["x1"] + (ma["xT"] - ma["x1"]) * t / (T - 1)
    if t < T - 1:
        z0[:, o + n:o + n + m] = rng.normal(size=(B, m))
res = sqp.solve_native(pn, z0, options=sqp.SQPOptions(max_iter=iters))
print({k: v for k, v in res.stats.items() if k != "phase_ms"})
assert res.stats["refactorisations"] > 0 and res.stats["corrections"] > 0 and res.stats["multi_trial_passes"] > 0
pn.close()
