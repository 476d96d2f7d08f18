"""BASELINE config 3 (acrobot T = 101 full solves) over several devices: one dto_sqp_solve per device side by side (no collective).
    python tools/solve_multi_gpu.py [B_total] [max_iter]   -> one JSON line per device count 1, 2, .., all visible (strong scaling)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dto_b200 as D  # noqa: E402
from examples import models as M  # noqa: E402
from solve_config3 import initial_guess  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ndev = torch.cuda.device_count()
model = M.build_acrobot(D, T=101)
z0 = initial_guess(model, B)
n = model["n"]
counts = sorted({1, ndev} | {c for c in (2, 4, 8) if c <= ndev})
for g in counts:
    s = D.solver_from(model, batch=B, devices=list(range(g)))
    s._initial[:] = z0
    for rep in range(2):                      # the first pass warms every device up (context, plan tables, allocations)
        t0 = time.perf_counter()
        res = s.solve(options=dict(max_iter=iters if rep else 2))
        dt = time.perf_counter() - t0
    Z = np.asarray(res.z)
    ok = (np.asarray(res.constraint_violation) < 1e-6) & (np.linalg.norm(Z[:, :n] - model["x1"], axis=1) < 1e-3) & \
         (np.linalg.norm(Z[:, -n:] - model["xT"], axis=1) < 1e-3)
    print(json.dumps(dict(workload=f"acrobot swing-up T=101, {B} problems, full solves (dto_sqp_solve per device)", n_gpus=g, B=B, seconds=dt,
                          solves_per_s=B / dt, accepted_frac=float(ok.mean()), iterations_median=float(np.median(np.asarray(res.iterations))))))
    s.nlp.close()
