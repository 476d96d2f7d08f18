"""End-to-end host call (pinned host z, sigma, lambda -> host J, H) against the number of pipeline chunks.
    python tools/e2e_chunks.py [B]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dto_b200 as D  # noqa: E402
from examples import models as M  # noqa: E402
from util import make_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = M.build_cartpole(D, T=101)
n = D.solver_from(model, batch=B).nlp
z, lam, sigma, w = make_inputs("cartpole", model, n.num_variables, n.num_constraint, n.num_parameter, B, 2)
n.set_parameters(w)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
zp, lp, sp_ = pin(z), pin(lam), pin(sigma)
Jp = torch.empty((B, n.num_jacobian), dtype=torch.float64).pin_memory().numpy()
Hp = torch.empty((B, n.num_hessian), dtype=torch.float64).pin_memory().numpy()
byt = 8 * B * (n.num_variables + n.num_constraint + 1 + n.num_jacobian + n.num_hessian)
for chunks in (1, 2, 4, 8, 12, 16, 24, 32, 64):
    for _ in range(3):
        n.eval_jacobian_hessian(Jp, Hp, zp, sp_, lp, chunks=chunks)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        n.eval_jacobian_hessian(Jp, Hp, zp, sp_, lp, chunks=chunks)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / reps
    print(json.dumps({"chunks": chunks, "ms": ms, "GBs": byt / ms / 1e6, "knot_evals_per_s": B * 101 / ms * 1e3}), flush=True)
