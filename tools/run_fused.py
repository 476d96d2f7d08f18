"""Launch the fused Jacobian+Hessian pass of one model a few times (target command for ncu).
    python tools/run_fused.py acrobot:T=101:4096 [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import dto_b200 as D  # noqa: E402
from dto_b200.evaluator import K_JAC_HESS  # noqa: E402
from examples import models as M  # noqa: E402
from util import make_inputs  # noqa: E402

parts = sys.argv[1].split(":")
name, kw, B = parts[0], {}, int(parts[2]) if len(parts) > 2 else 4096
if len(parts) > 1 and parts[1]:
    for kv in parts[1].split(","):
        k, v = kv.split("=")
        kw[k] = int(v) if v.isdigit() else v
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
model = M.BUILDERS[name](D, **kw)
s = D.solver_from(model, batch=B)
n0 = s.nlp
nl = [n0, n0.new_batch(), n0.new_batch()]
for i, n in enumerate(nl):
    z, lam, sig, w = make_inputs(name, model, n.num_variables, n.num_constraint, n.num_parameter, B, 2, i)
    if n.num_parameter:
        n.set_parameters(w)
    n.set_x(z)
    n.set_duals(sig, lam)
for i in range(iters):
    nl[i % 3].launch(K_JAC_HESS)
for n in nl:
    n.sync()
print("done", name, kw, B, "smem", n0.kernel_smem_bytes(5))
