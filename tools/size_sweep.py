"""Device-timed fused Jacobian+Hessian pass over a list of (model, T, B): config-5 style size sweep, one JSON
line per case.   python tools/size_sweep.py cartpole:T=101:4096 cartpole:T=101:65536 ..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import dto_b200 as D  # noqa: E402
from dto_b200.evaluator import K_JAC_HESS  # noqa: E402
from examples import models as M  # noqa: E402
from util import make_inputs  # noqa: E402

for spec in sys.argv[1:]:
    parts = spec.split(":")
    name, kw, B = parts[0], {}, int(parts[2])
    for kv in parts[1].split(","):
        if kv:
            k, v = kv.split("=")
            kw[k] = int(v) if v.isdigit() else v
    model = M.BUILDERS[name](D, **kw)
    s = D.solver_from(model, batch=B)
    n0 = s.nlp
    byt = n0.algorithmic_bytes_per_problem() * B
    R = 3 if byt < 4e9 else 1
    nl = [n0] + [n0.new_batch() for _ in range(R - 1)]
    st = torch.cuda.Stream()
    for i, n in enumerate(nl):
        z, lam, sig, w = make_inputs(name, model, n.num_variables, n.num_constraint, n.num_parameter, B, 2, i)
        if n.num_parameter:
            n.set_parameters(w)
        n.set_x(z)
        n.set_duals(sig, lam)
        n.set_stream(st.cuda_stream)
    steps = max(5, min(200, int(2e11 / byt)))
    with torch.cuda.stream(st):
        for i in range(5):
            nl[i % R].launch(K_JAC_HESS)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            nl[i % R].launch(K_JAC_HESS)
        e1.record(st)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"model": name, **kw, "B": B, "ms": ms, "GBs": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / 6549.4,
                      "evals_per_s": B * n0.T / ms * 1e3, "steps": steps}), flush=True)
    for n in nl:
        n.close()
