#!/usr/bin/env python
"""Times the device-resident KKT consumer (dto_kkt_*) for one model shape:

  kernels   right-hand side + banded LDL' factor/solve on the J, H, g, c already in HBM (CUDA events)
  step      gradient + constraint + fused Jacobian/Hessian callbacks + the two KKT kernels (CUDA events)
  e2e       host z, lambda (pinned) -> set_x/set_duals -> dto_kkt_solve -> host sol (wall clock), next to
            the e2e of shipping J and H to the host instead (what a CPU factorisation would need)

    python tools/run_kkt.py --model cartpole --T 101 --batch 4096 [--steps 50]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="cartpole")
    ap.add_argument("--T", type=int, default=101)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--obstacle", default="general")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch

    import dto_b200 as D
    from dto_b200 import kkt as PK
    from examples import models as M
    from util import make_inputs

    kw = dict(T=args.T)
    if args.model == "car":
        kw["obstacle"] = args.obstacle
    if args.model == "pendulum":
        kw = {}
    model = M.BUILDERS[args.model](D, **kw)
    B = args.batch
    nlp = D.solver_from(model, batch=B, devices=[0]).nlp
    config = {"pendulum": 1, "cartpole": 2, "acrobot": 3, "car": 4}.get(args.model, 0)
    z, lam, sigma, w = make_inputs(args.model, model, nlp.num_variables, nlp.num_constraint, nlp.num_parameter, B, config=config)
    sigma = np.ones(B)
    if nlp.num_parameter:
        nlp.set_parameters(w)
    nlp.set_x(z)
    nlp.set_duals(sigma, lam)
    stream = torch.cuda.Stream()
    nlp.set_stream(stream.cuda_stream)
    kkt = PK.KKTSystem(nlp)
    n = kkt.dim

    def timed(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    kkt.launch(True)
    torch.cuda.synchronize()
    ms_k = timed(lambda: kkt.launch(False), args.steps)
    ms_s = timed(lambda: kkt.launch(True), args.steps)
    ms_cb = timed(lambda: (nlp.launch(1), nlp.launch(2), nlp.launch(5)), args.steps)

    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    zp, lp, sp_ = pin(z), pin(lam), pin(sigma)
    solp = torch.empty((B, n), dtype=torch.float64).pin_memory().numpy()
    Jp = torch.empty((B, nlp.num_jacobian), dtype=torch.float64).pin_memory().numpy()
    Hp = torch.empty((B, nlp.num_hessian), dtype=torch.float64).pin_memory().numpy()
    gp = torch.empty((B, nlp.num_variables), dtype=torch.float64).pin_memory().numpy()
    cp = torch.empty((B, nlp.num_constraint), dtype=torch.float64).pin_memory().numpy()

    chunks_now = [0]

    def e2e_kkt():
        kkt.solve(solp, variables=zp, scaling=sp_, duals=lp, chunks=chunks_now[0])

    def e2e_ship():
        nlp.eval_jacobian_hessian(Jp, Hp, zp, sp_, lp)
        nlp.eval_objective_gradient(gp)
        nlp.eval_constraint(cp)

    res = {}
    runs = [("ship_JH", e2e_ship, 0)] + [(f"kkt_chunks{c}", e2e_kkt, c) for c in (1, 2, 4, 8, 16)]
    for name, fn, c in runs:
        chunks_now[0] = c
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        reps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        res[name] = 1e3 * (time.perf_counter() - t0) / reps

    # bytes the two KKT kernels must move per problem: read J, H, g, c, y once; write + read the
    # factor once; write h and sol, read h
    alg = 8 * (nlp.num_jacobian + nlp.num_hessian + nlp.num_variables + 2 * nlp.num_constraint + 3 * n) + 2 * kkt.factor_bytes_per_problem
    peak = 6549.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    line = {
        "tool": "run_kkt", "tag": args.tag, "model": args.model, "T": args.T, "B": B, "dim": n, "bandwidth": kkt.bandwidth,
        "row_width": kkt.row_width, "factor_bytes_per_problem": kkt.factor_bytes_per_problem,
        "ms_kkt_kernels": ms_k, "ms_callbacks": ms_cb, "ms_step_callbacks_plus_kkt": ms_s,
        "kkt_solves_per_s": B / (ms_k * 1e-3), "steps_per_s": B / (ms_s * 1e-3),
        "bytes_per_problem_kkt_kernels": alg, "kkt_GBs": alg * B / (ms_k * 1e-3) / 1e9, "kkt_frac_of_hbm_peak": alg * B / (ms_k * 1e-3) / 1e9 / peak,
        "e2e_ms_kkt_solution_to_host": min(v for k_, v in res.items() if k_.startswith("kkt_")),
        "e2e_ms_kkt_by_chunks": {k_[10:]: v for k_, v in res.items() if k_.startswith("kkt_")},
        "e2e_ms_ship_g_c_J_H_to_host": res["ship_JH"],
        "d2h_bytes_kkt": 8 * B * n, "d2h_bytes_ship": 8 * B * (nlp.num_jacobian + nlp.num_hessian + nlp.num_variables + nlp.num_constraint),
    }
    print(json.dumps(line))
    kkt.close()
    nlp.close()


if __name__ == "__main__":
    main()
