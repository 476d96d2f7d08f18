"""Interior-point mode of sqp.solve on the cartpole example (examples/cartpole/cartpole.jl: T = 101, |u| <= u_bnd, the
example's guess: constant controls 0.01 and the states of an explicit rollout; GUESS=interpolate for interpolated states and
controls GUESS_SIGMA * randn):   python tools/ip_cartpole.py [u_bnd] [B] [max_iter]  -> JSON line"""
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

def run(ub=3.0, B=32, iters=400, guess=None, sigma=None, distinct=None):
    """`distinct`: number of different guesses (rollouts are python loops); problem b uses guess b % distinct"""
    T = 101
    mc = M.build_cartpole(D, T=T)
    n, m = mc["n"], mc["m"]
    mc["bounds"] = [D.Bound(n, m, action_lower=[-ub], action_upper=[ub])] * (T - 1) + [D.Bound(n, 0)]   # host side: same model library
    s = D.solver_from(mc, batch=B)
    s.nlp.set_parameters(np.tile(np.concatenate([mc["x1"], mc["xT"]]), (B, 1)))      # w = [x1; xT] of every problem (BASELINE config 2 layout)
    rng = np.random.default_rng(7)
    guess = guess or os.environ.get("GUESS", "rollout")
    sigma = float(os.environ.get("GUESS_SIGMA", "0.01")) if sigma is None else sigma
    if guess == "rollout":
        # the example's guess (cartpole.jl:102-109): constant controls 0.01 (here 0.01 (1 + 0.2 randn) per problem) and the states
        # of an explicit RK3 rollout from x1
        rolls = []
        for _ in range(min(B, distinct or B)):
            u0 = np.array([0.01 * (1.0 + 0.2 * rng.normal())])
            xs = [mc["x1"].astype(float)]
            for t in range(T - 1):
                xs.append(np.array(M.cartpole_rk3_explicit(xs[-1], u0, np.zeros(0)), dtype=float))
            rolls.append((u0, xs))
        for b in range(B):
            u0, xs = rolls[b % len(rolls)]
            s.initialize_states(xs, problem=b)
            s.initialize_controls([u0 for _ in range(T - 1)], problem=b)
    else:   # states interpolated from x1 to xT, controls sigma * randn
        s.initialize_states(D.linear_interpolation(mc["x1"], mc["xT"], T))
        for b in range(B):
            s.initialize_controls([sigma * rng.normal(size=1) for _ in range(T - 1)], problem=b)
    so = dict(max_iter=iters)
    so.update(json.loads(os.environ.get("DTO_SQP_OPTIONS", "{}")))
    t0 = time.perf_counter()
    res = s.solve(options=so, method=os.environ.get("METHOD", "auto"))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    it, conv = _np(res.iterations), _np(res.converged)
    Z = _np(res.z)
    U = np.stack([Z[:, t * (n + m) + n] for t in range(T - 1)], axis=1)
    out = (dict(u_bnd=ub, guess=guess, guess_sigma=sigma, options=so, T=T, B=B, seconds=dt, converged=float(conv.mean()), it_median=float(np.median(it)), it_max=float(it.max()),
                          cv_max=float(_np(res.constraint_violation).max()), dr_median=float(np.median(_np(res.dual_residual))),
                          u_max=float(np.abs(U).max()), end_error_max=float(np.abs(Z[:, -n:] - mc["xT"]).max()), at_bound=float((np.abs(U) > 0.997 * ub).mean()), f_median=float(np.median(_np(res.objective)))))
    s.nlp.close()
    out["solves_per_s"] = B / dt
    return out


if __name__ == "__main__":
    a = sys.argv[1:]
    print(json.dumps(run(float(a[0]) if a else 3.0, int(a[1]) if len(a) > 1 else 32, int(a[2]) if len(a) > 2 else 400)))
