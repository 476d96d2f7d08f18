"""examples/cartpole/cartpole.jl of the reference on the B200 path -- for a whole batch at once.

    python examples/cartpole_swingup.py [batch]

Reference script (/root/reference/examples/cartpole/cartpole.jl:41-115): implicit RK3 dynamics, quadratic costs, the state
pinned at both ends by stage constraints, |u| <= 3 as Bound(action_lower, action_upper), the guess = constant controls
0.01 and the states of an explicit rollout, `solve!`, `get_trajectory`. Here the same calls go through `dto_b200`; the
model is BASELINE config 2's (start and goal enter as per-problem parameters w = [x1; xT]). Bounds on variables make
`solve()` pick the interior-point mode of the lock-step Newton-KKT solver (sqp.py; Ipopt is not in this
image): all problems advance together on the device, the barrier term sits on the diagonal of H inside the factor kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dto_b200 as D  # noqa: E402


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)

from examples import models as M  # noqa: E402


def main(batch=64, T=101):
    model = M.build_cartpole(D, T=T)                      # cartpole.jl:41-94 (u_bnd = 3.0)
    n, m, x1, xT = model["n"], model["m"], model["x1"], model["xT"]
    solver = D.solver_from(model, batch=batch)            # Solver(dyn, obj, cons, bounds) (cartpole.jl:98-99) + the batch size
    solver.nlp.set_parameters(np.tile(np.concatenate([x1, xT]), (batch, 1)))
    # ## initialize (cartpole.jl:102-109); every problem gets its own constant control near the example's 0.01
    rng = np.random.default_rng(0)
    for b in range(batch):
        u0 = np.array([0.01 * (1.0 + 0.2 * rng.normal())])
        xs = [x1.astype(float)]
        for _ in range(T - 1):
            xs.append(np.array(M.cartpole_rk3_explicit(xs[-1], u0, np.zeros(0)), dtype=float))
        solver.initialize_states(xs, problem=b)
        solver.initialize_controls([u0] * (T - 1), problem=b)
    # ## solve (cartpole.jl:112)
    res = solver.solve(options=dict(max_iter=600))
    # ## solution (cartpole.jl:115-118)
    ok, umax = 0, 0.0
    for b in range(batch):
        xs, us = solver.get_trajectory(b)
        umax = max(umax, max(abs(u[0]) for u in us))
        ok += int(np.linalg.norm(xs[0] - x1) < 1e-3 and np.linalg.norm(xs[-1] - xT) < 1e-3 and bool(res.converged[b]))
    print(f"{ok} of {batch} swing-ups solved (||x_1 - x1||, ||x_T - xT|| < 1e-3, KKT residuals converged); max |u| = {umax:.6f} <= 3; "
          f"median iterations {float(np.median(_np(res.iterations))):.0f}; kernels launched {solver.sqp_launches}")
    xs, us = solver.get_trajectory(0)
    print("x_1 =", xs[0], " x_T =", xs[-1])
    return ok


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
