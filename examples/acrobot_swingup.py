"""examples/acrobot/acrobot.jl of the reference, line for line, on the B200 path -- for a whole batch at once.

    python examples/acrobot_swingup.py [batch]

Reference script (/root/reference/examples/acrobot/acrobot.jl:94-139): build Dynamics / Cost / Constraint objects,
`Solver(dynamics, objective, constraints, bounds)`, `initialize_states!`, `initialize_controls!`, `solve!`,
`get_trajectory`. Here the same calls go through `dto_b200` (the python mirror of that API); `batch` problems --
each with its own random control guess, as the example's `randn` would give on every run -- are solved in lock step on
the device (Ipopt is not in this image)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dto_b200 as D  # noqa: E402
from examples import models as M  # noqa: E402


def main(batch=64, T=101):
    n, m = 4, 1
    x1 = np.array([0.0, 0.0, 0.0, 0.0])
    xT = np.array([np.pi, 0.0, 0.0, 0.0])
    # ## model / objective / constraints (acrobot.jl:93-118)
    dt = D.Dynamics(M.acrobot_midpoint, n, n, m, num_parameter=0, evaluate_hessian=True)
    ot = lambda x, u, w: 0.1 * M.dot(x[2:4], x[2:4]) + 0.1 * M.dot(u, u)  # noqa: E731
    oT = lambda x, u, w: 0.1 * M.dot(x[2:4], x[2:4])  # noqa: E731
    ct, cT = D.Cost(ot, n, m, evaluate_hessian=True), D.Cost(oT, n, 0, evaluate_hessian=True)
    cons = [D.Constraint(lambda x, u, w: x - x1, n, m, evaluate_hessian=True)] + [D.Constraint() for _ in range(2, T)] + \
           [D.Constraint(lambda x, u, w: x - xT, n, 0, evaluate_hessian=True)]
    bounds = [D.Bound(n, m)] * (T - 1) + [D.Bound(n, 0)]
    # ## problem (acrobot.jl:121-123) -- plus the batch size
    solver = D.Solver([dt] * (T - 1), [ct] * (T - 1) + [cT], cons, bounds, evaluate_hessian=True, batch=batch, name="acrobot")
    # ## initialize (acrobot.jl:126-131): one random control guess per problem
    rng = np.random.default_rng(0)
    solver.initialize_states(D.linear_interpolation(x1, xT, T))
    for b in range(batch):
        solver.initialize_controls([rng.normal(size=m) for _ in range(T - 1)], problem=b)
    # ## solve (acrobot.jl:134)
    res = solver.solve(options=dict(max_iter=300))
    # ## solution (acrobot.jl:137-140)
    ok = 0
    for b in range(batch):
        xs, us = solver.get_trajectory(b)
        ok += int(np.linalg.norm(xs[0] - x1) < 1e-3 and np.linalg.norm(xs[-1] - xT) < 1e-3 and bool(res.converged[b]))
    print(f"{ok} of {batch} swing-ups solved (||x_1 - x1||, ||x_T - xT|| < 1e-3, KKT residuals converged); "
          f"median iterations {float(np.median(res.iterations)):.0f}; kernels launched {solver.sqp_launches}")
    xs, us = solver.get_trajectory(0)
    print("x_1 =", xs[0], " x_T =", xs[-1], " max |u| =", max(abs(u[0]) for u in us))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
