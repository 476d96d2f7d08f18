"""examples/car/car.jl of the reference on the B200 path -- for a whole batch at once.

    python examples/car_obstacle.py [batch]

Reference script (/root/reference/examples/car/car.jl:20-75): midpoint car dynamics, cost u'u, |u| <= 0.5 and both end states
pinned as Bounds, a circular obstacle as one inequality row per knot (Constraint(obs, ...; indices_inequality = [1])), the
guess = interpolated states and controls 0.001 randn, `solve!`, `get_trajectory`. Here the same calls go through `dto_b200`
(exact Hessians on, which the reference's Solver leaves to Ipopt's limited-memory update). Inequalities make `solve()`
pick the interior-point mode of the lock-step Newton-KKT solver (sqp.py; Ipopt is not in this image)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dto_b200 as D  # noqa: E402
from examples import models as M  # noqa: E402


def main(batch=32, T=51):
    model = M.build_car(D, T=T, obstacle="stage")         # car.jl:20-62
    n, m, x1, xT = model["n"], model["m"], model["x1"], model["xT"]
    solver = D.solver_from(model, batch=batch)            # Solver(dynamics, objective, constraints, bounds) (car.jl:65) + the batch size
    # ## initialize (car.jl:68-72): one random control guess per problem
    solver.initialize_states(D.linear_interpolation(x1, xT, T))
    rng = np.random.default_rng(0)
    for b in range(batch):
        solver.initialize_controls([0.001 * rng.normal(size=m) for _ in range(T - 1)], problem=b)
    # ## solve (car.jl:75)
    res = solver.solve(options=dict(max_iter=300))
    # ## solution (car.jl:78-81)
    ok, closest = 0, np.inf
    p_obs, r_obs = np.array(M.P_OBS), M.R_OBS
    for b in range(batch):
        xs, us = solver.get_trajectory(b)
        d = min(np.linalg.norm(x[:2] - p_obs) for x in xs)
        good = (np.linalg.norm(xs[0] - x1) < 1e-3 and np.linalg.norm(xs[-1] - xT) < 1e-3 and bool(res.converged[b])
                and d > r_obs - 1e-6 and max(np.abs(u).max() for u in us) < 0.5)
        ok += int(good)
        if good:
            closest = min(closest, d)
    print(f"{ok} of {batch} paths solved (end points to 1e-3, |u| < 0.5, outside the obstacle, KKT residuals converged); "
          f"closest approach to the obstacle centre {closest:.6f} (radius {r_obs}); kernels launched {solver.sqp_launches}")
    return ok


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32)
