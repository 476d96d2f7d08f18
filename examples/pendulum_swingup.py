"""examples/pendulum/pendulum.jl of the reference (its solve part, :1-98) on the B200 path -- for a whole batch at once.

    python examples/pendulum_swingup.py [batch]

Reference script: midpoint pendulum dynamics, quadratic costs, the state pinned at both ends by stage constraints,
`initialize_states!` with interpolated states, `initialize_controls!`, `solve!`, `get_trajectory`. Equality constraints only:
`solve()` runs the native lock-step Newton-KKT solver inside libdto.so (dto_sqp_solve)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dto_b200 as D  # noqa: E402
from examples import models as M  # noqa: E402


def main(batch=256):
    model = M.build_pendulum(D)
    T, x1, xT = model["T"], model["x1"], model["xT"]
    solver = D.solver_from(model, batch=batch)
    solver.initialize_states(D.linear_interpolation(x1, xT, T))
    rng = np.random.default_rng(0)
    for b in range(batch):
        solver.initialize_controls([rng.normal(size=1) for _ in range(T - 1)], problem=b)
    res = solver.solve()
    ok = 0
    for b in range(batch):
        xs, us = solver.get_trajectory(b)
        ok += int(np.linalg.norm(xs[0] - x1) < 1e-3 and np.linalg.norm(xs[-1] - xT) < 1e-3 and bool(res.converged[b]))
    print(f"{ok} of {batch} swing-ups solved; median iterations {float(np.median(res.iterations)):.0f}; "
          f"kernels launched {solver.sqp_launches}; objective of problem 0: {float(res.objective[0]):.6f}")
    return ok


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
