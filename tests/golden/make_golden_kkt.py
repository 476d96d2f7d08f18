"""Generates tests/golden/kkt_*.npz: for the seeded inputs of the callback fixtures, the oracle's
restatement of /root/reference/examples/pendulum/pendulum.jl:124-211 -- assembled K (dense), h and
sol = K \\ h (QDLDL restated, identity ordering) with the script's regularisation 1e-5 / 1e-5 and
sigma = 1.0. The reference stores no output of that script (PARITY UNPINNED, see oracle/kkt.py);
these files freeze the oracle.

    python tests/golden/make_golden_kkt.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from examples import models as M  # noqa: E402
from oracle import api as O  # noqa: E402
from oracle import kkt as OK  # noqa: E402
from util import oracle_parameters  # noqa: E402



def tag(name, kw):  # same naming as make_golden.py
    return name + "".join(f"_{k}{v}" for k, v in sorted(kw.items()))


KKT_GOLDEN = [
    ("pendulum", dict(), 1),
    ("cartpole", dict(T=11), 2),
    ("acrobot", dict(T=9), 3),
    ("car", dict(T=12, obstacle="general"), 4),
]
REG = 1.0e-5


def main():
    for name, kw, config in KKT_GOLDEN:
        fx = np.load(os.path.join(HERE, tag(name, kw) + ".npz"))
        model = M.BUILDERS[name](O, **kw)
        solver = O.solver_from(model)
        Ks, hs, sols, refs = [], [], [], []
        for b in range(fx["z"].shape[0]):
            p = oracle_parameters(model, fx["w"][b])
            if p is not None:
                solver.set_parameters(p)
            r = OK.kkt_solve(solver.nlp, fx["z"][b], fx["lam"][b], REG, REG)
            Ks.append(r["K"]); hs.append(r["h"]); sols.append(r["sol"])
            # accurate reference: dense LU with partial pivoting + one step of iterative refinement (the
            # natural-ordering no-pivot factorisation above loses up to 5e-7 on the acrobot systems,
            # which are not quasi-definite for random multipliers)
            x = np.linalg.solve(r["K"], r["h"])
            x = x + np.linalg.solve(r["K"], r["h"] - r["K"] @ x)
            refs.append(x)
        path = os.path.join(HERE, "kkt_" + tag(name, kw) + ".npz")
        np.savez_compressed(path, K=np.array(Ks), h=np.array(hs), sol=np.array(sols), sol_lu=np.array(refs), reg=REG)
        print("wrote", path, np.array(Ks).shape)


if __name__ == "__main__":
    main()
