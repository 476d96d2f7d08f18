"""tests/fixtures_ctarget/*.json: the four BASELINE model shapes in the interchange format julia/DTOB200.jl
writes -- one Symbolics-C-target function per element ("evaluate_c": `void dyn_evaluate(double* out, const
double* y, const double* x, ...) { out[0] = ...; }`) plus the structural patterns and the
per-knot kinds. Julia cannot run in this image, so the text is produced by spec_io.c_function, a python
emulation of that printer (Julia Expr printing: spaced binary operators, shortest-repr literals, `a//b`
rationals, `pow(a, b)` for `^`, zero-based `x[i]` references); tests/test_frontend_cpu.py then checks that the
parser turns every fixture back into a model with the SAME content hash as the python front end.
    python tools/make_ctarget_fixtures.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dto_b200 as D  # noqa: E402
from dto_b200 import spec_io  # noqa: E402
from examples import models as M  # noqa: E402

FIXTURES = [("pendulum", dict()), ("cartpole", dict(T=11)), ("acrobot", dict(T=9)), ("car", dict(T=12, obstacle="general")),
            ("piecewise", dict())]


def tag(name, kw):
    return name + "".join(f"_{k}{v}" for k, v in sorted(kw.items()))


def main():
    out = os.path.join(ROOT, "tests", "fixtures_ctarget")
    os.makedirs(out, exist_ok=True)
    for name, kw in FIXTURES:
        mp = M.BUILDERS[name](D, **kw)
        s = D.solver_from(mp, batch=1)
        kd, kc, ks = [k.tolist() for k in s.nlp._keep[:3]]
        doc = spec_io.dump_spec(s.model.spec, dict(T=mp["T"], dynamics_kind=kd, cost_kind=kc, stage_kind=ks), style="ctarget")
        with open(os.path.join(out, tag(name, kw) + ".json"), "w") as f:
            json.dump(doc, f, indent=1)
        print("wrote", tag(name, kw), len(json.dumps(doc)), "bytes")


if __name__ == "__main__":
    main()
