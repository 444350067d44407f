#!/usr/bin/env python
"""Cost / parity trade-off of the solver options on one GPU: for every option set, (i) the cube24 fixture of
tests/golden (10 oracle steps, reference settings) is re-run and the relative L2 errors are recorded, (ii) the
benchmark cavity (n = 74) is stepped and timed.  One JSON line per option set.
  python tools/tune_newton.py [n=74] [steps=8]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flow_b200 import _lib  # noqa: E402
from flow_b200 import dolfin as d  # noqa: E402
from flow_b200 import navier_stokes as nav  # noqa: E402
from flow_b200._lib import lib  # noqa: E402

SETS = {
    "default": {},
    "cheb3": {"chebyshev_degree": 3},
    "cheb5": {"chebyshev_degree": 5},
    "cheb6": {"chebyshev_degree": 6},
    "rtol1e-5": {"momentum_rtol": 1e-5},
    "rtol1e-4": {"momentum_rtol": 1e-4},
    "rtol1e-4_cheb5": {"momentum_rtol": 1e-4, "chebyshev_degree": 5},
    "inner_cg4": {"inner_chebyshev": 0},
    "semi_implicit": {"semi_implicit": 1},
}


def cavity(n):
    mesh = d.UnitCubeMesh(n, n, n)
    W = d.VectorFunctionSpace(mesh, "CG", 2)
    P = d.FunctionSpace(mesh, "CG", 1)
    bcs = [d.DirichletBC(W, (0.0, 0.0, 0.0), "on_boundary"), d.DirichletBC(W, (1.0, 0.0, 0.0), lambda x, on: x[2] > 1 - 1e-12)]
    return mesh, W, P, bcs


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 74
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    names = sys.argv[3].split(",") if len(sys.argv) > 3 else list(SETS)
    fixture = sys.argv[4] if len(sys.argv) > 4 else "cube24"
    fx = np.load(os.path.join(ROOT, "tests", "golden", "parity_%s.npz" % fixture))
    zero = d.Constant((0.0, 0.0, 0.0))
    small = cavity(int(fx["n"]))
    big = cavity(n) if steps > 0 else None
    for name in names:
        nav.reset_options()
        nav.set_options(**SETS[name])
        rec = {"options": name}
        # parity on the fixture
        mesh, W, P, bcs = small
        u, p = d.Function(W), d.Function(P)
        errs = {}
        for k in range(1, 11):
            u, p = nav.IPCS().step(d.Constant(1e-2), {0: u}, p, bcs, [], d.Constant(1.0), d.Constant(1e-2), {0: zero, 1: zero}, verbose=False, tol=1e-10)
            if ("u_%d" % k) in fx:
                pv = p._vec - p._vec.mean()
                errs[k] = (float(np.linalg.norm(u._vec[fx["iu"]] - fx["u_%d" % k]) / np.linalg.norm(fx["u_%d" % k])),
                           float(np.linalg.norm(pv[fx["ip"]] - fx["p_%d" % k]) / np.linalg.norm(fx["p_%d" % k])))
        rec["cube24_err_u_p"] = errs
        rec["fixture"] = fixture
        if big is None:
            print(json.dumps(rec), flush=True)
            continue
        # timing at size
        mesh, W, P, bcs = big
        u, p = d.Function(W), d.Function(P)
        ms, its = [], []
        for k in range(3 + steps):
            u, p = nav.IPCS().step(d.Constant(1e-2), {0: u}, p, bcs, [], d.Constant(1.0), d.Constant(1e-2), {0: zero, 1: zero}, verbose=False, tol=1e-10)
            s = nav.last_stats()
            if k >= 3:
                ms.append(s["ms_total"])
                its.append((s["newton_its"], s["momentum_its"], s["momentum_inner_its"], s["jacobian_assemblies"]))
        rec.update(n=n, ms_per_step=float(np.mean(ms)), newton_momentum_inner_assemblies=[float(x) for x in np.mean(np.array(its), axis=0)],
                   last=dict((k, s[k]) for k in ("ms_tentative", "ms_pressure", "ms_correction", "ms_assembly_J", "ms_momentum_solve", "newton_residual")))
        print(json.dumps(rec), flush=True)
    nav.reset_options()


if __name__ == "__main__":
    main()
