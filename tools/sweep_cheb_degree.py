#!/usr/bin/env python
"""Chebyshev degree of the momentum preconditioner on BASELINE.json config 2 (2D cavity, n = 333, diffusion number
dt nu / h^2 ~ 11: kappa(D^-1 S) in the hundreds) -- outer iterations and ms per step for each degree.
  python tools/sweep_cheb_degree.py [degrees ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from flow_b200 import dolfin as d, navier_stokes as nav

    degrees = [int(a) for a in sys.argv[1:]] or [4, 8, 12, 16]
    n = 333
    mesh = d.UnitSquareMesh(n, n)
    W, P = d.VectorFunctionSpace(mesh, "CG", 2), d.FunctionSpace(mesh, "CG", 1)
    bcs = [d.DirichletBC(W, (0.0, 0.0), "on_boundary"), d.DirichletBC(W, (1.0, 0.0), lambda x, on: x[1] > 1 - 1e-12)]
    zero = d.Constant((0.0, 0.0))
    ref = None
    for deg in degrees:
        nav.reset_options()
        nav.set_options(chebyshev_degree=deg)
        u, p = d.Function(W), d.Function(P)
        st = nav.IPCS()
        hist = []
        for k in range(8):
            u, p = st.step(d.Constant(1e-2), {0: u}, p, bcs, [], d.Constant(1.0), d.Constant(1e-2), {0: zero, 1: zero}, verbose=False)
            hist.append(nav.last_stats())
        h = hist[3:]
        un = float(np.linalg.norm(u.vector().get_local()))
        if ref is None:
            ref = u.vector().get_local().copy()
        out = {"chebyshev_degree": deg, "ms_per_step_device": float(np.mean([s["ms_total"] for s in h])),
               "ms_tentative": float(np.mean([s.get("ms_tentative", 0.0) for s in h])),
               "momentum_its": float(np.mean([s["momentum_its"] for s in h])),
               "inner_its": float(np.mean([s.get("momentum_inner_its", 0.0) for s in h])),
               "newton": float(np.mean([s["newton_its"] for s in h])), "u_l2": un,
               "rel_diff_to_first": float(np.linalg.norm(u.vector().get_local() - ref) / np.linalg.norm(ref))}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
