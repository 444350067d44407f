#!/usr/bin/env python
"""Small workload for compute-sanitizer: assembly kernels (constant operators, residual, Jacobian one-pass and two-pass,
facets), the tile SpMM with its TMA/mbarrier pipeline (FB_TILE_MIN_ROWS=0), Chebyshev, FGMRES, AMG-free pressure CG,
correction CG, heat operator, Stokes -- two IPCS steps on a 4^3 cube and a 6x5 square."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FB_TILE_MIN_ROWS", "0")

from flow_b200 import dolfin as d, heat, navier_stokes as nav  # noqa: E402


def run(mesh, dim, opts):
    nav.reset_options()
    nav.set_options(**opts)
    W = d.VectorFunctionSpace(mesh, "CG", 2)
    P = d.FunctionSpace(mesh, "CG", 1)
    zero = (0.0,) * dim
    lid = tuple([1.0] + [0.0] * (dim - 1))
    bcs = [d.DirichletBC(W, zero, "on_boundary"), d.DirichletBC(W, lid, lambda x, on: x[dim - 1] > 1 - 1e-12)]
    f = d.Constant(tuple([0.0] * (dim - 1) + [-1.0]))
    u, p = d.Function(W), d.Function(P)
    for _ in range(2):
        u, p = nav.IPCS().step(d.Constant(1e-2), {0: u}, p, bcs, [], d.Constant(1.0), d.Constant(1e-2), {0: f, 1: f}, verbose=False)
    return u, p


for opts in ({}, {"deterministic_assembly": 1}, {"semi_implicit": 1}, {"inner_chebyshev": 0}):
    run(d.UnitCubeMesh(4, 4, 4), 3, opts)
    run(d.UnitSquareMesh(6, 5, "crossed"), 2, opts)
m = d.UnitSquareMesh(6, 5)
Q = d.FunctionSpace(m, "CG", 2)
Wm = d.VectorFunctionSpace(m, "CG", 2)
conv = d.Function(Wm)
conv.nodal_view()[:, 0] = 0.1
h = heat.ImplicitEuler(heat.Heat(Q, conv, 0.6, 1000.0, 4.0, [d.DirichletBC(Q, 300.0, "on_boundary")], d.Constant(0.0)))
th = d.interpolate(d.Constant(293.0), Q)
th = h.step(th, 0.0, 1.0)
print("sanitize_case: done")
