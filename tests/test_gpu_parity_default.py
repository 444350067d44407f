"""Parity at the solver options that ship and that bench.py times (fb_ns_opts_default, no overrides):
ten consecutive IPCS steps against the oracle's Newton + LU path, velocity and pressure (modulo its mean) within the
north-star tolerance of 1e-8 relative L2 after 10 steps.

* live oracle runs at sizes it finishes in seconds: the reference's guermond2 problem on UnitSquareMesh(32, 'crossed')
  (tests/test_navier_stokes.py:168-195, 403-410) stepped 10 times, and a small 3D lid-driven cavity;
* committed at-size fixtures generated from the oracle by tests/golden/make_parity_fixtures.py: the benchmark's 3D
  cavity at n = 24 / 32 (0.37 M / 0.86 M dofs) and the 2D cavity of config 2 at n = 128 and n = 333 (1.0 M dofs).

The reference's accepted Newton iterate is the output of an exact (LU) update; it lands at |F| = 1e-12 .. 1e-14 and is
itself 2e-10 .. 6e-10 (u) / 5e-10 .. 1e-8 (p) away from the root (measured with the oracle).  The
product iterates to newton_overshoot * newton_atol = 1e-13 for that reason (include/flowb200.h).
"""
import os

import numpy as np
import pytest

import mms_problems as mp
from oracle import fem, forms, navier_stokes as ons, util

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_PARITY = 1.0e-8  # north_star: relative L2 after 10 steps


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _rel_p(a, b):
    return _rel(a - a.mean(), b - b.mean())


def test_guermond2_ten_steps_default_options(gpu_ctx):
    """Config 1: guermond2 on UnitSquareMesh(32, 32, 'crossed'), rho = mu = 1, 10 steps of dt = 0.05 with the exact
    solution as (time dependent) Dirichlet data and the manufactured forcing; solver options untouched."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    nav.reset_options()
    pr = mp.problem_guermond2()
    n, dt, steps = 32, 0.05, 10
    om = fem.Mesh(*fem.unit_square_mesh(n, n, "crossed"))
    ost = ons.IPCS(om)
    bd = ost.W.boundary_dofs()
    uo = util.project(ost.W, pr["u"](0.0), pr["udeg"])
    po = util.project(ost.P, pr["p"](0.0), pr["pdeg"])

    mesh = d.UnitSquareMesh(n, n, "crossed")
    W = d.VectorFunctionSpace(mesh, "CG", 2)
    P = d.FunctionSpace(mesh, "CG", 1)
    u = d.Function(W, uo.copy())  # identical initial state on both sides (the projections are tested elsewhere)
    p = d.Function(P, po.copy())
    stepper = nav.IPCS()
    eu = ep = None
    for k in range(steps):
        t0, t1 = k * dt, (k + 1) * dt
        g = util.interpolate(ost.W, pr["u"](t1))[bd]
        l0 = forms.expression_load_vector(ost.W, pr["f"](t0), pr["fdeg"])
        l1 = forms.expression_load_vector(ost.W, pr["f"](t1), pr["fdeg"])
        uo, po = ost.step(dt, uo, po, (bd, g), None, pr["rho"], pr["mu"], l0, l1, tol=1e-10)
        bcs = [d.DirichletBC(W, d.Expression(pr["u"](t1), degree=pr["udeg"]), "on_boundary")]
        f = {0: d.Expression(pr["f"](t0), degree=pr["fdeg"]), 1: d.Expression(pr["f"](t1), degree=pr["fdeg"])}
        u, p = stepper.step(d.Constant(dt), {0: u}, p, bcs, [], d.Constant(pr["rho"]), d.Constant(pr["mu"]), f,
                            verbose=False, tol=1e-10)
        eu, ep = _rel(u._vec, uo), _rel_p(p._vec, po)
        assert eu < 10 * TOL_PARITY and ep < 100 * TOL_PARITY, (k, eu, ep)  # no excursion on the way
    print("guermond2 n=32, 10 steps, defaults: eu = %.2e, ep = %.2e, stats %s" % (eu, ep, nav.last_stats()))
    assert eu < TOL_PARITY, eu
    assert ep < TOL_PARITY, ep


def test_cavity3d_ten_steps_default_options_live(gpu_ctx):
    """Impulsively started 3D lid-driven cavity (the benchmark's problem) on UnitCubeMesh(6): 10 steps against the
    oracle's Newton + LU at the reference's settings."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    nav.reset_options()
    n, dt, rho, mu = 6, 1e-2, 1.0, 1e-2
    om = fem.Mesh(*fem.unit_cube_mesh(n, n, n))
    ost = ons.IPCS(om)
    bd = ost.W.boundary_dofs()
    g = np.zeros((ost.W.nnodes, 3))
    g[ost.W.node_coords[:, 2] > 1 - 1e-12, 0] = 1.0
    g = g.reshape(-1)
    mesh = d.UnitCubeMesh(n, n, n)
    W = d.VectorFunctionSpace(mesh, "CG", 2)
    P = d.FunctionSpace(mesh, "CG", 1)
    bcs = [d.DirichletBC(W, (0.0, 0.0, 0.0), "on_boundary"), d.DirichletBC(W, (1.0, 0.0, 0.0), lambda x, on: x[2] > 1 - 1e-12)]
    zero = d.Constant((0.0, 0.0, 0.0))
    uo, po = np.zeros(ost.W.ndofs), np.zeros(ost.P.nnodes)
    u, p = d.Function(W), d.Function(P)
    eu = ep = None
    for k in range(10):
        uo, po = ost.step(dt, uo, po, (bd, g[bd]), None, rho, mu, None, None, tol=1e-10)
        u, p = nav.IPCS().step(d.Constant(dt), {0: u}, p, bcs, [], d.Constant(rho), d.Constant(mu), {0: zero, 1: zero},
                               verbose=False, tol=1e-10)
        eu, ep = _rel(u._vec, uo), _rel_p(p._vec, po)
        assert eu < 10 * TOL_PARITY and ep < 100 * TOL_PARITY, (k, eu, ep)
    print("cavity3d n=6, 10 steps, defaults: eu = %.2e, ep = %.2e" % (eu, ep))
    assert eu < TOL_PARITY and ep < TOL_PARITY, (eu, ep)


def _run_fixture(case):
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    path = os.path.join(GOLD, "parity_%s.npz" % case)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated (tests/golden/make_parity_fixtures.py %s)" % (path, case))
    fx = np.load(path)
    nav.reset_options()
    n, dim = int(fx["n"]), int(fx["dim"])
    dt, rho, mu, tol = float(fx["dt"]), float(fx["rho"]), float(fx["mu"]), float(fx["tol"])
    if dim == 3:
        mesh = d.UnitCubeMesh(n, n, n)
        walls, lidv, zero = (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), d.Constant((0.0, 0.0, 0.0))
    else:
        mesh = d.UnitSquareMesh(n, n, "right")
        walls, lidv, zero = (0.0, 0.0), (1.0, 0.0), d.Constant((0.0, 0.0))
    W = d.VectorFunctionSpace(mesh, "CG", 2)
    P = d.FunctionSpace(mesh, "CG", 1)
    assert W.dim() == int(fx["ndofs_u"]) and P.dim() == int(fx["ndofs_p"])
    bcs = [d.DirichletBC(W, walls, "on_boundary"), d.DirichletBC(W, lidv, lambda x, on: x[dim - 1] > 1 - 1e-12)]
    u, p = d.Function(W), d.Function(P)
    iu, ip = fx["iu"], fx["ip"]
    stored = [int(s) for s in fx["steps"]]
    res = {}
    for k in range(1, max(stored) + 1):
        u, p = nav.IPCS().step(d.Constant(dt), {0: u}, p, bcs, [], d.Constant(rho), d.Constant(mu), {0: zero, 1: zero},
                               verbose=False, tol=tol)
        if k in stored:
            pv = p._vec - p._vec.mean()
            res[k] = (_rel(u._vec[iu], fx["u_%d" % k]), _rel(pv[ip], fx["p_%d" % k]),
                      abs(np.linalg.norm(u._vec) / float(fx["unorm_%d" % k]) - 1.0),
                      abs(np.linalg.norm(pv) / float(fx["pnorm_%d" % k]) - 1.0), nav.last_stats()["newton_residual"])
    print("fixture %s (n = %d, %d dofs): step -> (eu, ep, d|u|, d|p|, |F|): %s"
          % (case, n, W.dim() + P.dim(), {k: tuple("%.1e" % x for x in v) for k, v in res.items()}))
    return res, stored


@pytest.mark.parametrize("case", ["cube24", "cube32", "cavity2d_128"])
def test_ten_steps_default_options_fixture(gpu_ctx, case):
    res, stored = _run_fixture(case)
    for k in stored:
        eu, ep, du, dp, _ = res[k]
        assert eu < 10 * TOL_PARITY and ep < 100 * TOL_PARITY, (case, k, res[k])
    eu, ep, du, dp, _ = res[stored[-1]]
    assert stored[-1] == 10
    assert eu < TOL_PARITY and ep < TOL_PARITY and du < TOL_PARITY and dp < TOL_PARITY, (case, res[stored[-1]])


def test_config2_at_size_fixture(gpu_ctx):
    """Config 2 at its stated size (UnitSquareMesh(333): 1 001 334 dofs), 3 steps against the oracle's Newton + splu."""
    res, stored = _run_fixture("cavity2d_333")
    eu, ep, du, dp, _ = res[stored[-1]]
    assert eu < TOL_PARITY and ep < 10 * TOL_PARITY and du < TOL_PARITY, res
