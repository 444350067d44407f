#!/usr/bin/env python
"""BASELINE.json configs 2-4 at their stated sizes through the public facade (one JSON line per config).

  python tools/run_configs.py [cavity2d] [karman] [boussinesq] [--steps N]

These are the drivers of the reference's tests (tests/test_sealed_box.py, tests/test_karman_vortex_street.py,
tests/test_boussinesq.py) on synthetic structured meshes: every step goes through
flow_b200.navier_stokes / flow_b200.heat / flow_b200.stokes with host buffers, i.e. the numbers are end-to-end
steps/s of the facade, not kernel times.  Results of round 1 are kept in profiles/r1_configs.jsonl.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RHO_WATER, MU_WATER, CP_WATER, KAPPA_WATER = 998.21, 1.002e-3, 4184.0, 0.598


def cavity2d(steps):
    """config 2: unit square n = 333 (1 001 334 dofs), lid u = (1, 0), Re = 100, dt = 1e-2, IPCS; then the sealed-box
    invariant of tests/test_sealed_box.py:134-141 on the same mesh."""
    import flow_b200
    from flow_b200 import dolfin as d, navier_stokes as nav

    n = 333
    mesh = d.UnitSquareMesh(n, n)
    W, P = d.VectorFunctionSpace(mesh, "CG", 2), d.FunctionSpace(mesh, "CG", 1)
    bcs = [d.DirichletBC(W, (0.0, 0.0), "on_boundary"), d.DirichletBC(W, (1.0, 0.0), lambda x, on: x[1] > 1 - 1e-12)]
    u, p = d.Function(W), d.Function(P)
    zero = d.Constant((0.0, 0.0))
    st = nav.IPCS()
    hist, t0 = [], None
    for k in range(steps + 3):
        if k == 3:
            t0 = time.perf_counter()
        u, p = st.step(d.Constant(1e-2), {0: u}, p, bcs, [], d.Constant(1.0), d.Constant(1e-2), {0: zero, 1: zero}, verbose=False)
        hist.append(nav.last_stats())
    sec = (time.perf_counter() - t0) / steps
    # sealed box: f = (0, g) balanced by p0 = g y keeps u == 0
    g = -9.81
    u0, p0 = d.Function(W), d.interpolate(d.Expression("g*x[1]", degree=1, g=g), P)
    nos = [d.DirichletBC(W, (0.0, 0.0), "on_boundary")]
    for _ in range(2):
        u0, p0 = st.step(d.Constant(1e-2), {0: u0}, p0, nos, [], d.Constant(RHO_WATER), d.Constant(MU_WATER),
                         {0: d.Constant((0.0, g)), 1: d.Constant((0.0, g))}, verbose=False)
    umax = float(np.sqrt((u0.nodal() ** 2).sum(axis=1)).max())
    h = hist[3:]
    return {"config": "2: 2D lid-driven cavity, UnitSquareMesh(333), IPCS, Re=100, dt=1e-2", "dofs": W.dim() + P.dim(),
            "steps": steps, "steps_per_s_e2e": 1.0 / sec, "ms_per_step_device": float(np.mean([s["ms_total"] for s in h])),
            "newton": float(np.mean([s["newton_its"] for s in h])), "momentum_its": float(np.mean([s["momentum_its"] for s in h])),
            "pressure_its": float(np.mean([s["pressure_its"] for s in h])), "correction_its": float(np.mean([s["correction_its"] for s in h])),
            "sealed_box_umax": umax, "sealed_box_ok": umax < 1e-13}


def karman(steps):
    """config 3: channel [0,0.6]x[-0.07,0.07] with cylinder c=(0.1,0.01), r=0.02 on the structured holed-rectangle mesh,
    Re = 100 (U = 0.005, rho = 998.21, mu = 0.002), Stokes initial state, Rotational, dt = 1e-3 fixed."""
    import flow_b200
    from flow_b200 import dolfin as d, navier_stokes as nav

    x0, x1, y0, y1 = 0.0, 0.6, -0.07, 0.07
    mesh = d.RectangleWithHoleMesh(d.Point(x0, y0), d.Point(x1, y1), 480, 112, (0.1, 0.01), 0.02)
    eps = 1e-10
    left = lambda x, on: on and x[0] < x0 + eps  # noqa: E731
    right = lambda x, on: on and x[0] > x1 - eps  # noqa: E731
    walls = lambda x, on: on and (x[1] < y0 + eps or x[1] > y1 - eps)  # noqa: E731
    obstacle = lambda x, on: on and x0 + eps < x[0] < x1 - eps and y0 + eps < x[1] < y1 - eps  # noqa: E731
    WP = d.FunctionSpace(mesh, d.VectorElement("Lagrange", mesh.ufl_cell(), 2) * d.FiniteElement("Lagrange", mesh.ufl_cell(), 1))
    U, mu, rho = 0.005, 0.002, RHO_WATER
    prof = d.Expression("%e * (%e - x[1]) * (x[1] - %e) / %e" % (U, y1, y0, (0.5 * (y1 - y0)) ** 2), degree=2)

    def vbcs(V):
        return [d.DirichletBC(V, (0.0, 0.0), walls), d.DirichletBC(V, (0.0, 0.0), obstacle),
                d.DirichletBC(V.sub(0), prof, left), d.DirichletBC(V.sub(0), prof, right)]

    t0 = time.perf_counter()
    u, p = flow_b200.stokes.solve(WP, vbcs(WP.sub(0)), mu, f=d.Constant((0.0, 0.0)), verbose=False, tol=1e-10, max_iter=5000)
    t_stokes = time.perf_counter() - t0
    W, P = u.function_space(), p.function_space()
    ubcs, pbcs = vbcs(W), [d.DirichletBC(P, 0.0, right)]
    st = nav.Rotational()
    zero = d.Constant((0.0, 0.0))
    hist = []
    for k in range(steps + 3):
        if k == 3:
            t0 = time.perf_counter()
        u, p = st.step(d.Constant(1e-3), {0: u}, p, ubcs, pbcs, d.Constant(rho), d.Constant(mu), {0: zero, 1: zero}, verbose=False)
        hist.append(nav.last_stats())
    sec = (time.perf_counter() - t0) / steps
    h = hist[3:]
    ok = bool(np.isfinite(u.vector().get_local()).all() and np.isfinite(p.vector().get_local()).all())
    return {"config": "3: Karman channel with cylinder, holed-rectangle mesh 480x112, Rotational, Re=100, dt=1e-3", "dofs": W.dim() + P.dim(),
            "steps": steps, "steps_per_s_e2e": 1.0 / sec, "ms_per_step_device": float(np.mean([s["ms_total"] for s in h])),
            "stokes_bootstrap_s": t_stokes, "newton": float(np.mean([s["newton_its"] for s in h])),
            "momentum_its": float(np.mean([s["momentum_its"] for s in h])), "pressure_its": float(np.mean([s["pressure_its"] for s in h])),
            "umax": float(np.sqrt((u.nodal() ** 2).sum(axis=1)).max()), "finite": ok}


def boussinesq(steps):
    """config 4: unit square n = 667 (4 010 674 flow dofs + 1 782 225 temperature dofs), hot wall x = 0 (320 K), cold wall
    x = 1 (293 K), heat implicit Euler (new Heat object = full re-assembly every step, as the reference driver does) +
    Rotational with the Boussinesq force rho(theta) g, rho(theta) = rho0 (1 - beta (theta - theta0)), beta = 2.07e-4 1/K."""
    from flow_b200 import dolfin as d, heat, navier_stokes as nav

    n = 667
    mesh = d.UnitSquareMesh(n, n)
    W, P, Q = d.VectorFunctionSpace(mesh, "CG", 2), d.FunctionSpace(mesh, "CG", 1), d.FunctionSpace(mesh, "CG", 2)
    beta, g, T0 = 2.07e-4, -9.81, 293.0
    rho = lambda T: RHO_WATER * (1.0 - beta * (T - T0))  # noqa: E731
    theta = d.interpolate(d.Constant(T0), Q)
    u = d.Function(W)
    p = d.interpolate(d.Expression("r*g*x[1]", degree=1, r=RHO_WATER, g=g), P)
    hot = lambda x, on: on and x[0] < 1e-12  # noqa: E731
    cold = lambda x, on: on and x[0] > 1 - 1e-12  # noqa: E731
    heat_bcs = [d.DirichletBC(Q, 320.0, hot), d.DirichletBC(Q, T0, cold)]
    ubcs = [d.DirichletBC(W, (0.0, 0.0), "on_boundary")]
    st = nav.Rotational()
    dt, t = 1e-2, 0.0
    hist, t_heat, heat_its = [], 0.0, []
    for k in range(steps + 3):
        if k == 3:
            t0 = time.perf_counter()
            t_heat = 0.0
        th0 = time.perf_counter()
        stepper = heat.ImplicitEuler(heat.Heat(Q, u, KAPPA_WATER, RHO_WATER, CP_WATER, heat_bcs, d.Constant(0.0)))
        theta1 = stepper.step(theta, t, dt)
        t_heat += time.perf_counter() - th0
        heat_its.append(stepper.problem.last_iterations)
        f = d.Function(W)
        f.nodal_view()[:, 1] = rho(theta.vector().get_local()) * g
        u, p = st.step(d.Constant(dt), {0: u}, p, ubcs, [], RHO_WATER, d.Constant(MU_WATER), {0: f, 1: f}, verbose=False)
        theta = theta1
        t += dt
        hist.append(nav.last_stats())
    sec = (time.perf_counter() - t0) / steps
    h = hist[3:]
    th = theta.vector().get_local()
    return {"config": "4: Boussinesq heated cavity, UnitSquareMesh(667), heat implicit Euler + Rotational, dt=1e-2",
            "dofs": W.dim() + P.dim() + Q.dim(), "steps": steps, "steps_per_s_e2e": 1.0 / sec,
            "ms_per_step_ns_device": float(np.mean([s["ms_total"] for s in h])), "heat_s_per_step_e2e": t_heat / steps, "heat_solver_its": float(np.mean(heat_its[3:])),
            "newton": float(np.mean([s["newton_its"] for s in h])), "momentum_its": float(np.mean([s["momentum_its"] for s in h])),
            "pressure_its": float(np.mean([s["pressure_its"] for s in h])), "theta_min": float(th.min()), "theta_max": float(th.max()),
            "umax": float(np.sqrt((u.nodal() ** 2).sum(axis=1)).max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["cavity2d", "karman", "boussinesq"])
    ap.add_argument("--steps", type=int, default=0)
    args = ap.parse_args()
    runs = {"cavity2d": (cavity2d, 20), "karman": (karman, 1000), "boussinesq": (boussinesq, 20)}
    for name in args.which:
        fn, default_steps = runs[name]
        t0 = time.perf_counter()
        try:
            out = fn(args.steps or default_steps)
        except Exception as e:  # report and go on with the next config
            out = {"config": name, "error": "%s: %s" % (type(e).__name__, e)}
        out["wall_s"] = time.perf_counter() - t0
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
