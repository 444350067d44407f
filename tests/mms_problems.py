"""Manufactured-solution problems of the reference's own tests, as numpy callables.

Problem definitions follow /root/reference/tests/test_navier_stokes.py:34-229
(`_get_navier_stokes_rhs`, problem_flat / guermond1 / guermond2 / taylor) and
/root/reference/tests/test_stokes.py:20-99.  Each returns
``dict(mesh=(kind, args), u(t), p(t), f(t), mu, rho, degrees)`` where u/p/f map an
(npts, 2) coordinate array to values.
"""
import numpy as np
import sympy

MAX_DEGREE = 5


def _lambdify(exprs, extra=()):
    x0, x1 = sympy.symbols("x0 x1")
    X = sympy.DeferredVector("x")
    syms = (x0, x1) + tuple(extra)
    fs = [sympy.lambdify(syms, sympy.sympify(e).subs({X[0]: x0, X[1]: x1}), "numpy") for e in exprs]

    def call(pts, *params):
        pts = np.asarray(pts)
        cols = [np.broadcast_to(np.asarray(f(pts[:, 0], pts[:, 1], *params), dtype=float), pts.shape[:1]) for f in fs]
        return np.stack(cols, axis=1) if len(cols) > 1 else cols[0].copy()

    return call


def navier_stokes_rhs(u, p):
    """tests/test_navier_stokes.py:34-75."""
    x = sympy.DeferredVector("x")
    t, mu, rho = sympy.symbols("t, mu, rho")
    d = sympy.simplify(sympy.diff(u[0], x[0]) + sympy.diff(u[1], x[1]))
    assert d == 0
    f = []
    for i in range(2):
        f.append(
            sympy.simplify(
                rho * (sympy.diff(u[i], t) + u[0] * sympy.diff(u[i], x[0]) + u[1] * sympy.diff(u[i], x[1]))
                + sympy.diff(p, x[i])
                - mu * (sympy.diff(u[i], x[0], 2) + sympy.diff(u[i], x[1], 2))
            )
        )
    return tuple(f)


def _ns_problem(u, p, mesh, mu=1.0, rho=1.0, udeg=MAX_DEGREE, pdeg=MAX_DEGREE):
    t, smu, srho = sympy.symbols("t, mu, rho")
    f = navier_stokes_rhs(u, p)
    uf = _lambdify(u, (t,))
    pf = _lambdify([p], (t,))
    ff = _lambdify(f, (t, smu, srho))
    return dict(
        mesh=mesh,
        u=lambda tt: (lambda X: uf(X, tt)),
        p=lambda tt: (lambda X: pf(X, tt)),
        f=lambda tt: (lambda X: ff(X, tt, mu, rho)),
        mu=mu,
        rho=rho,
        udeg=udeg,
        pdeg=pdeg,
        fdeg=MAX_DEGREE,
    )


def problem_flat():
    x = sympy.DeferredVector("x")
    u = (0.0 * x[0], 0.0 * x[1])
    p = -x[1]
    return _ns_problem(u, p, ("unit_square", "left/right"), udeg=1, pdeg=1)


def problem_guermond1():
    x = sympy.DeferredVector("x")
    t = sympy.symbols("t")
    pi = sympy.pi
    m = sympy.sin(t)
    u = (
        +pi * m * 2 * sympy.sin(pi * x[1]) * sympy.cos(pi * x[1]) * sympy.sin(pi * x[0]) ** 2,
        -pi * m * 2 * sympy.sin(pi * x[0]) * sympy.cos(pi * x[0]) * sympy.sin(pi * x[1]) ** 2,
    )
    p = m * sympy.cos(pi * x[0]) * sympy.sin(pi * x[1])
    return _ns_problem(u, p, ("rectangle", (-1.0, -1.0), (1.0, 1.0), "crossed"))


def problem_guermond2():
    x = sympy.DeferredVector("x")
    t = sympy.symbols("t")
    u = (sympy.sin(x[0] + t) * sympy.sin(x[1] + t), sympy.cos(x[0] + t) * sympy.cos(x[1] + t))
    p = sympy.sin(x[0] - x[1] + t)
    return _ns_problem(u, p, ("unit_square", "crossed"))


def stokes_guermond1():
    """tests/test_stokes.py:69-99."""
    from sympy import cos, pi, sin

    x = sympy.DeferredVector("x")
    u = (
        +pi * 2 * sin(pi * x[1]) * cos(pi * x[1]) * sin(pi * x[0]) ** 2,
        -pi * 2 * sin(pi * x[0]) * cos(pi * x[0]) * sin(pi * x[1]) ** 2,
    )
    p = cos(pi * x[0]) * sin(pi * x[1])
    mu = 1.0
    f = []
    for i in range(2):
        f.append(sympy.simplify(-mu * (sympy.diff(u[i], x[0], 2) + sympy.diff(u[i], x[1], 2)) + sympy.diff(p, x[i])))
    return dict(mesh=("unit_square", "left/right"), u=_lambdify(u), p=_lambdify([p]), f=_lambdify(f), mu=mu)


def compute_numerical_order_of_convergence(Dt, errors):
    """tests/helpers.py:10-14."""
    return np.array([np.log(errors[k] / errors[k + 1]) / np.log(Dt[k] / Dt[k + 1]) for k in range(len(Dt) - 1)])
