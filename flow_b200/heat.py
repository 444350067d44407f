"""flow.heat.Heat with the reference's interface (flow/heat.py:12-122)."""
import ctypes as C

import numpy as np

from . import _lib, hostfem
from ._lib import lib
from .dolfin import Constant, Expression, Function, collect_bcs


class Heat(object):
    """u' = F(t, u) for the convection-diffusion operator of heat.py:54-58 with the
    vertex-lumped mass matrix of heat.py:39-45."""

    def __init__(self, V, conv, kappa, rho, cp, bcs, source, supg_stabilization=False):
        if isinstance(source, (int, float)):  # the reference's `source * v * dx` accepts plain numbers
            source = Constant(float(source))
        constant_source = isinstance(source, Constant)
        if supg_stabilization:
            assert conv is not None  # heat.py:74
            if source is not None and not constant_source:
                raise NotImplementedError("SUPG with a non-constant source term")
        self.V = V
        self.bcs = bcs
        mesh, ns = V.mesh(), V.nodes
        src = None
        if source is not None and not (constant_source and float(source) == 0.0):
            if constant_source:
                c0 = float(source)
                fn, deg = (lambda X: np.full(X.shape[0], c0)), 0
            elif isinstance(source, Function):
                # P_k nodal function: integrate its interpolant (the load of a Function coefficient)
                raise NotImplementedError("Function source terms: pass an Expression or a Constant")
            else:
                fn = source
                deg = source.degree() if isinstance(source, Expression) else 2
            src = hostfem.load_vector(mesh.coordinates(), mesh.cells(), ns.cell_nodes, ns.nnodes, ns.degree, 1,
                                      lambda X: np.asarray(fn(X)).reshape(-1, 1), deg)
        source_value = float(source) if constant_source else 0.0  # enters the SUPG residual term only
        h = _lib.vp()
        Wh = conv.function_space().handle() if conv is not None else None
        cv = _lib.as_pd(_lib.f64(conv._vec)) if conv is not None else None
        _lib.check(lib.fb_heat_create_supg(V.handle(), Wh, cv, float(kappa), float(rho), float(cp),
                                           _lib.as_pd(src) if src is not None else None, 1 if supg_stabilization else 0,
                                           source_value, C.byref(h)), mesh.ctx, "Heat")
        self._h = h

    def __del__(self):
        try:
            lib.fb_heat_destroy(self._h)
        except Exception:
            pass

    # heat.py:92-101 -- returns a vector (numpy array wrapped like a GenericVector)
    def eval_alpha_M_beta_F(self, alpha, beta, u, t):
        from .dolfin import Vector
        x = _lib.f64(u.vector().a if hasattr(u, "vector") else u)
        out = np.zeros_like(x)
        _lib.check(lib.fb_heat_eval(self._h, float(alpha), float(beta), _lib.as_pd(x), _lib.as_pd(out)),
                   self.V.mesh().ctx, "Heat.eval_alpha_M_beta_F")
        return Vector(out)

    # heat.py:103-122 -- b is modified in place by the Dirichlet conditions, as in the reference
    def solve_alpha_M_beta_F(self, alpha, beta, b, t, tol=1.0e-12, max_iter=5000):
        from .dolfin import Vector
        barr = b.a if isinstance(b, Vector) else b
        assert barr.flags["C_CONTIGUOUS"] and barr.dtype == np.float64
        d, v = collect_bcs(self.bcs, self.V)
        u = Function(self.V)
        its = C.c_int()
        st = lib.fb_heat_solve(self._h, float(alpha), float(beta), _lib.as_pd(barr), d.size, _lib.as_pi64(d),
                               _lib.as_pd(v), float(tol), int(max_iter), _lib.as_pd(u._vec), C.byref(its))
        _lib.check(st, self.V.mesh().ctx, "Heat.solve_alpha_M_beta_F")
        self.last_iterations = its.value
        return u


class ImplicitEuler(object):
    """parabolic.ImplicitEuler [EXT, third-party `parabolic` package used at
    tests/test_boussinesq.py:219-229]:  u1 = solve(1, -dt, eval(1, 0, u0, t), t + dt)."""

    order = 1.0

    def __init__(self, problem):
        self.problem = problem

    def step(self, u0, t, dt, tol=1.0e-12):
        L = self.problem.eval_alpha_M_beta_F(1.0, 0.0, u0, t)
        return self.problem.solve_alpha_M_beta_F(1.0, -dt, L, t + dt, tol=tol)
