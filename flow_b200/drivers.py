"""Driver-side pieces of the reference's time loops, evaluated on the device.

The reference's drivers (its tests are its only drivers) do per-step host work around `stepper.step`:
the CFL-like step-size control of tests/test_karman_vortex_street.py:261-287 projects the velocity magnitude onto P2
and takes its max norm; tests/test_sealed_box.py:134-141 asserts on the same quantity.  Here that projection (load
vector by quadrature, mass solve, max reduction) is one C-ABI call, `fb_ns_velocity_magnitude`; the controller
arithmetic itself is five scalar operations and stays in Python, line for line.
"""
import ctypes as C

import numpy as np

from . import _lib, hostfem
from ._lib import lib
from .dolfin import Function


def velocity_magnitude(u, P, quadrature_degree=4, want_function=False, rtol=1.0e-12):
    """(||unorm||_inf, max_i |u(x_i)|[, unorm]) with
        unorm = project(sqrt(ux**2 + uy**2 [+ uz**2]), FunctionSpace(mesh, 'Lagrange', 2),
                        form_compiler_parameters={'quadrature_degree': quadrature_degree})
    (tests/test_karman_vortex_street.py:262-268).  `P` is the pressure space of the stepper that owns the device
    operators (the engine is shared with `IPCS().step`).  Partitioned runs: maxima over the rank's owned nodes."""
    from .navier_stokes.pressure_correction import _engine

    W = u.function_space()
    ns = _engine(W, P)
    lam, w = hostfem.quadrature(W.mesh().dim, quadrature_degree)
    lam, w = _lib.f64(lam), _lib.f64(w)
    linf, nodal = C.c_double(), C.c_double()
    out = None
    if want_function:
        from .dolfin import FunctionSpace

        Q = FunctionSpace(W.mesh(), "CG", 2)
        out = Function(Q)
    _lib.check(lib.fb_ns_velocity_magnitude(ns, u._vec.ctypes.data_as(C.c_void_p), 0, w.size, _lib.as_pd(lam), _lib.as_pd(w),
                                            float(rtol), out._vec.ctypes.data_as(C.c_void_p) if out is not None else None,
                                            C.byref(linf), C.byref(nodal)), W.mesh().ctx, "fb_ns_velocity_magnitude")
    return (linf.value, nodal.value, out) if want_function else (linf.value, nodal.value)


def adapt_step_size(dt, unorm_linf, hmax, dt_max, alpha=0.5):
    """Step-size controller of tests/test_karman_vortex_street.py:270-284: approach target_dt = hmax / ||u||_inf by the
    aggressiveness factor alpha, at most doubling the step."""
    target_dt = 1.0 * hmax / unorm_linf
    return min(dt_max, dt * min(2.0, 1.0 + alpha * (target_dt - dt) / dt))
