"""flow.stokes.solve with the reference's signature (flow/stokes.py:13-148)."""
import ctypes as C

import numpy as np

from . import _lib, hostfem
from ._lib import lib
from .dolfin import Constant, Expression, Function, collect_bcs


def solve(WP, bcs, mu, f, verbose=True, tol=1.0e-13, max_iter=500):
    # stokes.py:23
    assert mu > 0.0
    W, P = WP.sub(0), WP.sub(1)
    Wc, Pc = W.collapse(), P.collapse()
    u_bcs = [bc for bc in bcs if bc.function_space().nodes is W.nodes]
    p_bcs = [bc for bc in bcs if bc.function_space().nodes is P.nodes]
    ud, uv = collect_bcs(u_bcs, Wc)
    pd_, pv = collect_bcs(p_bcs, Pc)
    mesh, ns = Wc.mesh(), Wc.nodes
    if isinstance(f, Constant):
        mode, fa = _lib.F_CONSTANT, _lib.f64(f.values())
    elif isinstance(f, Function):
        mode, fa = _lib.F_NODAL, _lib.f64(f._vec)
    else:
        deg = f.degree() if isinstance(f, Expression) else 2
        mode = _lib.F_LOAD
        fa = hostfem.load_vector(mesh.coordinates(), mesh.cells(), ns.cell_nodes, ns.nnodes, 2, Wc.ncomp, f, deg)
    u = Function(Wc)
    p = Function(Pc)
    its = C.c_int()
    st = lib.fb_stokes_solve(Wc.handle(), Pc.handle(), float(mu), mode, _lib.as_pd(fa), ud.size, _lib.as_pi64(ud),
                             _lib.as_pd(uv), pd_.size, _lib.as_pi64(pd_), _lib.as_pd(pv), float(tol), int(max_iter),
                             _lib.as_pd(u._vec), _lib.as_pd(p._vec), C.byref(its))
    _lib.check(st, mesh.ctx, "stokes.solve")
    if verbose:
        print("    Stokes GMRES iterations: %d" % its.value)
    return u, p
