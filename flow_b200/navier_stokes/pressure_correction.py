"""Chorin / IPCS / Rotational pressure-correction steppers with the reference's call
signature (flow/navier_stokes/pressure_correction.py:521-617).

The whole of ``_step`` (:468-518) -- Newton solve for the tentative velocity, pressure
Poisson solve, velocity correction -- is ONE call into the C ABI (``fb_ns_step``); all
intermediate state stays on the GPU.  Operators that the reference re-assembles on
every call (P1 stiffness, P2 mass, sparsity patterns, scatter maps) are built once per
(W, P) pair and cached on the function space.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .._lib import lib
from ..dolfin import Constant, Expression, Function, collect_bcs
from .. import hostfem
from ..message import Message

_SCHEMES = {"forward euler": _lib.FORWARD_EULER, "backward euler": _lib.BACKWARD_EULER,
            "crank-nicolson": _lib.CRANK_NICOLSON}

_last_stats = {}
_options = {}
_PRECONDS = {"jacobi": _lib.JACOBI, "amg": _lib.AMG}
_SOLVERS = {"bicgstab": _lib.BICGSTAB, "fgmres": _lib.GMRES, "gmres": _lib.GMRES}


def set_options(**kw):
    """Solver options of the engines created from now on (fields of ``fb_ns_opts``, include/flowb200.h):
    pressure_precond ("amg" | "jacobi"), jacobian_reuse, jacobian_across_steps, warm_start, jacobian_fp32,
    adaptive_forcing, momentum_maxit, pressure_maxit, correction_maxit, check_every.  The reference's counterpart
    is its hard-wired ``solver_parameters`` dicts (pressure_correction.py:228-253, :328-339, :453-464)."""
    known = {name for name, _ in _lib.NSOpts._fields_}
    for k, v in kw.items():
        if k not in known or k == "reserved":
            raise KeyError("unknown option %r" % (k,))
        if k == "pressure_precond" and isinstance(v, str):
            v = _PRECONDS[v]
        if k == "momentum_solver" and isinstance(v, str):
            v = _SOLVERS[v]
        _options[k] = v


def reset_options():
    _options.clear()


def last_stats():
    """Iteration counts and CUDA-event timings of the most recent step (dict)."""
    return dict(_last_stats)


def _scalar(c):
    return float(c.values()[0]) if isinstance(c, Constant) else float(c)


def _engine(W, P, opts=None):
    """fb_ns handle for the (W, P) pair; created on first use and kept on W's node space."""
    key = ("ns", id(P.nodes), tuple(sorted(_options.items())))
    cache = W.nodes.__dict__.setdefault("_engines", {})
    if key not in cache:
        if opts is None and _options:
            opts = _lib.NSOpts()
            lib.fb_ns_opts_default(C.byref(opts))
            for k, v in _options.items():
                setattr(opts, k, v)
        h = _lib.vp()
        _lib.check(lib.fb_ns_create(W.handle(), P.handle(), C.byref(opts) if opts is not None else None, C.byref(h)),
                   W.mesh().ctx, "fb_ns_create")
        cache[key] = h
        _replicate_pressure_amg(h, W, P)
    return cache[key]


def _replicate_pressure_amg(ns, W, P):
    """Partitioned runs: hand the engine the GLOBAL P1 stiffness matrix so that the AMG preconditioner of the
    pressure solve is the single-GPU hierarchy, replicated on every rank (fb_ns_set_pressure_amg_global).  The
    matrix is assembled redundantly from the global mesh every rank already holds; collective, set-up only."""
    mesh = W.mesh()
    if getattr(mesh, "partition", None) is None or getattr(mesh, "global_mesh", None) is None:
        return
    if not lib.fb_comm_uses_peer_memory(mesh.ctx):
        return
    from ..dolfin import FunctionSpace

    Pg = FunctionSpace(mesh.global_mesh, "CG", 1)
    A = _lib.vp()
    _lib.check(lib.fb_assemble_stiffness(Pg.handle(), C.byref(A)), mesh.ctx, "fb_assemble_stiffness(global P1)")
    plan = P.nodes.plan
    l2g = np.ascontiguousarray(plan.l2g[:plan.n_owned], dtype=np.int64)
    try:
        _lib.check(lib.fb_ns_set_pressure_amg_global(ns, Pg.handle(), A, _lib.as_pi64(l2g), l2g.size), mesh.ctx,
                   "fb_ns_set_pressure_amg_global")
    finally:
        lib.fb_mat_destroy(A)


def _forcing(f, W, theta):
    """Translate f = {0: f_n, 1: f_{n+1}} into (mode, f0, f1) host arrays for the C ABI."""
    def one(fi):
        if fi is None:
            return None, _lib.F_NONE
        if isinstance(fi, Constant):
            return _lib.f64(fi.values()), _lib.F_CONSTANT
        if isinstance(fi, Function):
            return _lib.f64(fi._vec), _lib.F_NODAL
        if isinstance(fi, Expression) or callable(fi):
            mesh, ns = W.mesh(), W.nodes
            deg = fi.degree() if isinstance(fi, Expression) else 2
            return hostfem.load_vector(mesh.coordinates(), mesh.cells(), ns.cell_nodes, ns.nnodes, 2, W.ncomp, fi,
                                       deg), _lib.F_LOAD
        if hasattr(fi, "load_vector"):
            return _lib.f64(fi.load_vector(W)), _lib.F_LOAD
        raise TypeError("unsupported forcing term %r" % (fi,))

    need0, need1 = theta != 1.0, theta != 0.0
    a0, m0 = one(f.get(0)) if need0 else (None, _lib.F_NONE)
    a1, m1 = one(f.get(1)) if need1 else (None, _lib.F_NONE)
    modes = {m for m in (m0, m1) if m != _lib.F_NONE}
    if not modes:
        return _lib.F_NONE, None, None
    if len(modes) > 1:
        raise TypeError("f[0] and f[1] must be the same kind of object")
    return modes.pop(), a0, a1


def _step(dt, u, p0, u_bcs, p_bcs, rho, mu, time_step_method, f, rotational_form=False, chorin=False, verbose=True,
          tol=1.0e-10):
    dtv, rhov, muv = _scalar(dt), _scalar(rho), _scalar(mu)
    # pressure_correction.py:488-489
    assert dtv > 0.0
    assert muv > 0.0
    u0 = u[0]
    W, P = u0.function_space(), p0.function_space()
    scheme = _SCHEMES[time_step_method]
    theta = {0: 0.0, 1: 1.0, 2: 0.5}[scheme]
    ns = _engine(W, P)
    mode, f0, f1 = _forcing(f, W, theta)
    ud, uv = collect_bcs(u_bcs, W)
    pd_, pv = collect_bcs(p_bcs, P)
    ctx = W.mesh().ctx
    u1 = Function(W, _lib.state_array(ctx, W.dim(), zero=False))  # fb_ns_step writes every entry
    p1 = Function(P, _lib.state_array(ctx, P.dim(), zero=False))
    flags = (_lib.ROTATIONAL if rotational_form else 0) | (_lib.CHORIN if chorin else 0)
    stats = _lib.NSStats()

    def ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    with Message("Computing tentative velocity, pressure, velocity correction", verbose):
        st = lib.fb_ns_step(ns, dtv, rhov, muv, scheme, flags, ptr(u0._vec), ptr(p0._vec), mode, ptr(f0), ptr(f1),
                            ud.size, _lib.as_pi64(ud), _lib.as_pd(uv), pd_.size, _lib.as_pi64(pd_), _lib.as_pd(pv),
                            float(tol), ptr(u1._vec), ptr(p1._vec), C.byref(stats))
    _lib.check(st, W.mesh().ctx, "navier_stokes step")
    _last_stats.clear()
    _last_stats.update(stats.as_dict())
    if verbose:
        print("    Newton its %d, Krylov its: momentum %d, pressure %d, correction %d; %.3f ms"
              % (stats.newton_its, stats.momentum_its, stats.pressure_its, stats.correction_its, stats.ms_total))
    return u1, p1


class Chorin(object):
    order = {"velocity": 1.0, "pressure": 0.5}

    def __init__(self):
        return

    # p0 only provides the function space (pressure_correction.py:543-552)
    def step(self, dt, u, p0, u_bcs, p_bcs, rho, mu, f, verbose=True, tol=1.0e-10):
        return _step(dt, u, p0, u_bcs, p_bcs, rho, mu, "backward euler", f, chorin=True, verbose=verbose, tol=tol)


class IPCS(object):
    order = {"velocity": 2.0, "pressure": 1.0}

    def __init__(self, time_step_method="backward euler"):
        self.time_step_method = time_step_method

    def step(self, dt, u, p0, u_bcs, p_bcs, rho, mu, f, verbose=True, tol=1.0e-10):
        return _step(dt, u, p0, u_bcs, p_bcs, rho, mu, self.time_step_method, f, verbose=verbose, tol=tol)


class Rotational(object):
    order = {"velocity": 2.0, "pressure": 1.5}

    def __init__(self, time_step_method="backward euler"):
        self.time_step_method = time_step_method

    def step(self, dt, u, p0, u_bcs, p_bcs, rho, mu, f, verbose=True, tol=1.0e-10):
        return _step(dt, u, p0, u_bcs, p_bcs, rho, mu, self.time_step_method, f, rotational_form=True, verbose=verbose,
                     tol=tol)
