"""Oracle data-entry / error-measurement helpers (TEST INFRASTRUCTURE).

Restates the DOLFIN helpers the reference's tests call: ``project``
(tests/test_navier_stokes.py:296-308), ``errornorm`` (:333, :360),
``assemble(f*dx)`` (:349-351), ``interpolate``.  Semantics [EXT] per SURVEY.md 8c.
"""
import numpy as np
import scipy.sparse as sp

from . import fem, forms, solvers


def project(space, func, degree):
    """L2 projection of Expression(func, degree) into `space` (mass solve by LU)."""
    b = forms.expression_load_vector(space, func, degree)
    scalar = fem.Space(space.mesh, space.degree, 1)
    M = forms.mass_matrix(scalar)
    if space.ncomp > 1:
        M = sp.kron(M, sp.eye(space.ncomp), format="csr")
    return solvers.lu_solve(M, b)


def interpolate(space, func):
    vals = np.asarray(func(space.node_coords))
    return vals.reshape(space.nnodes, -1).ravel() if space.ncomp > 1 else vals.reshape(-1)


def errornorm(space, func, uh, degree_rise=3):
    """errornorm(u, uh) [EXT]: interpolate both into P_{k+3}, L2 norm of the difference."""
    m = space.mesh
    k = space.degree + degree_rise
    al = fem.lattice(m.dim, k) / float(k)
    X = np.einsum("nm,cmk->cnk", al, m.points[m.cells])
    ue = np.asarray(func(X.reshape(-1, m.dim))).reshape(m.nc, al.shape[0], -1)
    phi, _ = space.tabulate(al)
    Uh = uh.reshape(space.nnodes, -1)[space.cell_nodes]  # (c,a,i)
    uhl = np.einsum("na,cai->cni", phi, Uh)
    diff = ue - uhl
    lam, w = fem.simplex_quadrature(m.dim, 2 * k)
    psi = fem.tabulate_pk(m.dim, k, lam)
    Mk = np.einsum("q,qn,qm->nm", w, psi, psi)
    e2 = np.einsum("cni,nm,cmi,c->", diff, Mk, diff, m.vol)
    return float(np.sqrt(max(e2, 0.0)))


def integrate_expression(mesh, func, degree):
    """assemble(Expression*dx(mesh)): P_k interpolant integrated exactly."""
    al = fem.lattice(mesh.dim, degree) / float(degree)
    X = np.einsum("nm,cmk->cnk", al, mesh.points[mesh.cells])
    vals = np.asarray(func(X.reshape(-1, mesh.dim))).reshape(mesh.nc, al.shape[0])
    lam, w = fem.simplex_quadrature(mesh.dim, degree)
    psi = fem.tabulate_pk(mesh.dim, degree, lam)
    return float(np.einsum("q,qn,cn,c->", w, psi, vals, mesh.vol))


def integrate_function(space, uh):
    m = space.mesh
    lam, w = fem.simplex_quadrature(m.dim, space.degree)
    phi, _ = space.tabulate(lam)
    return float(np.einsum("q,qa,ca,c->", w, phi, uh[space.cell_nodes], m.vol))


def l2_norm(space, uh):
    """norm(f, 'L2') = sqrt(int f^2)."""
    scalar = fem.Space(space.mesh, space.degree, 1)
    M = forms.mass_matrix(scalar)
    U = uh.reshape(space.nnodes, -1)
    return float(np.sqrt(sum(U[:, i] @ (M @ U[:, i]) for i in range(U.shape[1]))))
