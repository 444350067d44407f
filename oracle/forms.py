"""Oracle assembly of the hot-path forms (TEST INFRASTRUCTURE).

Each function cites the reference form it restates; every integral is taken
with a quadrature rule that is exact for the polynomial integrand on affine
simplices, so the result equals FFC's up to rounding.
"""
import numpy as np
import scipy.sparse as sp

from . import fem


def _es(*args):
    """einsum with contraction-order optimisation (BLAS where possible): same sums, ~10x faster."""
    return np.einsum(*args, optimize=True)


def _scatter_matrix(row_dofs, col_dofs, Ke, nrows, ncols):
    nlr = row_dofs.shape[1]
    nlc = col_dofs.shape[1]
    rows = np.repeat(row_dofs, nlc, axis=1).ravel()
    cols = np.tile(col_dofs, (1, nlr)).ravel()
    A = sp.csr_matrix((Ke.ravel(), (rows, cols)), shape=(nrows, ncols))
    A.sum_duplicates()
    A.sort_indices()
    return A


def vector_dofs(cell_nodes, d):
    """(nc, nl*d) interleaved dofs: local index a*d+i -> d*node_a+i."""
    return (cell_nodes[:, :, None] * d + np.arange(d)[None, None, :]).reshape(cell_nodes.shape[0], -1)


def phys_grads(mesh, dphi):
    """(nc, nq, nl, d) physical gradients from dphi/dlam (nq, nl, d+1)."""
    return _es("qam,cmk->cqak", dphi, mesh.glam)


# ---- constant scalar matrices --------------------------------------------
def mass_matrix(space):
    """inner(u, v)*dx on the node space (pressure_correction.py:442 per component)."""
    m = space.mesh
    lam, w = fem.simplex_quadrature(m.dim, 2 * space.degree)
    phi, _ = space.tabulate(lam)
    Mref = _es("q,qa,qb->ab", w, phi, phi)
    Ke = m.vol[:, None, None] * Mref[None]
    return _scatter_matrix(space.cell_nodes, space.cell_nodes, Ke, space.nnodes, space.nnodes)


def stiffness_matrix(space):
    """dot(grad(p), grad(q))*dx (pressure_correction.py:317)."""
    m = space.mesh
    lam, w = fem.simplex_quadrature(m.dim, max(0, 2 * (space.degree - 1)))
    _, dphi = space.tabulate(lam)
    g = phys_grads(m, dphi)
    Ke = _es("q,cqak,cqbk->cab", w, g, g) * m.vol[:, None, None]
    return _scatter_matrix(space.cell_nodes, space.cell_nodes, Ke, space.nnodes, space.nnodes)


def lumped_vertex_mass(space):
    """u*v*dx with quadrature_rule 'vertex' (heat.py:39-45): diagonal, zero on edge nodes."""
    m = space.mesh
    diag = np.zeros(space.nnodes)
    np.add.at(diag, m.cells.ravel(), np.repeat(m.vol / (m.dim + 1), m.dim + 1))
    return sp.diags(diag).tocsr()


def load_vector(space, fq, lam, w):
    """int f.v dx given f at quadrature points: fq (nc, nq, ncomp)."""
    m = space.mesh
    phi, _ = space.tabulate(lam)
    be = _es("q,cqi,qa,c->cai", w, fq, phi, m.vol)
    b = np.zeros(space.nnodes * space.ncomp)
    dofs = vector_dofs(space.cell_nodes, space.ncomp)
    np.add.at(b, dofs.ravel(), be.reshape(m.nc, -1).ravel())
    return b


def expression_load_vector(space, func, degree):
    """int I_k(f).v dx, I_k = per-cell P_k nodal interpolant: DOLFIN's meaning of
    Expression(..., degree=k) inside a form [EXT] (tests/test_navier_stokes.py:250-259)."""
    m = space.mesh
    lam, w = fem.simplex_quadrature(m.dim, degree + space.degree)
    if degree == 0:
        xc = m.points[m.cells].mean(axis=1)
        vals = np.atleast_2d(func(xc).T).T.reshape(m.nc, 1, -1)
        fq = np.repeat(vals, lam.shape[0], axis=1)
    else:
        al = fem.lattice(m.dim, degree) / float(degree)
        X = _es("nm,cmk->cnk", al, m.points[m.cells])  # lattice points per cell
        vals = func(X.reshape(-1, m.dim)).reshape(m.nc, al.shape[0], -1)
        psi = fem.tabulate_pk(m.dim, degree, lam)
        fq = _es("qn,cni->cqi", psi, vals)
    return load_vector(space, fq, lam, w)


# ---- momentum residual / Jacobian ------------------------------------------
def _facet_rule(dim, degree):
    """Facet quadrature expressed in the barycentrics of the parent cell, per local facet."""
    flam, fw = fem.simplex_quadrature(dim - 1, degree)
    out = []
    for f in range(dim + 1):
        lam = np.zeros((flam.shape[0], dim + 1))
        others = [i for i in range(dim + 1) if i != f]
        lam[:, others] = flam
        out.append(lam)
    return out, fw


def momentum_residual_jacobian(W, P, ui, u0, p0, load, dt, rho, mu, theta, want_J=True, semi_implicit=False):
    """F1 (pressure_correction.py:169-190) and J = derivative(F1, ui) (:202).

    theta = 0 / 1 / 0.5 for forward Euler / backward Euler / Crank-Nicolson.
    `load` is the already time-weighted load vector int f.v dx:
    f[0] (FE), f[1] (BE), 0.5*(f[0]+f[1]) (CN).
    _rhs_weak (pressure_correction.py:135-144):
       R(u;v) = (f,v) - rho/2[((grad u)u, v) - ((grad v)u, u)] - (sigma(u,p0), eps(v))
                - (p0 n, v)_ds + mu((grad u)^T n, v)_ds
    F1 = (ui-u0, v) - dt/rho * [(1-theta) R(u0) + theta R(ui)]
    """
    m = W.mesh
    d = m.dim
    nl = W.nl
    lam, w = fem.simplex_quadrature(d, 5)
    phi, dphi = W.tabulate(lam)
    g = phys_grads(m, dphi)  # (c,q,a,k)
    psi, _ = P.tabulate(lam)
    vol = m.vol
    p0q = _es("qa,ca->cq", psi, p0[P.cell_nodes])

    def cell_R(u, adv=None):
        # adv: advecting velocity of the convective term (None: u itself, the reference's form :138-139;
        # semi-implicit option: u0, i.e. ((grad u) u0, v) - ((grad v) u0, u), cf. the notes at :96-101)
        U = u.reshape(-1, d)[W.cell_nodes]  # (c,a,i)
        uq = _es("qa,cai->cqi", phi, U)
        gu = _es("cqak,cai->cqik", g, U)  # d u_i / d x_k
        wq = uq if adv is None else _es("qa,cai->cqi", phi, adv.reshape(-1, d)[W.cell_nodes])
        conv1 = _es("cqik,cqk->cqi", gu, wq)  # (grad u) w
        ugphi = _es("cqk,cqak->cqa", wq, g)  # w . grad phi_a
        eps = 0.5 * (gu + np.swapaxes(gu, 2, 3))
        R = -rho * 0.5 * (
            _es("q,cqi,qa->cai", w, conv1, phi) - _es("q,cqa,cqi->cai", w, ugphi, uq)
        )
        R += -2 * mu * _es("q,cqik,cqak->cai", w, eps, g)
        R += _es("q,cq,cqai->cai", w, p0q, g)  # + p0 div v
        return R * vol[:, None, None], (U, uq, gu, ugphi)

    Ui = ui.reshape(-1, d)[W.cell_nodes]
    U0 = u0.reshape(-1, d)[W.cell_nodes]
    Mref = _es("q,qa,qb->ab", w, phi, phi)
    Fe = _es("ab,cbi,c->cai", Mref, Ui - U0, vol)
    ctx = None
    if theta != 0.0:
        Ri, ctx = cell_R(ui, u0 if semi_implicit else None)
        Fe -= dt / rho * theta * Ri
    if theta != 1.0:
        R0, _ = cell_R(u0)
        Fe -= dt / rho * (1 - theta) * R0

    dofs = vector_dofs(W.cell_nodes, d)
    F = np.zeros(W.nnodes * d)
    np.add.at(F, dofs.ravel(), Fe.reshape(m.nc, -1).ravel())
    F -= dt / rho * load

    # exterior-facet terms: - (p0 n, v)_ds + mu ((grad u)^T n, v)_ds
    flams, fw = _facet_rule(d, 4)
    bc_ = m.bfacet_cell
    bf_ = m.bfacet_local
    if want_J:
        Je = np.zeros((m.nc, nl, d, nl, d))
    for f in range(d + 1):
        sel = bc_[bf_ == f]
        if sel.size == 0:
            continue
        fphi, fdphi = W.tabulate(flams[f])
        fpsi, _ = P.tabulate(flams[f])
        gl = m.glam[sel]
        # n * |facet| = -grad(lambda_f) * d * vol
        nA = -gl[:, f, :] * (d * vol[sel])[:, None]
        gf = _es("qam,cmk->cqak", fdphi, gl)
        p0f = _es("qa,ca->cq", fpsi, p0[P.cell_nodes[sel]])

        def facet_R(u):
            U = u.reshape(-1, d)[W.cell_nodes[sel]]
            gu = _es("cqak,cai->cqik", gf, U)
            gtn = _es("cqki,ck->cqi", gu, nA)  # (grad u)^T n
            return -_es("q,cq,ci,qa->cai", fw, p0f, nA, fphi) + mu * _es(
                "q,cqi,qa->cai", fw, gtn, fphi
            )

        Rf = np.zeros((sel.size, nl, d))
        if theta != 0.0:
            Rf += theta * facet_R(ui)
        if theta != 1.0:
            Rf += (1 - theta) * facet_R(u0)
        np.add.at(F, dofs[sel].ravel(), (-dt / rho * Rf).reshape(sel.size, -1).ravel())
        if want_J and theta != 0.0:
            # -theta dt/rho mu ((grad delta)^T n, v)_ds ; delta = phi_b e_j, v = phi_a e_i
            # ((grad delta)^T n)_i = d_i phi_b n_j
            Jf = -theta * dt / rho * mu * _es("q,qa,cqbi,cj->caibj", fw, fphi, gf, nA)
            np.add.at(Je, sel, Jf)

    if not want_J:
        return F, None

    eye = np.eye(d)
    Je += _es("ab,ij,c->caibj", Mref, eye, vol)
    if theta != 0.0:
        U, uq, gu, ugphi = ctx
        c1 = theta * dt * 0.5  # rho cancels
        # ((grad delta) ui, v) - ((grad v) ui, delta): component-diagonal, skew in (a,b)
        S = _es("q,cqb,qa->cab", w, ugphi, phi)
        S = (S - np.swapaxes(S, 1, 2)) * vol[:, None, None]
        Je += c1 * _es("cab,ij->caibj", S, eye)
        if not semi_implicit:  # the advecting velocity is u0: these two terms (its derivative) vanish
            # ((grad ui) delta, v): phi_a phi_b d_j ui_i
            Je += c1 * _es("q,qa,qb,cqij,c->caibj", w, phi, phi, gu, vol)
            # -((grad v) delta, ui): - d_j phi_a phi_b ui_i
            Je -= c1 * _es("q,cqaj,qb,cqi,c->caibj", w, g, phi, uq, vol)
        c2 = theta * dt / rho * mu
        # 2 eps(delta):eps(v) = delta_ij grad phi_a.grad phi_b + d_i phi_b d_j phi_a
        K = _es("q,cqak,cqbk,c->cab", w, g, g, vol)
        Je += c2 * _es("cab,ij->caibj", K, eye)
        Je += c2 * _es("q,cqbi,cqaj,c->caibj", w, g, g, vol)
    J = _scatter_matrix(dofs, dofs, Je.reshape(m.nc, nl * d, nl * d), W.nnodes * d, W.nnodes * d)
    return F, J


# ---- pressure Poisson and velocity correction -----------------------------
def divergence_at(W, lam, u):
    """div(u) at barycentric points: (nc, nq)."""
    m = W.mesh
    _, dphi = W.tabulate(lam)
    g = phys_grads(m, dphi)
    U = u.reshape(-1, m.dim)[W.cell_nodes]
    return _es("cqai,cai->cq", g, U)


def grad_div(W, u):
    """grad(div u): constant per cell for P2 (nc, d)."""
    m = W.mesh
    H = fem.p2_second_derivs(m.dim)  # (a,m,n)
    U = u.reshape(-1, m.dim)[W.cell_nodes]
    # d_k d_i phi_a = sum_mn H[a,m,n] glam[m,i] glam[n,k]
    return _es("amn,cmi,cnk,cai->ck", H, m.glam, m.glam, U)


def pressure_rhs(W, P, ui, p0, dt, rho, mu, rotational, alpha=1.0):
    """L2 = -alpha rho/dt div(ui) q + grad p0.grad q [- mu grad(div ui).grad q]
    (pressure_correction.py:318-323)."""
    m = W.mesh
    lam, w = fem.simplex_quadrature(m.dim, 2)
    psi, dpsi = P.tabulate(lam)
    gq = phys_grads(m, dpsi)[:, 0]  # P1 gradients constant: (c,a,k)
    div = divergence_at(W, lam, ui)
    be = -alpha * rho / dt * _es("q,cq,qa,c->ca", w, div, psi, m.vol)
    gp0 = _es("cak,ca->ck", gq, p0[P.cell_nodes])
    be += _es("ck,cak,c->ca", gp0, gq, m.vol)
    if rotational:
        be -= mu * _es("ck,cak,c->ca", grad_div(W, ui), gq, m.vol)
    b = np.zeros(P.nnodes)
    np.add.at(b, P.cell_nodes.ravel(), be.ravel())
    return b


def correction_rhs(W, P, ui, p1, p0, dt, rho, mu, rotational):
    """L3 = (ui, v) - dt/rho (grad phi, v), phi = p1 - p0 [+ mu div ui]
    (pressure_correction.py:444-449)."""
    m = W.mesh
    d = m.dim
    lam, w = fem.simplex_quadrature(d, 4)
    phi, _ = W.tabulate(lam)
    psi, dpsi = P.tabulate(lam)
    gq = phys_grads(m, dpsi)[:, 0]
    U = ui.reshape(-1, d)[W.cell_nodes]
    Mref = _es("q,qa,qb->ab", w, phi, phi)
    be = _es("ab,cbi,c->cai", Mref, U, m.vol)
    gphi = _es("cak,ca->ck", gq, (p1 - p0)[P.cell_nodes])
    if rotational:
        gphi = gphi + mu * grad_div(W, ui)
    be -= dt / rho * _es("q,qa,ck,c->cak", w, phi, gphi, m.vol)
    b = np.zeros(W.nnodes * d)
    np.add.at(b, vector_dofs(W.cell_nodes, d).ravel(), be.reshape(m.nc, -1).ravel())
    return b


# ---- heat -------------------------------------------------------------------
def heat_operator(V, W, conv, kappa, rho_cp):
    """A from lhs of  -kappa grad u.grad(v/rho_cp) - (conv.grad u) v  (heat.py:54-58, 88)."""
    m = V.mesh
    lam, w = fem.simplex_quadrature(m.dim, 2 + 2 * V.degree - 1)
    phi, dphi = V.tabulate(lam)
    g = phys_grads(m, dphi)
    Ke = -(kappa / rho_cp) * _es("q,cqak,cqbk->cab", w, g, g)
    if conv is not None:
        wphi, _ = W.tabulate(lam)
        cq = _es("qa,cai->cqi", wphi, conv.reshape(-1, m.dim)[W.cell_nodes])
        Ke -= _es("q,cqk,cqbk,qa->cab", w, cq, g, phi)
    Ke *= m.vol[:, None, None]
    return _scatter_matrix(V.cell_nodes, V.cell_nodes, Ke, V.nnodes, V.nnodes)


# ---- Stokes -------------------------------------------------------------------
def stokes_blocks(W, P, mu):
    """Blocks of a = mu grad u:grad v - p div v - q div u (stokes.py:40-42) and of the
    preconditioner form mu grad u:grad v - p q (stokes.py:55-56)."""
    m = W.mesh
    d = m.dim
    K = stiffness_matrix(fem.Space(m, 2, 1)) * mu
    lam, w = fem.simplex_quadrature(d, 2)
    _, dphi = W.tabulate(lam)
    psi, _ = P.tabulate(lam)
    g = phys_grads(m, dphi)
    # B[q_a, (b,j)] = - int psi_a d_j phi_b
    Be = -_es("q,qa,cqbj,c->cabj", w, psi, g, m.vol).reshape(m.nc, P.nl, -1)
    B = _scatter_matrix(P.cell_nodes, vector_dofs(W.cell_nodes, d), Be, P.nnodes, W.nnodes * d)
    Mp = mass_matrix(P)
    A = sp.kron(K, sp.eye(d), format="csr")
    return A, B, Mp


# ---- Dirichlet application ---------------------------------------------------
def apply_bc_rows(A, b, dofs, vals):
    """DirichletBC.apply(A, b): rows -> identity, b -> g (non-symmetric) [EXT]."""
    A = A.tocsr(copy=True)
    mask = np.zeros(A.shape[0], bool)
    mask[dofs] = True
    keep = sp.diags((~mask).astype(float))
    A = keep @ A + sp.diags(mask.astype(float))
    if b is not None:
        b = b.copy()
        b[dofs] = vals
    return A.tocsr(), b


def apply_bc_symmetric(A, b, dofs, vals):
    """assemble_system / solve(symmetric=True): zero row+column, unit diagonal, lifted RHS [EXT]."""
    n = A.shape[0]
    mask = np.zeros(n, bool)
    mask[dofs] = True
    g = np.zeros(n)
    g[dofs] = vals
    b = b - A @ g
    b[dofs] = vals
    keep = sp.diags((~mask).astype(float))
    A = keep @ A @ keep + sp.diags(mask.astype(float))
    return A.tocsr(), b
