"""Oracle restatement of flow.heat.Heat (TEST INFRASTRUCTURE).

Follows /root/reference/flow/heat.py: vertex-quadrature mass (:39-45), convection-diffusion
operator (:54-58, :88), eval_alpha_M_beta_F (:92-101), solve_alpha_M_beta_F (:103-122, sparse LU)
and the implicit-Euler wrapper of the third-party `parabolic` package used at
/root/reference/tests/test_boussinesq.py:219-229 [EXT]:  u1 = solve(1, -dt, eval(1, 0, u0, t), t + dt).
"""
import numpy as np

from . import fem, forms, solvers


def supg_tau(mesh, W, conv, eps, p):
    """stabilization.py:50-143 evaluated at the three vertices of every triangle: (nc, 3).
    h = directed element diameter (:74-111), Pe = |b| h / (2 p eps), tau = h^2/(4 eps p) xi(Pe)."""
    assert mesh.dim == 2
    X = mesh.points[mesh.cells]                      # (nc, 3, 2)
    V = conv.reshape(-1, 2)[W.cell_nodes[:, :3]]     # convection at the vertices (P2 vertex dofs)
    nrm = np.sqrt((V ** 2).sum(axis=2))
    s = np.zeros_like(nrm)
    for i in range(3):
        for j in range(i + 1, 3):
            e = X[:, i, :] - X[:, j, :]
            s += np.abs(e[:, None, 1] * V[:, :, 0] - e[:, None, 0] * V[:, :, 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        h = 4.0 * nrm * mesh.vol[:, None] / s
        Pe = 0.5 * nrm * h / (p * eps)
        xi = np.where(Pe > 1.0e-5, (1.0 / np.tanh(Pe) - 1.0 / Pe) / Pe, 1.0 / 3.0 - Pe ** 2 / 45.0 + 2.0 / 945.0 * Pe ** 4)
        tau = h * h / 4.0 / eps / p * xi
    tau = np.where(nrm < 1.0e-10, 0.0, tau)
    if (tau > 1.0e3).any():
        raise RuntimeError("SUPG tau > 1e3 (stabilization.py:132-140)")
    return tau


def supg_terms(V, W, conv, kappa, rho_cp, source, p):
    """heat.py:79-86: Msupg = int u tau conv.grad v;  Asupg = int [kappa Lap u / rho_cp - conv.grad u] tau conv.grad v;
    bsupg = int source/rho_cp tau conv.grad v  (tau = P1 interpolant of the vertex values, degree-7 quadrature)."""
    import scipy.sparse as sp

    m = V.mesh
    lam, w = fem.simplex_quadrature(2, 7)
    phi, dphi = V.tabulate(lam)
    g = forms.phys_grads(m, dphi)                                  # (c,q,a,k)
    wphi, _ = W.tabulate(lam)
    cq = np.einsum("qa,cai->cqi", wphi, conv.reshape(-1, 2)[W.cell_nodes])
    tau = np.einsum("qv,cv->cq", lam, supg_tau(m, W, conv, kappa, p))
    test = np.einsum("cq,cqk,cqak->cqa", tau, cq, g)               # tau conv.grad phi_a
    if V.degree == 2:
        H = fem.p2_second_derivs(2)
        lap = np.einsum("bmn,cmk,cnk->cb", H, m.glam, m.glam)      # Laplacian of phi_b, constant per cell
    else:
        lap = np.zeros((m.nc, V.nl))
    Me = np.einsum("q,qb,cqa,c->cab", w, phi, test, m.vol)
    cgb = np.einsum("cqk,cqbk->cqb", cq, g)
    Ae = np.einsum("q,cqb,cqa,c->cab", w, kappa * lap[:, None, :] / rho_cp - cgb, test, m.vol)
    be = source / rho_cp * np.einsum("q,cqa,c->ca", w, test, m.vol)
    Ms = forms._scatter_matrix(V.cell_nodes, V.cell_nodes, Me, V.nnodes, V.nnodes)
    As = forms._scatter_matrix(V.cell_nodes, V.cell_nodes, Ae, V.nnodes, V.nnodes)
    bs = np.zeros(V.nnodes)
    np.add.at(bs, V.cell_nodes.ravel(), be.ravel())
    return Ms, As, bs


class Heat:
    def __init__(self, mesh, degree, conv, kappa, rho, cp, bc, source_load=None, supg=False, source=0.0):
        self.V = fem.Space(mesh, degree, 1)
        self.W = fem.Space(mesh, 2, mesh.dim)
        self.bc = bc  # (dofs, vals) or None
        self.M = forms.lumped_vertex_mass(self.V)
        self.A = forms.heat_operator(self.V, self.W, conv, kappa, rho * cp)
        self.b = np.zeros(self.V.nnodes) if source_load is None else source_load
        if supg:
            assert conv is not None  # heat.py:74
            Ms, As, bs = supg_terms(self.V, self.W, conv, kappa, rho * cp, source, degree)
            self.M = (self.M + Ms).tocsr()
            self.A = (self.A + As).tocsr()
            self.b = self.b + bs

    def eval_alpha_M_beta_F(self, alpha, beta, u, t=None):
        return alpha * (self.M @ u) + beta * (self.A @ u + self.b)

    def solve_alpha_M_beta_F(self, alpha, beta, b, t=None):
        A = alpha * self.M + beta * self.A
        if self.bc is not None and len(self.bc[0]):
            A, b = forms.apply_bc_rows(A, b, self.bc[0], self.bc[1])
        return solvers.lu_solve(A, b)


def implicit_euler_step(heat, u0, t, dt):
    return heat.solve_alpha_M_beta_F(1.0, -dt, heat.eval_alpha_M_beta_F(1.0, 0.0, u0, t), t + dt)
