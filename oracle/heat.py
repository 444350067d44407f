"""Oracle restatement of flow.heat.Heat (TEST INFRASTRUCTURE).

Follows /root/reference/flow/heat.py: vertex-quadrature mass (:39-45), convection-diffusion
operator (:54-58, :88), eval_alpha_M_beta_F (:92-101), solve_alpha_M_beta_F (:103-122, sparse LU)
and the implicit-Euler wrapper of the third-party `parabolic` package used at
/root/reference/tests/test_boussinesq.py:219-229 [EXT]:  u1 = solve(1, -dt, eval(1, 0, u0, t), t + dt).
"""
import numpy as np

from . import fem, forms, solvers


class Heat:
    def __init__(self, mesh, degree, conv, kappa, rho, cp, bc, source_load=None):
        self.V = fem.Space(mesh, degree, 1)
        self.W = fem.Space(mesh, 2, mesh.dim)
        self.bc = bc  # (dofs, vals) or None
        self.M = forms.lumped_vertex_mass(self.V)
        self.A = forms.heat_operator(self.V, self.W, conv, kappa, rho * cp)
        self.b = np.zeros(self.V.nnodes) if source_load is None else source_load

    def eval_alpha_M_beta_F(self, alpha, beta, u, t=None):
        return alpha * (self.M @ u) + beta * (self.A @ u + self.b)

    def solve_alpha_M_beta_F(self, alpha, beta, b, t=None):
        A = alpha * self.M + beta * self.A
        if self.bc is not None and len(self.bc[0]):
            A, b = forms.apply_bc_rows(A, b, self.bc[0], self.bc[1])
        return solvers.lu_solve(A, b)


def implicit_euler_step(heat, u0, t, dt):
    return heat.solve_alpha_M_beta_F(1.0, -dt, heat.eval_alpha_M_beta_F(1.0, 0.0, u0, t), t + dt)
