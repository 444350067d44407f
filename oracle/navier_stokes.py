"""Oracle restatement of flow.navier_stokes (TEST INFRASTRUCTURE).

Follows /root/reference/flow/navier_stokes/pressure_correction.py:
``_compute_tentative_velocity`` (:147-255), ``_compute_pressure`` (:258-433),
``_compute_velocity_correction`` (:436-465), ``_step`` (:468-518) and the
Chorin / IPCS / Rotational classes (:521-617).  All state is plain numpy in
the canonical numbering of oracle/fem.py.
"""
import numpy as np

from . import fem, forms, solvers

THETA = {"forward euler": 0.0, "backward euler": 1.0, "crank-nicolson": 0.5}


class Stepper:
    """One object per (mesh, W, P); constant operators are cached (the reference
    re-assembles them every step, which changes nothing numerically)."""

    order = {"velocity": 2.0, "pressure": 1.0}

    def __init__(self, mesh, time_step_method="backward euler", rotational=False, chorin=False, linear="lu", semi_implicit=False):
        self.mesh = mesh
        # semi_implicit: (u0 . grad) ui instead of (ui . grad) ui -- the linearisation the reference's notes discuss
        # (:96-101, :204-219) but do not implement; oracle of the product's opts.semi_implicit
        self.semi_implicit = semi_implicit
        self.linear = linear  # "lu": the reference's Newton+LU; "krylov": C/OpenMP Jacobi-Krylov (CPU timing)
        self.W = fem.Space(mesh, 2, mesh.dim)
        self.P = fem.Space(mesh, 1, 1)
        self.Wn = fem.Space(mesh, 2, 1)
        self.method = time_step_method
        self.rotational = rotational
        self.chorin = chorin
        self.A_p = forms.stiffness_matrix(self.P)
        self.M_u = forms.mass_matrix(self.Wn)
        import scipy.sparse as sp

        self.M_vec = sp.kron(self.M_u, sp.eye(mesh.dim), format="csr")
        self.info = {}

    # -- pressure_correction.py:147-255
    def tentative_velocity(self, u0, p0, load0, load1, u_bc, rho, mu, dt, tol=1e-10):
        theta = THETA[self.method]
        load = (1 - theta) * (0 if load0 is None else load0) + theta * (0 if load1 is None else load1)
        if np.isscalar(load):
            load = np.zeros(self.W.ndofs)
        dofs, vals = u_bc

        def rj(x, want_J):
            return forms.momentum_residual_jacobian(
                self.W, self.P, x, u0, p0, load, dt, rho, mu, theta, want_J=want_J, semi_implicit=self.semi_implicit
            )

        if self.linear == "tight-krylov":  # the LU iterates, computed without the (infeasible) 3D LU fill
            ui, its = solvers.newton(rj, u0, dofs, vals, atol=tol, maxit=10, report=self.info,
                                     linear_solve=solvers.tight_krylov_solve)
        else:
            newton = solvers.newton if self.linear == "lu" else solvers.newton_krylov
            ui, its = newton(rj, u0, dofs, vals, atol=tol, maxit=10, report=self.info)
        self.info["newton_its"] = its
        return ui

    # -- pressure_correction.py:258-433
    def pressure(self, ui, p0, p_bc, rho, mu, dt, tol):
        b = forms.pressure_rhs(self.W, self.P, ui, p0, dt, rho, mu, self.rotational)
        if p_bc is not None and len(p_bc[0]) > 0:
            A, b = forms.apply_bc_symmetric(self.A_p, b, p_bc[0], p_bc[1])
            p1, its = solvers.pcg(A, b, tol, 100 * 50)  # reference: maxit 100 with AMG
        else:
            cg = solvers.pcg if self.linear == "lu" else solvers.c_pcg
            p1, its = cg(self.A_p, b, tol, 1000 * 50)  # reference: maxit 1000 with AMG
        self.info["pressure_its"] = its
        return p1

    # -- pressure_correction.py:436-465
    def velocity_correction(self, ui, p1, p0, u_bc, rho, mu, dt, tol):
        b = forms.correction_rhs(self.W, self.P, ui, p1, p0, dt, rho, mu, self.rotational)
        A, b = forms.apply_bc_symmetric(self.M_vec, b, u_bc[0], u_bc[1])
        cg = solvers.pcg if self.linear == "lu" else solvers.c_pcg
        u1, its = cg(A, b, tol, 100 * 50)
        self.info["correction_its"] = its
        return u1

    # -- pressure_correction.py:468-518 and the three .step methods
    def step(self, dt, u0, p0, u_bc, p_bc, rho, mu, load0=None, load1=None, tol=1e-10):
        assert dt > 0.0 and mu > 0.0  # :488-489
        if self.chorin:
            p0 = np.zeros_like(p0)  # :545
        ui = self.tentative_velocity(u0, p0, load0, load1, u_bc, rho, mu, dt, tol=1e-10)  # :499
        p1 = self.pressure(ui, p0, p_bc, rho, mu, dt, tol)
        u1 = self.velocity_correction(ui, p1, p0, u_bc, rho, mu, dt, tol)
        self.info["ui"] = ui
        return u1, p1


def Chorin(mesh):
    s = Stepper(mesh, "backward euler", rotational=False, chorin=True)
    s.order = {"velocity": 1.0, "pressure": 0.5}
    return s


def IPCS(mesh, time_step_method="backward euler", linear="lu", semi_implicit=False):
    s = Stepper(mesh, time_step_method, rotational=False, linear=linear, semi_implicit=semi_implicit)
    s.order = {"velocity": 2.0, "pressure": 1.0}
    return s


def Rotational(mesh, time_step_method="backward euler"):
    s = Stepper(mesh, time_step_method, rotational=True)
    s.order = {"velocity": 2.0, "pressure": 1.5}
    return s
