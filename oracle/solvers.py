"""Oracle linear / nonlinear solvers (TEST INFRASTRUCTURE).

Restates the solver settings the reference passes to DOLFIN/PETSc:
Newton (pressure_correction.py:224-254), CG (:325-339, :414-432, :451-464),
LU (heat.py:117-121), GMRES (stokes.py:59-143).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class ConvergenceError(RuntimeError):
    """error_on_nonconvergence: True (pressure_correction.py:236, 337, 424, 462)."""


def pcg(A, b, rtol, maxit, Minv=None, x0=None):
    """Preconditioned CG, PETSc-style stopping test [EXT]: ||M^-1 r||_2 <= rtol ||M^-1 b||_2,
    atol = 0.  M^-1 defaults to Jacobi (the reference's hypre BoomerAMG is not available;
    the converged x agrees to O(rtol))."""
    n = b.shape[0]
    if Minv is None:
        dinv = 1.0 / A.diagonal()
        Minv = lambda r: dinv * r  # noqa: E731
    x = np.zeros(n) if x0 is None else x0.copy()
    r = b - A @ x
    z = Minv(r)
    ref = np.linalg.norm(Minv(b))
    if ref == 0.0:
        return x, 0
    p = z.copy()
    rz = r @ z
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = Minv(r)
        if np.linalg.norm(z) <= rtol * ref:
            return x, it
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    raise ConvergenceError("CG did not converge in %d iterations" % maxit)


def lu_solve(A, b):
    return spla.splu(sp.csc_matrix(A)).solve(b)


def newton(residual_jacobian, x0, bc_dofs, bc_vals, atol=1e-10, maxit=10, report=None):
    """DOLFIN NewtonSolver [EXT]: residual test on ||F||_2 with BC rows F_i = x_i - g_i,
    relaxation 1, sparse LU for the update, raise after `maxit` (pressure_correction.py:228-236)."""
    x = x0.copy()

    def eval_F(want_J):
        F, J = residual_jacobian(x, want_J)
        F = F.copy()
        F[bc_dofs] = x[bc_dofs] - bc_vals
        return F, J

    F, _ = eval_F(False)
    r = np.linalg.norm(F)
    its = 0
    hist = [r]
    while r >= atol:
        if its >= maxit:
            raise ConvergenceError("Newton solver did not converge (|F| = %g)" % r)
        _, J = eval_F(True)
        mask = np.zeros(x.size, bool)
        mask[bc_dofs] = True
        keep = sp.diags((~mask).astype(float))
        J = (keep @ J + sp.diags(mask.astype(float))).tocsc()
        x -= spla.splu(J).solve(F)
        its += 1
        F, _ = eval_F(False)
        r = np.linalg.norm(F)
        hist.append(r)
    if report is not None:
        report["newton_residuals"] = hist
    return x, its
