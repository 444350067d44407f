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


def newton(residual_jacobian, x0, bc_dofs, bc_vals, atol=1e-10, maxit=10, report=None, linear_solve=None):
    """DOLFIN NewtonSolver [EXT]: residual test on ||F||_2 with BC rows F_i = x_i - g_i,
    relaxation 1, sparse LU for the update, raise after `maxit` (pressure_correction.py:228-236).

    `linear_solve(J_csr, F) -> delta` replaces the LU factorisation where its fill is infeasible (3D meshes beyond
    ~1e5 dofs); it must solve to rounding level so that the Newton iterates are those of the LU path
    (see `tight_krylov_solve`)."""
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
        J = keep @ J + sp.diags(mask.astype(float))
        if linear_solve is None:
            x -= spla.splu(J.tocsc()).solve(F)
        else:
            x -= linear_solve(J.tocsr(), F, bc_dofs)
        its += 1
        F, _ = eval_F(False)
        r = np.linalg.norm(F)
        hist.append(r)
    if report is not None:
        report["newton_residuals"] = hist
    return x, its


# ---- "Krylov CPU" variant: the same algorithm the GPU path runs,
# ---- on the host cores through oracle/_cbaseline.so (C + OpenMP).  Used for CPU timing only.
_cb = None


def cbaseline():
    global _cb
    if _cb is None:
        import ctypes as C
        import os

        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cbaseline.so")
        lib = C.CDLL(path)
        pd, pi64, pi32 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        lib.cb_pcg.argtypes = [C.c_int64, pi64, pi32, pd, pd, pd, pd, C.c_double, C.c_int]
        lib.cb_bicgstab.argtypes = [C.c_int64, pi64, pi32, pd, pd, pd, pd, C.c_double, C.c_int]
        lib.cb_spmv.argtypes = [C.c_int64, pi64, pi32, pd, pd, pd]
        lib.cb_spmv.restype = None
        _cb = lib
    return _cb


def _csr_args(A):
    import ctypes as C

    A = sp.csr_matrix(A)
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    keep = (indptr, indices, data)
    return keep, (A.shape[0], indptr.ctypes.data_as(C.POINTER(C.c_int64)), indices.ctypes.data_as(C.POINTER(C.c_int32)),
                  data.ctypes.data_as(C.POINTER(C.c_double)))


def c_pcg(A, b, rtol, maxit):
    import ctypes as C

    pd = C.POINTER(C.c_double)
    keep, args = _csr_args(A)
    dinv = np.ascontiguousarray(1.0 / A.diagonal())
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    its = cbaseline().cb_pcg(*args, dinv.ctypes.data_as(pd), b.ctypes.data_as(pd), x.ctypes.data_as(pd), rtol, maxit)
    if its < 0:
        raise ConvergenceError("CG did not converge in %d iterations" % maxit)
    return x, its


def c_bicgstab(A, b, atol, maxit):
    import ctypes as C

    pd = C.POINTER(C.c_double)
    keep, args = _csr_args(A)
    dinv = np.ascontiguousarray(1.0 / A.diagonal())
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    its = cbaseline().cb_bicgstab(*args, dinv.ctypes.data_as(pd), b.ctypes.data_as(pd), x.ctypes.data_as(pd), atol, maxit)
    if its < 0:
        raise ConvergenceError("BiCGStab did not converge in %d iterations" % maxit)
    return x, its


def newton_krylov(residual_jacobian, x0, bc_dofs, bc_vals, atol=1e-10, maxit=10, report=None):
    """Same Newton loop as `newton`, linear solves by Jacobi-BiCGStab to 0.1*atol (the GPU path's setting)."""
    x = x0.copy()

    def eval_F(want_J):
        F, J = residual_jacobian(x, want_J)
        F = F.copy()
        F[bc_dofs] = x[bc_dofs] - bc_vals
        return F, J

    F, _ = eval_F(False)
    r = np.linalg.norm(F)
    its = 0
    kits = 0
    while r >= atol:
        if its >= maxit:
            raise ConvergenceError("Newton solver did not converge (|F| = %g)" % r)
        _, J = eval_F(True)
        mask = np.zeros(x.size, bool)
        mask[bc_dofs] = True
        keep = sp.diags((~mask).astype(float))
        J = (keep @ J + sp.diags(mask.astype(float))).tocsr()
        # lift the Dirichlet dofs (their update is known exactly); BiCGStab breaks down otherwise
        dg = np.zeros_like(F)
        dg[bc_dofs] = F[bc_dofs]
        b = F - J @ dg
        b[bc_dofs] = 0.0
        dx, k = c_bicgstab(J, b, 0.1 * atol, 1000)
        dx += dg
        kits += k
        x -= dx
        its += 1
        F, _ = eval_F(False)
        r = np.linalg.norm(F)
    if report is not None:
        report["momentum_its"] = kits
    return x, its


def tight_krylov_solve(J, F, bc_dofs, rel=1e-13, maxit=5000):
    """Stand-in for the LU solve of `newton` on meshes where LU fill is infeasible: Jacobi-BiCGStab
    (oracle/_cbaseline.so, C + OpenMP) driven to ||r|| <= rel * ||b|| (b = lifted right-hand side), i.e. to rounding level, restarted on the true
    residual until it is met.  The Dirichlet rows (identity) are lifted first, as their update is known exactly."""
    dg = np.zeros_like(F)
    dg[bc_dofs] = F[bc_dofs]
    b = F - J @ dg
    b[bc_dofs] = 0.0
    target = rel * np.linalg.norm(b)
    dx = np.zeros_like(F)
    r = b.copy()
    for _ in range(6):
        if np.linalg.norm(r) <= target:
            break
        try:
            e, _k = c_bicgstab(J, r, 0.3 * target, maxit)
        except ConvergenceError:
            break
        dx += e
        r = b - J @ dx
    if np.linalg.norm(r) > 50 * target:
        raise ConvergenceError("tight_krylov_solve stalled at %g (target %g)" % (np.linalg.norm(r), target))
    return dx + dg
