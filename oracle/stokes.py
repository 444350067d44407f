"""Oracle restatement of flow.stokes.solve (TEST INFRASTRUCTURE).

Follows /root/reference/flow/stokes.py:40-46 (saddle-point form, symmetric Dirichlet
elimination through assemble_system) and :134-146 (GMRES to rtol 1e-13, here a sparse LU
solve of the same system, which the reference's iteration converges to).
"""
import numpy as np
import scipy.sparse as sp

from . import fem, forms, solvers


def solve(mesh, mu, load, u_bc, p_bc):
    """Returns (u, p).  load = int f.v dx; u_bc / p_bc = (dofs, values) or None."""
    assert mu > 0.0  # stokes.py:23
    W = fem.Space(mesh, 2, mesh.dim)
    P = fem.Space(mesh, 1, 1)
    A, B, _ = forms.stokes_blocks(W, P, mu)
    nu, npp = W.ndofs, P.nnodes
    S = sp.bmat([[A, B.T], [B, None]], format="csr")
    b = np.concatenate([load, np.zeros(npp)])
    dofs, vals = [], []
    if u_bc is not None and len(u_bc[0]):
        dofs.append(np.asarray(u_bc[0]))
        vals.append(np.asarray(u_bc[1]))
    if p_bc is not None and len(p_bc[0]):
        dofs.append(nu + np.asarray(p_bc[0]))
        vals.append(np.asarray(p_bc[1]))
    if dofs:
        S, b = forms.apply_bc_symmetric(S, b, np.concatenate(dofs), np.concatenate(vals))
    x = solvers.lu_solve(S, b)
    return x[:nu], x[nu:]
