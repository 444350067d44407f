"""ctypes wrapper of oracle/_cstep.so: the compiled C++/OpenMP CPU baseline of one IPCS step
(BENCH INFRASTRUCTURE -- see oracle/cstep/cstep.cpp).  The numpy oracle provides meshes, dof maps,
sparsity patterns and the constant matrices; the C++ side assembles and solves with all host threads."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import fem, forms

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cstep.so")
        if not os.path.exists(path):
            import subprocess

            subprocess.check_call(["make", "-C", os.path.join(os.path.dirname(path), "cstep")])
        _lib = C.CDLL(path)
    return _lib


class CavityCPU(object):
    """Lid-driven cavity on UnitCubeMesh(n), IPCS / backward Euler, stepped on the host cores."""

    def __init__(self, n):
        self.mesh = fem.Mesh(*fem.unit_cube_mesh(n, n, n))
        m = self.mesh
        self.W = fem.Space(m, 2, 3)
        self.P = fem.Space(m, 1, 1)
        Wn = fem.Space(m, 2, 1)
        indptr, indices = Wn.pattern()
        node = sp.csr_matrix((np.ones(indices.size), indices, indptr), shape=(Wn.nnodes,) * 2)
        Jp = sp.kron(node, np.ones((3, 3)), format="csr")
        Jp.sort_indices()
        self.Jptr = Jp.indptr.astype(np.int64)
        self.Jidx = Jp.indices.astype(np.int32)
        self.Jval = np.zeros(self.Jidx.size)
        A = forms.stiffness_matrix(self.P).tocsr()
        A.sort_indices()
        self.A = (A.indptr.astype(np.int64), A.indices.astype(np.int32), np.ascontiguousarray(A.data))
        M = sp.kron(forms.mass_matrix(Wn), sp.eye(3), format="csr")
        M.sort_indices()
        self.M = (M.indptr.astype(np.int64), M.indices.astype(np.int32), np.ascontiguousarray(M.data))
        bd = self.W.boundary_dofs().astype(np.int64)
        g = np.zeros((self.W.nnodes, 3))
        g[self.W.node_coords[:, 2] > 1 - 1e-12, 0] = 1.0
        self.bc = (np.ascontiguousarray(bd), np.ascontiguousarray(g.reshape(-1)[bd]))
        self.ndofs = self.W.ndofs + self.P.nnodes
        self.cell_nodes = np.ascontiguousarray(self.W.cell_nodes, dtype=np.int32)
        self.xyz = np.ascontiguousarray(m.points)

    def threads(self):
        return lib().cs_num_threads()

    def step(self, u0, p0, dt=1e-2, rho=1.0, mu=1e-2, tol=1e-10):
        pd = C.POINTER(C.c_double)
        pi64, pi32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        f = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
        u1, p1 = np.zeros_like(u0), np.zeros_like(p0)
        stats = (C.c_int * 4)()
        st = lib().cs_ipcs_step(
            C.c_int64(self.mesh.nc), f(self.cell_nodes, pi32), f(self.xyz, pd), C.c_int64(self.W.ndofs), C.c_int64(self.P.nnodes),
            f(self.Jptr, pi64), f(self.Jidx, pi32), f(self.Jval, pd), f(self.A[0], pi64), f(self.A[1], pi32), f(self.A[2], pd),
            f(self.M[0], pi64), f(self.M[1], pi32), f(self.M[2], pd), C.c_double(dt), C.c_double(rho), C.c_double(mu),
            f(u0, pd), f(p0, pd), C.c_int64(self.bc[0].size), f(self.bc[0], pi64), f(self.bc[1], pd), C.c_double(tol),
            f(u1, pd), f(p1, pd), stats)
        if st != 0:
            raise RuntimeError("cs_ipcs_step failed with status %d" % st)
        return u1, p1, list(stats)
