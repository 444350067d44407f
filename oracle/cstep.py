"""ctypes wrapper of oracle/_cstep.so: the compiled C++/OpenMP CPU implementation of one IPCS step
(BENCH INFRASTRUCTURE -- see oracle/cstep/cstep.cpp).  The numpy oracle provides the mesh, the P2 dof map, the
quadrature rule and the basis tables; patterns, constant matrices, assembly and solvers run in C++ on all host
threads.  Nothing here is imported by flow_b200."""
import ctypes as C
import os

import numpy as np

from . import fem

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cstep.so")
        if not os.path.exists(path):
            import subprocess

            subprocess.check_call(["make", "-C", os.path.join(os.path.dirname(path), "cstep")])
        L = C.CDLL(path)
        pd, pi64, pi32 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        L.cs_create.restype = C.c_void_p
        L.cs_create.argtypes = [C.c_int64, pi32, C.c_int64, C.c_int64, pd, C.c_int, pd, pd, pd, pd, C.c_int64, pi64, pd]
        L.cs_destroy.argtypes = [C.c_void_p]
        L.cs_destroy.restype = None
        L.cs_step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, pd, pd, C.c_double, pd, pd, C.POINTER(C.c_int)]
        L.cs_spmv_bench.argtypes = [C.c_void_p, C.c_int, pd, pd]
        L.cs_spmv_bench.restype = None
        L.cs_nnz.argtypes = [C.c_void_p, C.c_int]
        L.cs_nnz.restype = C.c_int64
        L.cs_set_threads.argtypes = [C.c_int]
        L.cs_set_threads.restype = None
        _lib = L
    return _lib


def set_threads(n):
    """OpenMP thread count of the CPU arm (torchrun exports OMP_NUM_THREADS=1; bench.py asks for all cores itself)."""
    lib().cs_set_threads(int(n))


class CavityCPU(object):
    """Lid-driven cavity on UnitCubeMesh(n), IPCS / backward Euler, stepped on the host cores."""

    def __init__(self, n):
        pd, pi64, pi32 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        self.mesh = fem.Mesh(*fem.unit_cube_mesh(n, n, n))
        m = self.mesh
        self.W = fem.Space(m, 2, 3)
        self.P = fem.Space(m, 1, 1)
        lam, w = fem.simplex_quadrature(3, 5)
        phi, dphi = fem.tabulate_p2(lam)
        bd = self.W.boundary_dofs().astype(np.int64)
        g = np.zeros((self.W.nnodes, 3))
        g[self.W.node_coords[:, 2] > 1 - 1e-12, 0] = 1.0
        self.bc = (np.ascontiguousarray(bd), np.ascontiguousarray(g.reshape(-1)[bd]))
        self.ndofs = self.W.ndofs + self.P.nnodes
        cn = np.ascontiguousarray(self.W.cell_nodes, dtype=np.int32)
        xyz = np.ascontiguousarray(m.points, dtype=np.float64)
        f = lambda a, t: np.ascontiguousarray(a).ctypes.data_as(t)  # noqa: E731
        lam, w, phi, dphi = (np.ascontiguousarray(a, dtype=np.float64) for a in (lam, w / w.sum(), phi, dphi))
        self.h = lib().cs_create(m.nc, f(cn, pi32), self.W.nnodes, self.P.nnodes, f(xyz, pd), lam.shape[0], f(lam, pd), f(w, pd),
                                 f(phi, pd), f(dphi, pd), self.bc[0].size, f(self.bc[0], pi64), f(self.bc[1], pd))
        if not self.h:
            raise RuntimeError("cs_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().cs_destroy(self.h)
            self.h = None

    def threads(self):
        return lib().cs_num_threads()

    def step(self, u0, p0, dt=1e-2, rho=1.0, mu=1e-2, tol=1e-10):
        pd = C.POINTER(C.c_double)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        u1, p1 = np.zeros_like(u0), np.zeros_like(p0)
        stats = (C.c_int * 4)()
        st = lib().cs_step(self.h, dt, rho, mu, u0.ctypes.data_as(pd), p0.ctypes.data_as(pd), tol, u1.ctypes.data_as(pd),
                           p1.ctypes.data_as(pd), stats)
        if st != 0:
            raise RuntimeError("cs_step failed with status %d" % st)
        return u1, p1, list(stats)

    def spmv_bench(self, reps=5):
        """OpenMP SpMV throughput (algorithmic GB/s, SURVEY.md 8d byte count) of the block Jacobian and of the scalar
        P2 mass matrix x 3 components."""
        gbs, ms = (C.c_double * 2)(), (C.c_double * 2)()
        lib().cs_spmv_bench(self.h, reps, gbs, ms)
        return {"jacobian_bsr3_GBs": gbs[0], "jacobian_ms": ms[0], "p2_mass_x3_GBs": gbs[1], "p2_mass_x3_ms": ms[1]}
