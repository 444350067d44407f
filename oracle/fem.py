"""Oracle building blocks: structured simplicial meshes, Lagrange tables,
dof maps, quadrature.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Semantics restated from DOLFIN 2017.x (third-party, not in /root/reference;
every choice below is listed as [EXT] in SURVEY.md section 8c):

* mesh generators ``RectangleMesh``/``UnitSquareMesh``/``BoxMesh`` -- used at
  /root/reference/tests/test_navier_stokes.py:82,144,176 and
  /root/reference/tests/test_stokes.py:65,98;
* UFC local ordering: cell vertices ascending by global index, triangle edge i
  opposite vertex i, tet edges (2,3),(1,3),(1,2),(0,3),(0,2),(0,1), P2 local
  dofs = vertices then edges;
* canonical global numbering (ours, DOLFIN's graph reordering is not
  reproducible): vertex nodes first in vertex order, then edge nodes in
  lexicographic (min,max) order; vector dofs interleaved ``d*node+comp``.
"""
import math
from functools import lru_cache

import numpy as np
from scipy.special import roots_jacobi


# --------------------------------------------------------------------------
# meshes
# --------------------------------------------------------------------------
def rectangle_mesh(p0, p1, nx, ny, diagonal="right"):
    """DOLFIN RectangleMesh(Point(p0), Point(p1), nx, ny, diagonal) [EXT]."""
    x = np.linspace(p0[0], p1[0], nx + 1)
    y = np.linspace(p0[1], p1[1], ny + 1)
    # x = a + ix*(b-a)/nx exactly as DOLFIN computes it
    x = p0[0] + (p1[0] - p0[0]) * np.arange(nx + 1) / nx
    y = p0[1] + (p1[1] - p0[1]) * np.arange(ny + 1) / ny
    X, Y = np.meshgrid(x, y, indexing="xy")
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ix = ix.ravel()
    iy = iy.ravel()
    v0 = iy * (nx + 1) + ix
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v1 + (nx + 1)
    if diagonal == "crossed":
        xm = p0[0] + (p1[0] - p0[0]) * (np.arange(nx) + 0.5) / nx
        ym = p0[1] + (p1[1] - p0[1]) * (np.arange(ny) + 0.5) / ny
        XM, YM = np.meshgrid(xm, ym, indexing="xy")
        pts = np.vstack([pts, np.stack([XM.ravel(), YM.ravel()], axis=1)])
        vm = (nx + 1) * (ny + 1) + iy * nx + ix
        cells = np.stack(
            [
                np.stack([v0, v1, vm], 1),
                np.stack([v0, v2, vm], 1),
                np.stack([v1, v3, vm], 1),
                np.stack([v2, v3, vm], 1),
            ],
            axis=1,
        ).reshape(-1, 3)
    else:
        if diagonal == "right":
            right = np.ones(nx * ny, bool)
        elif diagonal == "left":
            right = np.zeros(nx * ny, bool)
        elif diagonal == "left/right":
            right = ((ix + iy) % 2) == 0
        elif diagonal == "right/left":
            right = ((ix + iy) % 2) == 1
        else:
            raise ValueError("unknown diagonal %r" % diagonal)
        a = np.where(right[:, None], np.stack([v0, v1, v3], 1), np.stack([v0, v1, v2], 1))
        b = np.where(right[:, None], np.stack([v0, v2, v3], 1), np.stack([v1, v2, v3], 1))
        cells = np.stack([a, b], axis=1).reshape(-1, 3)
    cells = np.sort(cells, axis=1).astype(np.int32)
    return pts, cells


def unit_square_mesh(nx, ny, diagonal="right"):
    return rectangle_mesh((0.0, 0.0), (1.0, 1.0), nx, ny, diagonal)


def box_mesh(p0, p1, nx, ny, nz):
    """DOLFIN BoxMesh: 6 tetrahedra per cube [EXT]."""
    x = p0[0] + (p1[0] - p0[0]) * np.arange(nx + 1) / nx
    y = p0[1] + (p1[1] - p0[1]) * np.arange(ny + 1) / ny
    z = p0[2] + (p1[2] - p0[2]) * np.arange(nz + 1) / nz
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    v0 = iz * sy + iy * sx + ix
    v1 = v0 + 1
    v2 = v0 + sx
    v3 = v1 + sx
    v4 = v0 + sy
    v5 = v1 + sy
    v6 = v2 + sy
    v7 = v3 + sy
    tets = [
        (v0, v1, v3, v7),
        (v0, v1, v7, v5),
        (v0, v5, v7, v4),
        (v0, v3, v2, v7),
        (v0, v6, v4, v7),
        (v0, v2, v6, v7),
    ]
    cells = np.stack([np.stack(t, 1) for t in tets], axis=1).reshape(-1, 4)
    cells = np.sort(cells, axis=1).astype(np.int32)
    return pts, cells


def unit_cube_mesh(nx, ny, nz):
    return box_mesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), nx, ny, nz)


# --------------------------------------------------------------------------
# quadrature (conical-product Gauss-Jacobi: positive weights, any degree)
# --------------------------------------------------------------------------
@lru_cache(maxsize=None)
def simplex_quadrature(dim, degree):
    """Barycentric points (nq, dim+1) and weights (sum 1) exact to `degree`."""
    n = max(1, (degree + 2) // 2)
    if dim == 1:
        xi, w = roots_jacobi(n, 0, 0)
        x = 0.5 * (xi + 1)
        lam = np.stack([1 - x, x], 1)
        return lam, 0.5 * w
    if dim == 2:
        xa, wa = roots_jacobi(n, 1, 0)
        xb, wb = roots_jacobi(n, 0, 0)
        xa = 0.5 * (xa + 1)
        wa = wa * 0.25
        xb = 0.5 * (xb + 1)
        wb = wb * 0.5
        X, T = np.meshgrid(xa, xb, indexing="ij")
        W = np.outer(wa, wb)
        x = X.ravel()
        y = ((1 - X) * T).ravel()
        lam = np.stack([1 - x - y, x, y], 1)
        w = W.ravel()
        return lam, w / w.sum()
    if dim == 3:
        xa, wa = roots_jacobi(n, 2, 0)
        xb, wb = roots_jacobi(n, 1, 0)
        xc, wc = roots_jacobi(n, 0, 0)
        xa = 0.5 * (xa + 1)
        xb = 0.5 * (xb + 1)
        xc = 0.5 * (xc + 1)
        A, B, C = np.meshgrid(xa, xb, xc, indexing="ij")
        W = wa[:, None, None] * wb[None, :, None] * wc[None, None, :]
        x = A
        y = (1 - A) * B
        z = (1 - A) * (1 - B) * C
        lam = np.stack([(1 - x - y - z).ravel(), x.ravel(), y.ravel(), z.ravel()], 1)
        w = W.ravel()
        return lam, w / w.sum()
    raise ValueError(dim)


# --------------------------------------------------------------------------
# Lagrange bases in barycentric coordinates
# --------------------------------------------------------------------------
TRI_EDGES = ((1, 2), (0, 2), (0, 1))
TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))


def local_edges(dim):
    return TRI_EDGES if dim == 2 else TET_EDGES


def tabulate_p1(lam):
    """phi (nq, d+1), dphi/dlam (nq, d+1, d+1)."""
    nq, nv = lam.shape
    return lam.copy(), np.broadcast_to(np.eye(nv), (nq, nv, nv)).copy()


def tabulate_p2(lam):
    """UFC-ordered P2 basis: phi (nq, nl), dphi/dlam (nq, nl, d+1)."""
    nq, nv = lam.shape
    dim = nv - 1
    edges = local_edges(dim)
    nl = nv + len(edges)
    phi = np.zeros((nq, nl))
    dphi = np.zeros((nq, nl, nv))
    for i in range(nv):
        phi[:, i] = lam[:, i] * (2 * lam[:, i] - 1)
        dphi[:, i, i] = 4 * lam[:, i] - 1
    for e, (a, b) in enumerate(edges):
        phi[:, nv + e] = 4 * lam[:, a] * lam[:, b]
        dphi[:, nv + e, a] = 4 * lam[:, b]
        dphi[:, nv + e, b] = 4 * lam[:, a]
    return phi, dphi


def p2_second_derivs(dim):
    """d2phi/dlam_m dlam_n, constant: (nl, d+1, d+1)."""
    nv = dim + 1
    edges = local_edges(dim)
    H = np.zeros((nv + len(edges), nv, nv))
    for i in range(nv):
        H[i, i, i] = 4.0
    for e, (a, b) in enumerate(edges):
        H[nv + e, a, b] = 4.0
        H[nv + e, b, a] = 4.0
    return H


@lru_cache(maxsize=None)
def lattice(dim, k):
    """Multi-indices alpha (|alpha|=k) of the equispaced P_k lattice, (n, dim+1)."""
    out = []

    def rec(prefix, left, slots):
        if slots == 1:
            out.append(prefix + (left,))
            return
        for a in range(left, -1, -1):
            rec(prefix + (a,), left - a, slots - 1)

    rec((), k, dim + 1)
    return np.array(out, dtype=np.int64)


def tabulate_pk(dim, k, lam):
    """Equispaced Lagrange P_k basis at barycentric points lam: (nq, n_k)."""
    if k == 0:
        return np.ones((lam.shape[0], 1))
    al = lattice(dim, k)
    out = np.ones((lam.shape[0], al.shape[0]))
    for n, alpha in enumerate(al):
        for m, am in enumerate(alpha):
            for j in range(am):
                out[:, n] *= (k * lam[:, m] - j) / (j + 1.0)
    return out


# --------------------------------------------------------------------------
# topology and dof maps
# --------------------------------------------------------------------------
class Mesh:
    def __init__(self, points, cells):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.cells = np.ascontiguousarray(np.sort(cells, axis=1), dtype=np.int32)
        self.dim = self.points.shape[1]
        assert self.cells.shape[1] == self.dim + 1
        self.nv = self.points.shape[0]
        self.nc = self.cells.shape[0]
        self._build_edges()
        self._build_geometry()
        self._build_boundary()

    def _build_edges(self):
        le = np.array(local_edges(self.dim))
        ev = self.cells[:, le]  # (nc, ne, 2) ascending because cells sorted
        key = ev[..., 0].astype(np.int64) * self.nv + ev[..., 1]
        uniq, inv = np.unique(key.ravel(), return_inverse=True)
        self.edges = np.stack([uniq // self.nv, uniq % self.nv], 1).astype(np.int32)
        self.cell_edges = inv.reshape(key.shape).astype(np.int32)
        self.ne = self.edges.shape[0]

    def _build_geometry(self):
        d = self.dim
        X = self.points[self.cells]  # (nc, d+1, d)
        Jm = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))  # columns = edge vectors
        det = np.linalg.det(Jm)
        Jinv = np.linalg.inv(Jm)  # rows = grad of lambda_1..d
        glam = np.empty((self.nc, d + 1, d))
        glam[:, 1:, :] = Jinv
        glam[:, 0, :] = -Jinv.sum(axis=1)
        self.glam = glam
        self.vol = np.abs(det) / math.factorial(d)

    def _build_boundary(self):
        d = self.dim
        nc = self.nc
        # facet f = the one opposite local vertex f
        fac = np.stack([np.delete(self.cells, f, axis=1) for f in range(d + 1)], axis=1)  # (nc,d+1,d)
        key = np.zeros((nc, d + 1), dtype=np.int64)
        for j in range(d):
            key = key * self.nv + fac[..., j]
        flat = key.ravel()
        uniq, inv, cnt = np.unique(flat, return_inverse=True, return_counts=True)
        on_b = (cnt[inv] == 1).reshape(nc, d + 1)
        c, f = np.nonzero(on_b)
        self.bfacet_cell = c.astype(np.int32)
        self.bfacet_local = f.astype(np.int32)
        self.bfacet_verts = fac[c, f]
        bv = np.zeros(self.nv, bool)
        bv[self.bfacet_verts.ravel()] = True
        self.boundary_vertex = bv
        # boundary edges: edges of boundary facets
        be = np.zeros(self.ne, bool)
        if d == 2:
            # the facet opposite vertex f IS local edge f
            be[self.cell_edges[c, f]] = True
        else:
            le = np.array(TET_EDGES)
            for floc in range(4):
                sel = f == floc
                for e in range(6):
                    if floc not in le[e]:
                        be[self.cell_edges[c[sel], e]] = True
        self.boundary_edge = be

    def hmax(self):
        X = self.points[self.edges]
        return float(np.sqrt(((X[:, 0] - X[:, 1]) ** 2).sum(1)).max())


class Space:
    """Scalar node space P1 or P2 on a Mesh; `ncomp` interleaved components."""

    def __init__(self, mesh, degree, ncomp=1):
        self.mesh = mesh
        self.degree = degree
        self.ncomp = ncomp
        if degree == 1:
            self.cell_nodes = mesh.cells.copy()
            self.nnodes = mesh.nv
            self.node_coords = mesh.points.copy()
            self.boundary_node = mesh.boundary_vertex.copy()
        elif degree == 2:
            self.cell_nodes = np.hstack([mesh.cells, mesh.nv + mesh.cell_edges]).astype(np.int32)
            self.nnodes = mesh.nv + mesh.ne
            mid = 0.5 * (mesh.points[mesh.edges[:, 0]] + mesh.points[mesh.edges[:, 1]])
            self.node_coords = np.vstack([mesh.points, mid])
            self.boundary_node = np.concatenate([mesh.boundary_vertex, mesh.boundary_edge])
        else:
            raise ValueError(degree)
        self.nl = self.cell_nodes.shape[1]
        self.ndofs = self.nnodes * ncomp

    def tabulate(self, lam):
        return tabulate_p1(lam) if self.degree == 1 else tabulate_p2(lam)

    def dof_coords(self):
        return np.repeat(self.node_coords, self.ncomp, axis=0)

    def boundary_dofs(self, comp=None):
        nodes = np.nonzero(self.boundary_node)[0]
        if self.ncomp == 1:
            return nodes
        comps = range(self.ncomp) if comp is None else [comp]
        return np.sort(np.concatenate([nodes * self.ncomp + c for c in comps]))

    def pattern(self):
        """Node-level CSR pattern (indptr, indices), columns ascending."""
        import scipy.sparse as sp

        nl = self.nl
        rows = np.repeat(self.cell_nodes, nl, axis=1).ravel()
        cols = np.tile(self.cell_nodes, (1, nl)).ravel()
        A = sp.csr_matrix((np.ones(rows.size, np.int8), (rows, cols)), shape=(self.nnodes,) * 2)
        A.sum_duplicates()
        A.sort_indices()
        return A.indptr.astype(np.int64), A.indices.astype(np.int32)
