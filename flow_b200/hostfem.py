"""Host-side (numpy) data-entry helpers of the facade: mesh generators, Lagrange
tabulation, quadrature, load vectors, error norms.

These cover what the reference's *drivers* ask DOLFIN for outside the time-stepping hot
path: building structured meshes, turning an ``Expression`` into
numbers, measuring errors.  They feed host buffers to the C ABI; none of the hot-path
arithmetic (assembly, SpMV, Krylov) lives here.
"""
import math
from functools import lru_cache

import numpy as np
from scipy.special import roots_jacobi

# UFC local edges: triangle edge i is opposite vertex i; tetrahedron per UFC
LOCAL_EDGES = {2: ((1, 2), (0, 2), (0, 1)), 3: ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))}


# ------------------------------------------------------------------ meshes
def structured_rectangle(a, b, nx, ny, diagonal):
    """Vertices/cells of DOLFIN's RectangleMesh [EXT]: row-major grid vertices, optional
    cell-centre vertices ('crossed'), diagonal direction per cell otherwise."""
    xs = a[0] + (b[0] - a[0]) * np.arange(nx + 1) / nx
    ys = a[1] + (b[1] - a[1]) * np.arange(ny + 1) / ny
    pts = np.empty(((nx + 1) * (ny + 1), 2))
    pts[:, 0] = np.tile(xs, ny + 1)
    pts[:, 1] = np.repeat(ys, nx + 1)
    j, i = np.divmod(np.arange(nx * ny), nx)
    ll = j * (nx + 1) + i
    lr, ul, ur = ll + 1, ll + nx + 1, ll + nx + 2
    if diagonal == "crossed":
        cx = a[0] + (b[0] - a[0]) * (np.arange(nx) + 0.5) / nx
        cy = a[1] + (b[1] - a[1]) * (np.arange(ny) + 0.5) / ny
        mid = np.empty((nx * ny, 2))
        mid[:, 0] = np.tile(cx, ny)
        mid[:, 1] = np.repeat(cy, nx)
        pts = np.concatenate([pts, mid])
        cc = (nx + 1) * (ny + 1) + j * nx + i
        quad = np.array([[ll, lr, cc], [ll, ul, cc], [lr, ur, cc], [ul, ur, cc]])  # (4, 3, ncell)
        cells = quad.transpose(2, 0, 1).reshape(-1, 3)
    else:
        if diagonal == "right":
            use_right = np.ones(nx * ny, dtype=bool)
        elif diagonal == "left":
            use_right = np.zeros(nx * ny, dtype=bool)
        elif diagonal == "left/right":
            use_right = (i + j) % 2 == 0
        elif diagonal == "right/left":
            use_right = (i + j) % 2 == 1
        else:
            raise ValueError("unknown diagonal %r" % (diagonal,))
        t0 = np.where(use_right, [ll, lr, ur], [ll, lr, ul])
        t1 = np.where(use_right, [ll, ul, ur], [lr, ul, ur])
        cells = np.stack([t0.T, t1.T], axis=1).reshape(-1, 3)
    return pts, np.sort(cells, axis=1).astype(np.int32)


def structured_box(a, b, nx, ny, nz):
    """DOLFIN BoxMesh [EXT]: six tetrahedra per hexahedron sharing the main diagonal v0-v7."""
    xs = a[0] + (b[0] - a[0]) * np.arange(nx + 1) / nx
    ys = a[1] + (b[1] - a[1]) * np.arange(ny + 1) / ny
    zs = a[2] + (b[2] - a[2]) * np.arange(nz + 1) / nz
    return structured_box_lattice(xs, ys, zs)


def structured_box_lattice(xs, ys, zs):
    """The same mesh on given lattice coordinates (a cut-out of a larger box keeps the larger box's coordinates)."""
    nx, ny, nz = len(xs) - 1, len(ys) - 1, len(zs) - 1
    npl = (nx + 1) * (ny + 1)
    pts = np.empty((npl * (nz + 1), 3))
    pts[:, 0] = np.tile(xs, (ny + 1) * (nz + 1))
    pts[:, 1] = np.tile(np.repeat(ys, nx + 1), nz + 1)
    pts[:, 2] = np.repeat(zs, npl)
    idx = np.arange(nx * ny * nz)
    k, rem = np.divmod(idx, nx * ny)
    j, i = np.divmod(rem, nx)
    v = [None] * 8
    v[0] = k * npl + j * (nx + 1) + i
    v[1] = v[0] + 1
    v[2] = v[0] + nx + 1
    v[3] = v[2] + 1
    for m in range(4):
        v[4 + m] = v[m] + npl
    corner_sets = ((0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7))
    cells = np.stack([np.stack([v[c] for c in cs], axis=1) for cs in corner_sets], axis=1).reshape(-1, 4)
    return pts, np.sort(cells, axis=1).astype(np.int32)


# ------------------------------------------------------------------ quadrature
@lru_cache(maxsize=None)
def quadrature(dim, degree):
    """Collapsed Gauss-Jacobi rule on the reference simplex; returns barycentric points
    (nq, dim+1) and weights normalised to sum 1."""
    n = max(1, degree // 2 + 1)
    pts1 = []
    for alpha in range(dim - 1, -1, -1):
        x, w = roots_jacobi(n, alpha, 0)
        pts1.append((0.5 * (x + 1.0), w))
    grids = np.meshgrid(*[p for p, _ in pts1], indexing="ij")
    wgrids = np.meshgrid(*[w for _, w in pts1], indexing="ij")
    w = np.ones_like(grids[0])
    for wg in wgrids:
        w = w * wg
    coords = []
    scale = np.ones_like(grids[0])
    for g in grids:
        coords.append(scale * g)
        scale = scale * (1.0 - g)
    cart = np.stack([c.ravel() for c in coords], axis=1)
    lam = np.concatenate([1.0 - cart.sum(axis=1, keepdims=True), cart], axis=1)
    w = w.ravel()
    return lam, w / w.sum()


# ------------------------------------------------------------------ bases
def lagrange_p1(lam):
    return lam


def lagrange_p2(lam):
    dim = lam.shape[1] - 1
    cols = [lam[:, i] * (2.0 * lam[:, i] - 1.0) for i in range(dim + 1)]
    cols += [4.0 * lam[:, a] * lam[:, b] for a, b in LOCAL_EDGES[dim]]
    return np.stack(cols, axis=1)


def lagrange(degree, lam):
    return lagrange_p1(lam) if degree == 1 else lagrange_p2(lam)


@lru_cache(maxsize=None)
def pk_lattice(dim, k):
    """Barycentric multi-indices of the equispaced degree-k lattice."""
    if dim == 2:
        idx = [(k - i - j, i, j) for i in range(k + 1) for j in range(k + 1 - i)]
    else:
        idx = [(k - i - j - l, i, j, l) for i in range(k + 1) for j in range(k + 1 - i) for l in range(k + 1 - i - j)]
    return np.array(idx, dtype=np.int64)


def pk_basis(dim, k, lam):
    """Equispaced Lagrange basis of degree k evaluated at barycentric points lam."""
    if k == 0:
        return np.ones((lam.shape[0], 1))
    idx = pk_lattice(dim, k)
    out = np.empty((lam.shape[0], idx.shape[0]))
    for n, alpha in enumerate(idx):
        val = np.ones(lam.shape[0])
        for m, am in enumerate(alpha):
            for s in range(am):
                val = val * (k * lam[:, m] - s) / (s + 1.0)
        out[:, n] = val
    return out


# ------------------------------------------------------------------ per-cell data
def cell_volumes(points, cells):
    dim = points.shape[1]
    X = points[cells]
    E = X[:, 1:, :] - X[:, :1, :]
    return np.abs(np.linalg.det(E)) / math.factorial(dim)


def cell_points(points, cells, lam):
    """Physical coordinates of barycentric points lam in every cell: (nc, npts, dim)."""
    return np.einsum("pm,cmk->cpk", lam, points[cells])


def expression_at_quadrature(points, cells, func, degree, lam):
    """Values at barycentric points `lam` of the per-cell P_degree interpolant of func
    (DOLFIN's reading of Expression(..., degree=k) inside a form [EXT])."""
    dim = points.shape[1]
    nc = cells.shape[0]
    if degree == 0:
        ctr = points[cells].mean(axis=1)
        v = np.asarray(func(ctr), dtype=float).reshape(nc, 1, -1)
        return np.repeat(v, lam.shape[0], axis=1)
    nodes = pk_lattice(dim, degree) / float(degree)
    X = cell_points(points, cells, nodes)
    vals = np.asarray(func(X.reshape(-1, dim)), dtype=float).reshape(nc, nodes.shape[0], -1)
    return np.einsum("qn,cni->cqi", pk_basis(dim, degree, lam), vals)


def load_vector(points, cells, cell_nodes, nnodes, sdegree, ncomp, func, fdegree):
    """int I_k(f) . v dx for all test functions of the (vector) Lagrange space."""
    dim = points.shape[1]
    lam, w = quadrature(dim, fdegree + sdegree)
    fq = expression_at_quadrature(points, cells, func, fdegree, lam)
    if fq.shape[2] != ncomp:
        raise ValueError("expression has %d components, space has %d" % (fq.shape[2], ncomp))
    phi = lagrange(sdegree, lam)
    vol = cell_volumes(points, cells)
    be = np.einsum("q,cqi,qa,c->cai", w, fq, phi, vol)
    out = np.zeros(nnodes * ncomp)
    dofs = (cell_nodes[:, :, None] * ncomp + np.arange(ncomp)).reshape(cells.shape[0], -1)
    np.add.at(out, dofs.ravel(), be.reshape(cells.shape[0], -1).ravel())
    return out


def errornorm_l2(points, cells, cell_nodes, sdegree, uh, func, rise=3):
    """DOLFIN errornorm(u, uh) [EXT]: both interpolated into P_{k+3}, exact L2 norm of the difference."""
    dim = points.shape[1]
    k = sdegree + rise
    nodes = pk_lattice(dim, k) / float(k)
    X = cell_points(points, cells, nodes)
    nc = cells.shape[0]
    exact = np.asarray(func(X.reshape(-1, dim)), dtype=float).reshape(nc, nodes.shape[0], -1)
    U = uh.reshape(-1, exact.shape[2])[cell_nodes]
    approx = np.einsum("na,cai->cni", lagrange(sdegree, nodes), U)
    diff = exact - approx
    lam, w = quadrature(dim, 2 * k)
    B = pk_basis(dim, k, lam)
    G = np.einsum("q,qn,qm->nm", w, B, B)
    val = np.einsum("cni,nm,cmi,c->", diff, G, diff, cell_volumes(points, cells))
    return float(np.sqrt(max(val, 0.0)))


def integrate_nodal(points, cells, cell_nodes, sdegree, uh):
    dim = points.shape[1]
    lam, w = quadrature(dim, sdegree)
    phi = lagrange(sdegree, lam)
    return float(np.einsum("q,qa,ca,c->", w, phi, uh[cell_nodes], cell_volumes(points, cells)))


def integrate_expression(points, cells, func, degree):
    dim = points.shape[1]
    lam, w = quadrature(dim, max(degree, 1))
    fq = expression_at_quadrature(points, cells, func, degree, lam)[:, :, 0]
    return float(np.einsum("q,cq,c->", w, fq, cell_volumes(points, cells)))


def structured_rectangle_with_hole(a, b, nx, ny, center, radius, diagonal="left/right"):
    """Synthetic stand-in for the gmsh geometries of the reference's drivers (rectangle with a circular
    hole: tests/test_sealed_box.py:32-53, tests/test_boussinesq.py:25-79, the cylinder of
    tests/test_karman_vortex_street.py:18-45): a structured triangulation whose vertices near the circle
    are projected onto it; cells inside the circle are removed.  Returns (points, cells)."""
    pts, cells = structured_rectangle(a, b, nx, ny, diagonal)
    c = np.asarray(center, dtype=float)
    h = max((b[0] - a[0]) / nx, (b[1] - a[1]) / ny)
    d = pts - c
    r = np.sqrt((d * d).sum(axis=1))
    snap = np.abs(r - radius) < 0.5 * h
    pts = pts.copy()
    pts[snap] = c + d[snap] * (radius / r[snap])[:, None]
    r = np.sqrt(((pts - c) ** 2).sum(axis=1))
    inside_v = r < radius * (1 - 1e-9)
    cen = pts[cells].mean(axis=1)
    rc = np.sqrt(((cen - c) ** 2).sum(axis=1))
    keep = ~(inside_v[cells].any(axis=1) | (rc < radius * 0.98))
    cells = cells[keep]
    # drop degenerate cells created by snapping and renumber the used vertices
    vol = cell_volumes(pts, cells)
    # ... and slivers (longest edge^2 / area > 16; a right isosceles triangle has 4): on fine grids the radial
    # snap squashes a few cells next to the circle to aspect ratios of several hundred
    tri = pts[cells]
    emax2 = np.max([((tri[:, i] - tri[:, j]) ** 2).sum(axis=1) for i, j in ((0, 1), (1, 2), (0, 2))], axis=0)
    cells = cells[(vol > 1e-3 * h * h) & (emax2 < 16.0 * np.maximum(vol, 1e-300))]
    used = np.unique(cells)
    remap = -np.ones(pts.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.size)
    return pts[used], np.sort(remap[cells], axis=1).astype(np.int32)


def read_msh(path_or_text):
    """Minimal gmsh MSH 2.x ASCII reader: returns (points, cells) of the highest-dimensional
    simplices (type 4 tetrahedra, else type 2 triangles).  Enough to load the meshes the reference's drivers
    generate with pygmsh (tests/test_sealed_box.py:32-53) when a .msh file is at hand; gmsh itself is not needed."""
    import os

    text = open(path_or_text).read() if os.path.exists(str(path_or_text)) else str(path_or_text)
    lines = text.splitlines()

    def section(name):
        i = lines.index("$" + name)
        j = lines.index("$End" + name)
        return lines[i + 1:j]

    fmt = section("MeshFormat")[0].split()
    if not fmt[0].startswith("2") or fmt[1] != "0":
        raise ValueError("only MSH 2.x ASCII files are supported (got %r)" % (fmt,))
    nodes = section("Nodes")
    n = int(nodes[0])
    ids = np.empty(n, dtype=np.int64)
    xyz = np.empty((n, 3))
    for k, ln in enumerate(nodes[1:n + 1]):
        t = ln.split()
        ids[k] = int(t[0])
        xyz[k] = [float(t[1]), float(t[2]), float(t[3])]
    elems = section("Elements")
    tris, tets = [], []
    for ln in elems[1:int(elems[0]) + 1]:
        t = ln.split()
        etype, ntags = int(t[1]), int(t[2])
        conn = [int(v) for v in t[3 + ntags:]]
        if etype == 2:
            tris.append(conn)
        elif etype == 4:
            tets.append(conn)
    cells = np.array(tets if tets else tris, dtype=np.int64)
    if cells.size == 0:
        raise ValueError("no triangles or tetrahedra in the file")
    lookup = -np.ones(ids.max() + 1, dtype=np.int64)
    lookup[ids] = np.arange(n)
    cells = lookup[cells]
    dim = 3 if tets else (2 if np.abs(xyz[:, 2]).max() == 0.0 else 3)
    used = np.unique(cells)
    remap = -np.ones(n, dtype=np.int64)
    remap[used] = np.arange(used.size)
    pts = xyz[used][:, :2] if (dim == 2 and not tets) else xyz[used]
    return np.ascontiguousarray(pts), np.sort(remap[cells], axis=1).astype(np.int32)
