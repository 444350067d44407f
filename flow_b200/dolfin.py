"""A thin stand-in for the DOLFIN names that flow's API takes as arguments.

The reference's boundary is its Python API (`flow.navier_stokes.*.step`,
`flow.heat.Heat`, `flow.stokes.solve`), whose arguments are DOLFIN objects
(SURVEY.md 8b).  This module supplies just those objects -- meshes, Lagrange
function spaces, functions, constants, nodal expressions, Dirichlet conditions --
as light host-side containers around handles of the C ABI (`include/flowb200.h`).
Names and call signatures follow the imports of the reference's tests
(tests/test_navier_stokes.py:10-14, tests/test_sealed_box.py:9-13,
tests/test_boussinesq.py:14-18, tests/test_stokes.py:5-8).
"""
import ctypes as C
import math

import numpy as np

from . import _lib, hostfem
from ._lib import lib

DOLFIN_EPS = 3.0e-16
pi = math.pi
triangle = "triangle"
tetrahedron = "tetrahedron"


# ---------------------------------------------------------------- logging
_indent = [0]


def info(msg):
    print("  " * _indent[0] + str(msg))


def begin(msg):
    info(msg)
    _indent[0] += 1


def end():
    _indent[0] = max(0, _indent[0] - 1)


# ---------------------------------------------------------------- mesh
class Point(object):
    def __init__(self, *xs):
        self.x = tuple(float(v) for v in xs)

    def __getitem__(self, i):
        return self.x[i]

    def __len__(self):
        return len(self.x)


class Mesh(object):
    """Simplicial mesh given by vertex coordinates and cell connectivity."""

    def __init__(self, points, cells, device=None):
        pts = np.ascontiguousarray(points, dtype=np.float64)
        cl = np.ascontiguousarray(np.sort(np.asarray(cells), axis=1), dtype=np.int32)
        self._points, self._cells = pts, cl
        self.dim = pts.shape[1]
        self.ctx = _lib.context(device)
        h = _lib.vp()
        _lib.check(
            lib.fb_mesh_create(self.ctx, self.dim, pts.shape[0], _lib.as_pd(pts), cl.shape[0], _lib.as_pi32(cl), C.byref(h)),
            self.ctx, "fb_mesh_create")
        self.handle = h
        self._spaces = {}
        self._vol = None
        self.partition = None  # set by flow_b200.parallel.distributed_mesh
        # None: canonical node numbering (vertices, then edges in lexicographic order -- the numbering the oracle and
        # the parity tests use).  "lexicographic": nodes numbered by coordinate (z, then y, then x): the columns of a
        # matrix row and the rows a warp works on then sit in neighbouring cache lines (experimental, single GPU;
        # vectors are in that numbering, like the rank-local vectors of a partitioned run).  Set before the first
        # FunctionSpace is created on the mesh.
        self.node_order = None

    def coordinates(self):
        return self._points

    def cells(self):
        return self._cells

    def num_vertices(self):
        return self._points.shape[0]

    def num_cells(self):
        return self._cells.shape[0]

    def ufl_cell(self):
        return triangle if self.dim == 2 else tetrahedron

    def geometry(self):
        return self

    def topology(self):
        return self

    def volumes(self):
        if self._vol is None:
            self._vol = hostfem.cell_volumes(self._points, self._cells)
        return self._vol

    def _edge_lengths(self):
        n = _lib.i64()
        lib.fb_mesh_info(self.handle, None, None, C.byref(n), None)
        p = _lib.pi32()
        lib.fb_mesh_edges(self.handle, C.byref(p))
        e = np.ctypeslib.as_array(p, shape=(n.value, 2))
        d = self._points[e[:, 0]] - self._points[e[:, 1]]
        return np.sqrt((d * d).sum(axis=1))

    def hmax(self):
        return float(self._edge_lengths().max())

    def hmin(self):
        return float(self._edge_lengths().min())

    def node_space(self, degree):
        """Shared scalar node space (dof map, pattern, device copy) of a degree."""
        if degree not in self._spaces:
            self._spaces[degree] = _NodeSpace(self, degree)
        return self._spaces[degree]


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    return Mesh(*hostfem.structured_rectangle((p0[0], p0[1]), (p1[0], p1[1]), nx, ny, diagonal))


def UnitSquareMesh(nx, ny, diagonal="right"):
    return Mesh(*hostfem.structured_rectangle((0.0, 0.0), (1.0, 1.0), nx, ny, diagonal))


def BoxMesh(p0, p1, nx, ny, nz):
    return Mesh(*hostfem.structured_box(tuple(p0[i] for i in range(3)), tuple(p1[i] for i in range(3)), nx, ny, nz))


def UnitCubeMesh(nx, ny, nz):
    return Mesh(*hostfem.structured_box((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), nx, ny, nz))


def RectangleWithHoleMesh(p0, p1, nx, ny, center, radius, diagonal="left/right"):
    """Synthetic replacement of the pygmsh rectangle-with-circular-hole geometries used by the reference's
    drivers (gmsh is not available); see hostfem.structured_rectangle_with_hole."""
    return Mesh(*hostfem.structured_rectangle_with_hole((p0[0], p0[1]), (p1[0], p1[1]), nx, ny, center, radius, diagonal))


def MshMesh(path_or_text):
    """Mesh from a gmsh MSH 2.x ASCII file (or its text)."""
    return Mesh(*hostfem.read_msh(path_or_text))


class _NodeSpace(object):
    """fb_space handle of a scalar Lagrange node set (P1 or P2) + host views of its tables."""

    def __init__(self, mesh, degree):
        self.mesh, self.degree = mesh, degree
        self._vec_handles = {}
        self.plan = mesh.partition.plans[degree] if mesh.partition is not None else None
        self.perm = None  # canonical -> this numbering, when mesh.node_order asks for one (single GPU)
        if self.plan is None and getattr(mesh, "node_order", None):
            self.perm = self._coordinate_order(mesh.node_order)
        self.handle = self._create(1)
        h = self.handle
        nn, nd, nl = _lib.i64(), _lib.i64(), C.c_int()
        lib.fb_space_info(h, C.byref(nn), C.byref(nd), C.byref(nl))
        self.nnodes, self.nl = nn.value, nl.value
        p = _lib.pi32()
        lib.fb_space_dofmap(h, C.byref(p))
        self.cell_nodes = np.ctypeslib.as_array(p, shape=(mesh.num_cells(), self.nl))
        q = _lib.pd()
        lib.fb_space_node_coords(h, C.byref(q))
        self.coords = np.ctypeslib.as_array(q, shape=(self.nnodes, mesh.dim))
        b = _lib.pu8()
        lib.fb_space_boundary_nodes(h, C.byref(b))
        self.on_boundary = np.ctypeslib.as_array(b, shape=(self.nnodes,)).astype(bool)
        self._mass = None

    def _coordinate_order(self, kind):
        """perm[canonical node] = new index for a numbering by coordinate."""
        if kind != "lexicographic":
            raise ValueError("unknown node order %r" % (kind,))
        mesh = self.mesh
        h = _lib.vp()
        _lib.check(lib.fb_space_create(mesh.handle, self.degree, 1, C.byref(h)), mesh.ctx, "fb_space_create")
        nn, nd, nl = _lib.i64(), _lib.i64(), C.c_int()
        lib.fb_space_info(h, C.byref(nn), C.byref(nd), C.byref(nl))
        q = _lib.pd()
        lib.fb_space_node_coords(h, C.byref(q))
        X = np.ctypeslib.as_array(q, shape=(nn.value, mesh.dim)).copy()
        lib.fb_space_destroy(h)
        order = np.lexsort(tuple(X[:, k] for k in range(mesh.dim)))  # last key (highest coordinate index) is the primary one
        perm = np.empty(nn.value, dtype=np.int32)
        perm[order] = np.arange(nn.value, dtype=np.int32)
        return perm

    def _create(self, ncomp):
        h = _lib.vp()
        mesh, pl = self.mesh, self.plan
        if pl is None and self.perm is not None:
            _lib.check(lib.fb_space_create_numbered(mesh.handle, self.degree, ncomp, _lib.as_pi32(self.perm), self.perm.size,
                                                    C.byref(h)), mesh.ctx, "fb_space_create_numbered")
            return h
        if pl is None:
            _lib.check(lib.fb_space_create(mesh.handle, self.degree, ncomp, C.byref(h)), mesh.ctx, "fb_space_create")
            return h
        # distributed: owned-first numbering + halo plan (flow_b200/parallel.py)
        _lib.check(lib.fb_space_create_numbered(mesh.handle, self.degree, ncomp, _lib.as_pi32(pl.perm), pl.n_owned, C.byref(h)),
                   mesh.ctx, "fb_space_create_numbered")
        _lib.check(lib.fb_space_set_halo(h, pl.ranks.size, _lib.as_pi32(pl.ranks), _lib.as_pi64(pl.send_ptr),
                                         _lib.as_pi32(pl.send_nodes), _lib.as_pi64(pl.recv_ptr)), mesh.ctx, "fb_space_set_halo")
        return h

    def vector_handle(self, ncomp):
        """fb_space handle with `ncomp` interleaved components on the same nodes."""
        if ncomp == 1:
            return self.handle
        if ncomp not in self._vec_handles:
            self._vec_handles[ncomp] = self._create(ncomp)
        return self._vec_handles[ncomp]

    def mass(self):
        if self._mass is None:
            h = _lib.vp()
            _lib.check(lib.fb_assemble_mass(self.handle, C.byref(h)), self.mesh.ctx, "fb_assemble_mass")
            self._mass = h
        return self._mass


# ---------------------------------------------------------------- elements / spaces
class FiniteElement(object):
    def __init__(self, family, cell=None, degree=1):
        assert family in ("Lagrange", "CG", "P"), family
        self.family, self.cell, self._degree, self.ncomp = "Lagrange", cell, degree, 1

    def degree(self):
        return self._degree

    def __mul__(self, other):
        return MixedElement([self, other])


class VectorElement(FiniteElement):
    def __init__(self, family, cell=None, degree=1, dim=None):
        FiniteElement.__init__(self, family, cell, degree)
        self.ncomp = dim if dim is not None else (2 if cell in (triangle, None) else 3)


class MixedElement(object):
    def __init__(self, elements):
        self.elements = list(elements)


class FunctionSpace(object):
    """Lagrange space: scalar (ncomp 1), vector (ncomp == gdim) or mixed (list of sub-spaces)."""

    def __init__(self, mesh, family, degree=None, _ncomp=1):
        self._mesh = mesh
        self.parent, self.component, self.offset = None, None, 0
        if isinstance(family, MixedElement):
            self.subspaces = [FunctionSpace(mesh, e) for e in family.elements]
            off = 0
            for i, s in enumerate(self.subspaces):
                s.parent, s.index, s.offset = self, i, off
                off += s.dim()
            self._dim = off
            self.nodes, self.ncomp = None, None
            return
        self.subspaces = None
        if isinstance(family, FiniteElement):
            degree = family.degree()
            _ncomp = family.ncomp if not isinstance(family, VectorElement) else mesh.dim
        self.nodes = mesh.node_space(degree)
        self.ncomp = _ncomp
        self._dim = self.nodes.nnodes * _ncomp

    def mesh(self):
        return self._mesh

    def dim(self):
        return self._dim

    def degree(self):
        return self.nodes.degree

    def ufl_element(self):
        return self

    def is_mixed(self):
        return self.subspaces is not None

    def sub(self, i):
        if self.subspaces is not None:
            return self.subspaces[i]
        s = FunctionSpace.__new__(FunctionSpace)
        s.__dict__.update(self.__dict__)
        s.parent, s.component = self, i
        return s

    def collapse(self):
        if self.component is not None:
            return FunctionSpace(self._mesh, "Lagrange", self.nodes.degree)
        c = FunctionSpace.__new__(FunctionSpace)
        c.__dict__.update(self.__dict__)
        c.parent, c.offset = None, 0
        return c

    def handle(self):
        return self.nodes.vector_handle(self.ncomp)

    def tabulate_dof_coordinates(self):
        return np.repeat(self.nodes.coords, self.ncomp, axis=0)


def VectorFunctionSpace(mesh, family, degree, dim=None):
    return FunctionSpace(mesh, family, degree, _ncomp=dim or mesh.dim)


# ---------------------------------------------------------------- vectors / functions
class Vector(object):
    """Minimal GenericVector: a view on a numpy array."""

    def __init__(self, array):
        self.a = array

    def __getitem__(self, k):
        # DOLFIN's vector()[:] / get_local() hand out copies; a view would alias the (pooled, pinned) state buffer
        r = self.a[k]
        return r.copy() if isinstance(r, np.ndarray) else r

    def __setitem__(self, k, v):
        self.a[k] = v.a if isinstance(v, Vector) else v

    def __len__(self):
        return self.a.size

    def get_local(self):
        return self.a.copy()

    array = get_local

    def set_local(self, v):
        self.a[:] = v

    def copy(self):
        return Vector(self.a.copy())

    def norm(self, kind="l2"):
        if kind == "linf":
            return float(np.abs(self.a).max()) if self.a.size else 0.0
        if kind == "l1":
            return float(np.abs(self.a).sum())
        return float(np.sqrt(self.a @ self.a))

    def inner(self, other):
        return float(self.a @ other.a)

    def _arr(self, o):
        return o.a if isinstance(o, Vector) else o

    def __add__(self, o):
        return Vector(self.a + self._arr(o))

    __radd__ = __add__

    def __sub__(self, o):
        return Vector(self.a - self._arr(o))

    def __rsub__(self, o):
        return Vector(self._arr(o) - self.a)

    def __mul__(self, o):
        return Vector(self.a * self._arr(o))

    __rmul__ = __mul__

    def __neg__(self):
        return Vector(-self.a)

    def __iadd__(self, o):
        self.a += self._arr(o)
        return self

    def __isub__(self, o):
        self.a -= self._arr(o)
        return self

    def __imul__(self, o):
        self.a *= self._arr(o)
        return self

    def __bool__(self):
        return True


class Function(object):
    def __init__(self, V, values=None):
        self.V = V
        self._vec = np.zeros(V.dim()) if values is None else np.ascontiguousarray(values, dtype=np.float64)
        assert self._vec.shape == (V.dim(),)
        self._name = "f"

    def function_space(self):
        return self.V

    def vector(self):
        return Vector(self._vec)

    def assign(self, other):
        self._vec[:] = other._vec if isinstance(other, Function) else _const_values(other, self.V)

    def copy(self, deepcopy=True):
        return Function(self.V, self._vec.copy())

    def rename(self, name, label):
        self._name = name

    def name(self):
        return self._name

    def ufl_element(self):
        return self.V

    def split(self, deepcopy=False):
        V = self.V
        if V.is_mixed():
            return tuple(Function(s.collapse(), self._vec[s.offset:s.offset + s.dim()].copy()) for s in V.subspaces)
        S = FunctionSpace(V.mesh(), "Lagrange", V.nodes.degree)
        return tuple(Function(S, self._vec[i::V.ncomp].copy()) for i in range(V.ncomp))

    def sub(self, i):
        return self.split()[i]

    def __call__(self, *x):
        raise NotImplementedError("point evaluation is not part of the hot path")

    def nodal(self):
        """(nnodes, ncomp) copy of the nodal values."""
        return self._vec.reshape(self.V.nodes.nnodes, self.V.ncomp).copy()

    def nodal_view(self):
        """(nnodes, ncomp) writable view of the nodal values (valid while this Function is alive)."""
        return self._vec.reshape(self.V.nodes.nnodes, self.V.ncomp)


class Constant(object):
    def __init__(self, value, cell=None):
        self._v = np.atleast_1d(np.asarray(value, dtype=np.float64)).copy()
        self.scalar = np.ndim(value) == 0

    def values(self):
        return self._v

    def assign(self, value):
        self._v[:] = np.atleast_1d(np.asarray(value.values() if isinstance(value, Constant) else value, dtype=float))

    def __float__(self):
        return float(self._v[0])

    def degree(self):
        return 0

    def __call__(self, X):
        X = np.asarray(X)
        out = np.broadcast_to(self._v, (X.shape[0], self._v.size)).copy()
        return out[:, 0] if self.scalar else out

    def __mul__(self, o):
        if isinstance(o, Expression):  # Constant * coordinate expression: the expression's reflected operator takes over
            return NotImplemented
        return Constant(self._v * float(o)) if not self.scalar else Constant(float(self) * float(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return Constant(float(self) / float(o))

    def __rtruediv__(self, o):
        return Constant(float(o) / float(self))


def _to_float(v):
    return float(v)


class _X(object):
    """`x[i]` inside a C-string expression -> coordinate column i."""

    def __init__(self, X):
        self.X = X

    def __getitem__(self, i):
        return self.X[:, i]


_MATH = {k: getattr(np, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "sinh", "cosh", "tanh", "fabs", "floor", "ceil")}
_MATH.update(pow=np.power, atan=np.arctan, asin=np.arcsin, acos=np.arccos, atan2=np.arctan2, abs=np.abs, pi=math.pi,
             M_PI=math.pi, DOLFIN_EPS=DOLFIN_EPS, fmin=np.minimum, fmax=np.maximum)


class Expression(object):
    """Nodal expression: C-string(s) in `x`, or a Python callable X(npts, gdim) -> values.
    Keyword parameters (t=..., mu=...) are attributes, as in DOLFIN."""

    def __init__(self, code, degree=None, element=None, cell=None, domain=None, **params):
        object.__setattr__(self, "user_parameters", dict(params))
        self.cppcode = code
        self._degree = degree if degree is not None else (element.degree() if element is not None else 2)
        self.cell = cell
        if callable(code):
            self._fn, self.ncomp = code, None
        else:
            codes = (code,) if isinstance(code, str) else tuple(code)
            self._compiled = [compile(c, "<expression>", "eval") for c in codes]
            self.ncomp = len(codes)
            self.scalar = isinstance(code, str)

    def __setattr__(self, k, v):
        if "user_parameters" in self.__dict__ and k in self.user_parameters:
            self.user_parameters[k] = v
        else:
            object.__setattr__(self, k, v)

    def __getattr__(self, k):
        up = self.__dict__.get("user_parameters", {})
        if k in up:
            return up[k]
        raise AttributeError(k)

    def degree(self):
        return self._degree

    def ufl_element(self):
        return FiniteElement("Lagrange", self.cell, self._degree)

    def __call__(self, X):
        X = np.atleast_2d(np.asarray(X, dtype=float))
        if callable(self.cppcode):
            return np.asarray(self._fn(X, **self.user_parameters) if self.user_parameters else self._fn(X))
        ns = dict(_MATH)
        ns.update(self.user_parameters)
        ns["x"] = _X(X)
        cols = [np.broadcast_to(np.asarray(eval(c, {"__builtins__": {}}, ns), dtype=float), (X.shape[0],)) for c in self._compiled]
        return cols[0].copy() if self.scalar else np.stack(cols, axis=1)


class SubDomain(object):
    def inside(self, x, on_boundary):
        raise NotImplementedError


def _const_values(value, V):
    """Values of a constant-like object broadcast over the dofs of V."""
    if isinstance(value, Constant):
        v = value.values()
    else:
        v = np.atleast_1d(np.asarray(value, dtype=float))
    n = V.nodes.nnodes
    if V.ncomp == 1 or V.component is not None:
        return np.full(n, v[0])
    return np.tile(v, n)


class DirichletBC(object):
    """DirichletBC(V, value, where): V may be a full space, a component `W.sub(i)` or a
    sub-space of a mixed space; where = 'on_boundary', a SubDomain, or f(x, on_boundary)."""

    def __init__(self, V, value, where="on_boundary", method="topological"):
        self.V, self.value, self.where = V, value, where
        self._nodes = None

    def function_space(self):
        return self.V

    def nodes(self):
        if self._nodes is None:
            ns = self.V.nodes
            cand = np.nonzero(ns.on_boundary)[0]
            if isinstance(self.where, str):
                assert self.where == "on_boundary", self.where
                sel = cand
            else:
                fn = self.where.inside if isinstance(self.where, SubDomain) else self.where
                X = ns.coords[cand]
                try:
                    keep = np.asarray(fn(X.T, True))
                    if keep.shape != (cand.size,):
                        raise ValueError
                except Exception:
                    keep = np.array([bool(fn(x, True)) for x in X], dtype=bool)
                sel = cand[keep.astype(bool)]
            self._nodes = sel
        return self._nodes

    def dofs_values(self):
        """(dofs, values) in the numbering of the *collapsed* space this BC constrains."""
        V = self.V
        nodes = self.nodes()
        X = V.nodes.coords[nodes]
        if V.component is not None:  # one component of a vector space
            nc = V.ncomp
            dofs = nodes * nc + V.component
            vals = self._eval(X, 1).reshape(-1)
            return dofs.astype(np.int64), np.ascontiguousarray(vals, dtype=np.float64)
        nc = V.ncomp
        vals = self._eval(X, nc).reshape(nodes.size, nc)
        dofs = (nodes[:, None] * nc + np.arange(nc)[None, :]).reshape(-1)
        return dofs.astype(np.int64), np.ascontiguousarray(vals.reshape(-1), dtype=np.float64)

    def _eval(self, X, nc):
        v = self.value
        if isinstance(v, Function):
            return v.nodal()[self.nodes()]
        if isinstance(v, (Expression, Constant)):
            return np.asarray(v(X), dtype=float)
        arr = np.atleast_1d(np.asarray(v, dtype=float))
        return np.broadcast_to(arr, (X.shape[0], arr.size)).copy()


def collect_bcs(bcs, V):
    """Merge a list of DirichletBCs acting on (components of) V into sorted unique (dofs, vals);
    later conditions win, like successive bc.apply calls."""
    if not bcs:
        return np.zeros(0, np.int64), np.zeros(0)
    merged = {}
    dl, vl = [], []
    for bc in bcs:
        d, v = bc.dofs_values()
        dl.append(d)
        vl.append(v)
    d = np.concatenate(dl)
    v = np.concatenate(vl)
    # keep the last occurrence of each dof
    _, idx = np.unique(d[::-1], return_index=True)
    idx = d.size - 1 - idx
    order = np.argsort(d[idx])
    del merged
    return np.ascontiguousarray(d[idx][order]), np.ascontiguousarray(v[idx][order])


# ---------------------------------------------------------------- integrals used by the tests
class _Integral(object):
    def __init__(self, integrand, mesh):
        self.integrand, self.mesh = integrand, mesh


class _Measure(object):
    def __init__(self, mesh=None):
        self.mesh = mesh

    def __call__(self, mesh=None, **kw):
        return _Measure(mesh)

    def __rmul__(self, f):
        return _Integral(f, self.mesh)


dx = _Measure()


def assemble(form):
    """Scalar functionals  f*dx(mesh)  with f a number, Constant, Expression or scalar Function."""
    assert isinstance(form, _Integral), "only scalar functionals f*dx are supported"
    f, mesh = form.integrand, form.mesh
    if isinstance(f, Function):
        mesh = f.function_space().mesh()
        ns = f.function_space().nodes
        return hostfem.integrate_nodal(mesh.coordinates(), mesh.cells(), ns.cell_nodes, ns.degree, f._vec)
    assert mesh is not None, "dx(mesh) needed"
    if isinstance(f, Expression):
        return hostfem.integrate_expression(mesh.coordinates(), mesh.cells(), f, f.degree())
    return float(f) * float(mesh.volumes().sum())


class _CoordinateExpr(Expression):
    """Polynomial expression in the spatial coordinates with the arithmetic the reference's drivers use on
    `SpatialCoordinate(mesh)[i]` (`g * y`, tests/test_sealed_box.py:84-88; `rho * g * y`, tests/test_boussinesq.py:152):
    a callable of the points plus its polynomial degree, closed under +, -, * with numbers, `Constant`s and each other."""

    def __init__(self, fn, degree):
        Expression.__init__(self, fn, degree=degree)

    @staticmethod
    def _wrap(o):
        if isinstance(o, _CoordinateExpr):
            return o._fn, o.degree()
        v = float(o)  # numbers and scalar Constants
        return (lambda X: np.full(np.atleast_2d(X).shape[0], v)), 0

    def __mul__(self, o):
        f, g = self._fn, None
        g, dg = self._wrap(o)
        return _CoordinateExpr(lambda X: f(X) * g(X), self.degree() + dg)

    __rmul__ = __mul__

    def __add__(self, o):
        f = self._fn
        g, dg = self._wrap(o)
        return _CoordinateExpr(lambda X: f(X) + g(X), max(self.degree(), dg))

    __radd__ = __add__

    def __neg__(self):
        f = self._fn
        return _CoordinateExpr(lambda X: -f(X), self.degree())

    def __sub__(self, o):
        return self + (-o if isinstance(o, _CoordinateExpr) else -float(o))

    def __rsub__(self, o):
        return (-self) + o

    def __truediv__(self, o):
        return self * (1.0 / float(o))

    def __pow__(self, k):
        f = self._fn
        assert int(k) == k and k >= 0
        return _CoordinateExpr(lambda X: f(X) ** int(k), self.degree() * int(k))


class SpatialCoordinate(object):
    """`SpatialCoordinate(mesh)[i]`: coordinate i as a degree-1 expression (see _CoordinateExpr)."""

    def __init__(self, mesh):
        self._dim = int(np.asarray(mesh.coordinates()).shape[1])

    def __getitem__(self, i):
        if not 0 <= i < self._dim:
            raise IndexError(i)
        return _CoordinateExpr(lambda X, i=i: np.atleast_2d(np.asarray(X, dtype=float))[:, i], 1)


def sqrt(x):
    """Square root of numbers / arrays.  The one finite-element use in the reference's drivers,
    `project(sqrt(ux**2 + uy**2), P2, quadrature_degree 4)` (tests/test_karman_vortex_street.py:262-268,
    tests/test_sealed_box.py:134-140), is one device call here: `flow_b200.drivers.velocity_magnitude(u, P)`."""
    if isinstance(x, (Function, Expression)):
        raise NotImplementedError("sqrt of a finite-element expression: use flow_b200.drivers.velocity_magnitude(u, P)")
    return np.sqrt(x)


def plot(*args, **kwargs):
    """No-op (the reference's drivers call dolfin.plot / interactive only behind `if show:` switches)."""
    return None


def interactive(*args, **kwargs):
    return None


def _callable_of(f):
    if isinstance(f, (Expression, Constant)):
        return f, f.degree()
    if callable(f):
        return f, 2
    c = Constant(f)
    return c, 0


def interpolate(f, V):
    fn, _ = _callable_of(f)
    vals = np.asarray(fn(V.nodes.coords), dtype=float)
    return Function(V, np.ascontiguousarray(vals.reshape(V.nodes.nnodes, -1)[:, :V.ncomp].reshape(-1)))


def project(f, V, tol=1e-14):
    """L2 projection: mass-matrix solve on the GPU (the reference's project() uses LU)."""
    if isinstance(f, Function):
        raise NotImplementedError("project(Function) is not needed by the hot path")
    fn, deg = _callable_of(f)
    mesh, ns = V.mesh(), V.nodes
    b = hostfem.load_vector(mesh.coordinates(), mesh.cells(), ns.cell_nodes, ns.nnodes, ns.degree, V.ncomp, fn, deg)
    x = np.zeros_like(b)
    its = C.c_int()
    _lib.check(lib.fb_mat_solve_cg(ns.mass(), V.ncomp, _lib.as_pd(b), _lib.as_pd(x), 0, None, None, tol, 2000, C.byref(its)),
               mesh.ctx, "project")
    return Function(V, x)


def errornorm(u, uh, norm_type="L2", degree_rise=3):
    assert norm_type.lower() == "l2"
    V = uh.function_space()
    mesh, ns = V.mesh(), V.nodes
    fn, _ = _callable_of(u)
    return hostfem.errornorm_l2(mesh.coordinates(), mesh.cells(), ns.cell_nodes, ns.degree, uh._vec, fn, degree_rise)


def norm(f, norm_type="L2"):
    if isinstance(f, Vector):
        return f.norm(norm_type.lower())
    assert norm_type.lower() == "l2"
    V = f.function_space()
    ns = V.nodes
    x = np.ascontiguousarray(f._vec)
    y = np.zeros_like(x)
    _lib.check(lib.fb_mat_spmv(ns.mass(), V.ncomp, _lib.as_pd(x), _lib.as_pd(y)), V.mesh().ctx, "norm")
    return float(np.sqrt(max(x @ y, 0.0)))


# result output used by the reference's drivers (host-side I/O, flow_b200/io.py)
from .io import File, XDMFFile, mpi_comm_world  # noqa: E402,F401
