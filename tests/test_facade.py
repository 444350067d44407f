"""Host-side facade logic: Expression parsing, Dirichlet sets, sub-spaces, the numpy data-entry
helpers (vs the oracle's independent implementations)."""
import numpy as np
import pytest

from oracle import fem, forms, util


def test_expression_cstring_and_parameters():
    from flow_b200 import dolfin as d

    e = d.Expression(("sin(x[0] + t)*pow(x[1], 2)", "mu*cos(pi*x[0])"), degree=5, t=0.0, mu=2.0)
    X = np.array([[0.1, 0.2], [0.3, 0.4]])
    assert np.allclose(e(X), np.stack([np.sin(X[:, 0]) * X[:, 1] ** 2, 2 * np.cos(np.pi * X[:, 0])], 1))
    e.t = 0.5
    assert np.allclose(e(X)[:, 0], np.sin(X[:, 0] + 0.5) * X[:, 1] ** 2)
    assert e.user_parameters["t"] == 0.5 and e.degree() == 5
    s = d.Expression("x[0]*x[1]", degree=2)
    assert s(X).shape == (2,)


def test_dirichlet_sets_and_merge():
    from flow_b200 import dolfin as d

    m = d.UnitSquareMesh(4, 4)
    W = d.VectorFunctionSpace(m, "CG", 2)
    om = fem.Mesh(*fem.unit_square_mesh(4, 4))
    Wo = fem.Space(om, 2, 2)
    dofs, vals = d.collect_bcs([d.DirichletBC(W, (1.0, 2.0), "on_boundary")], W)
    assert np.array_equal(dofs, Wo.boundary_dofs())
    assert np.allclose(vals.reshape(-1, 2), [1.0, 2.0])
    left = d.DirichletBC(W.sub(0), d.Expression("x[1]", degree=1), lambda x, on: on and x[0] < 1e-12)

    class Top(d.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary and x[1] > 1 - 1e-12

    top = d.DirichletBC(W, d.Constant((3.0, 4.0)), Top())
    dofs, vals = d.collect_bcs([left, top], W)
    assert len(np.unique(dofs)) == dofs.size and (np.diff(dofs) > 0).all()
    X = W.tabulate_dof_coordinates()
    for k, v in zip(dofs, vals):
        if X[k, 1] > 1 - 1e-12:
            assert v == (3.0 if k % 2 == 0 else 4.0)  # later BC wins at the corner
        else:
            assert k % 2 == 0 and X[k, 0] < 1e-12 and v == X[k, 1]


def test_mixed_space_and_split():
    from flow_b200 import dolfin as d

    m = d.UnitSquareMesh(3, 3)
    WP = d.FunctionSpace(m, d.VectorElement("Lagrange", m.ufl_cell(), 2) * d.FiniteElement("Lagrange", m.ufl_cell(), 1))
    assert WP.dim() == WP.sub(0).dim() + WP.sub(1).dim() == 2 * 49 + 16
    up = d.Function(WP)
    up.vector()[:] = np.arange(WP.dim())
    u, p = up.split(True)
    assert u.function_space().ncomp == 2 and p.function_space().dim() == 16
    ux, uy = u.split()
    assert np.array_equal(ux.vector().get_local(), u.vector().get_local()[0::2])


@pytest.mark.parametrize("dim", [2, 3])
def test_hostfem_matches_oracle(dim):
    from flow_b200 import dolfin as d, hostfem

    if dim == 2:
        m, om = d.UnitSquareMesh(3, 4, "crossed"), fem.Mesh(*fem.unit_square_mesh(3, 4, "crossed"))
        f = lambda X: np.stack([np.sin(X[:, 0] + 2 * X[:, 1]), np.cos(X[:, 0] * X[:, 1])], 1)
    else:
        m, om = d.UnitCubeMesh(2, 2, 1), fem.Mesh(*fem.unit_cube_mesh(2, 2, 1))
        f = lambda X: np.stack([np.sin(X[:, 0] + 2 * X[:, 1]), np.cos(X[:, 0] * X[:, 2]), X[:, 1] ** 3], 1)
    ns = m.node_space(2)
    Wo = fem.Space(om, 2, dim)
    for deg in (0, 1, 2, 5):
        a = hostfem.load_vector(m.coordinates(), m.cells(), ns.cell_nodes, ns.nnodes, 2, dim, f, deg)
        b = forms.expression_load_vector(Wo, f, deg)
        assert np.allclose(a, b, rtol=1e-12, atol=1e-15)
    uh = np.random.default_rng(0).standard_normal(Wo.ndofs)
    assert abs(hostfem.errornorm_l2(m.coordinates(), m.cells(), ns.cell_nodes, 2, uh, f) - util.errornorm(Wo, f, uh)) < 1e-12
    lam, w = hostfem.quadrature(dim, 6)
    lo, wo = fem.simplex_quadrature(dim, 6)
    assert np.allclose(np.sort(w), np.sort(wo))


MSH = """$MeshFormat
2.2 0 8
$EndMeshFormat
$Nodes
5
1 0 0 0
2 1 0 0
3 1 1 0
4 0 1 0
7 0.5 0.5 0
$EndNodes
$Elements
6
1 15 2 0 1 1
2 1 2 0 1 1 2
3 2 2 0 5 1 2 7
4 2 2 0 5 2 3 7
5 2 2 0 5 3 4 7
6 2 2 0 5 4 1 7
$EndElements
"""


def test_msh_reader(tmp_path):
    from flow_b200 import dolfin as d

    m = d.MshMesh(MSH)
    assert m.num_vertices() == 5 and m.num_cells() == 4 and m.dim == 2
    assert abs(m.volumes().sum() - 1.0) < 1e-15
    f = tmp_path / "unit.msh"
    f.write_text(MSH)
    m2 = d.MshMesh(str(f))
    assert np.array_equal(m.cells(), m2.cells())
    assert m.node_space(2).nnodes == 5 + 8 and m.node_space(1).on_boundary.sum() == 4


def test_pinned_block_outlives_views_of_a_dropped_function(monkeypatch):
    """views derived from a pooled pinned array (numpy collapses their .base to the buffer owner) must keep
    the block out of the pool; vector()[:] and nodal() hand out copies like DOLFIN."""
    import gc

    from flow_b200 import _lib

    class FakeBlock(object):  # stands in for _PinnedBlock on a machine without a GPU
        def __init__(self, ctx, nbytes):
            self._mem = np.zeros(nbytes, dtype=np.uint8)
            self.ptr, self.nbytes = None, nbytes
            self.__array_interface__ = self._mem.__array_interface__

    monkeypatch.setattr(_lib, "_PinnedBlock", FakeBlock)
    monkeypatch.setattr(_lib, "_pinned_pool", {})
    a = _lib.pinned_empty(None, 100)
    a[:] = 1.0
    view = a.reshape(50, 2)[:, 0]
    del a
    gc.collect()
    assert _lib._pinned_pool[800] == []      # the view keeps the block alive
    b = _lib.pinned_empty(None, 100)         # must be a different block
    b[:] = 2.0
    assert (view == 1.0).all()
    del view
    gc.collect()
    assert len(_lib._pinned_pool[800]) == 1  # now it is back in the pool


def test_spatial_coordinate_expressions_and_display_stubs():
    """`SpatialCoordinate(mesh)[1]` with the arithmetic of the reference's drivers (`g * y`, tests/test_sealed_box.py:84-88;
    `rho * g * y`, tests/test_boussinesq.py:152), and the display calls they make (`plot`, `interactive`)."""
    from flow_b200 import dolfin as d

    mesh = d.UnitSquareMesh(3, 4)
    x, y = d.SpatialCoordinate(mesh)[0], d.SpatialCoordinate(mesh)[1]
    g, rho = -9.81, 998.2
    X = np.array([[0.5, 0.25], [1.0, 2.0], [0.0, 0.0]])
    e = d.Constant(g) * y * rho + 3.0 - x ** 2
    assert np.allclose(e(X), g * rho * X[:, 1] + 3.0 - X[:, 0] ** 2)
    assert e.degree() == 2 and (g * y).degree() == 1 and (y / 2.0 - 1.0).degree() == 1
    assert np.allclose((1.0 - y)(X), 1.0 - X[:, 1]) and np.allclose((-y)(X), -X[:, 1])
    with pytest.raises(IndexError):
        d.SpatialCoordinate(mesh)[2]
    # a degree-1 expression interpolates exactly into P1 (what project(g * y, P1) returns in the reference's driver)
    P = d.FunctionSpace(mesh, "CG", 1)
    p0 = d.interpolate(g * y, P)
    assert np.allclose(p0.vector().get_local(), g * P.nodes.coords[:, 1])
    # Constant arithmetic with plain numbers is unchanged
    assert (d.Constant(2.0) * 3.0).values()[0] == 6.0 and (2 * d.Constant(2.0)).values()[0] == 4.0
    assert d.sqrt(4.0) == 2.0 and d.plot(p0) is None and d.interactive() is None
    with pytest.raises(NotImplementedError):
        d.sqrt(p0)
