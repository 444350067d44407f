"""Bit-exact integer work: meshes, edge numbering, dof maps, boundary sets and CSR sparsity
patterns of the C++ library vs the oracle (runs without a GPU: host-only context)."""
import ctypes as C

import numpy as np
import pytest

from oracle import fem

CASES = {
    "sq_crossed": (lambda d: d.UnitSquareMesh(5, 7, "crossed"), lambda: fem.unit_square_mesh(5, 7, "crossed")),
    "sq_leftright": (lambda d: d.UnitSquareMesh(6, 5, "left/right"), lambda: fem.unit_square_mesh(6, 5, "left/right")),
    "sq_rightleft": (lambda d: d.UnitSquareMesh(3, 4, "right/left"), lambda: fem.unit_square_mesh(3, 4, "right/left")),
    "sq_left": (lambda d: d.UnitSquareMesh(2, 2, "left"), lambda: fem.unit_square_mesh(2, 2, "left")),
    "rect_right": (lambda d: d.RectangleMesh(d.Point(-1, -1), d.Point(1, 2), 4, 3, "right"),
                   lambda: fem.rectangle_mesh((-1, -1), (1, 2), 4, 3, "right")),
    "single_cell_pair": (lambda d: d.UnitSquareMesh(1, 1), lambda: fem.unit_square_mesh(1, 1)),
    "cube": (lambda d: d.UnitCubeMesh(3, 4, 5), lambda: fem.unit_cube_mesh(3, 4, 5)),
    "cube1": (lambda d: d.UnitCubeMesh(1, 1, 1), lambda: fem.unit_cube_mesh(1, 1, 1)),
    "box": (lambda d: d.BoxMesh(d.Point(0, -1, 0.5), d.Point(2, 0, 1), 2, 3, 2), lambda: fem.box_mesh((0, -1, 0.5), (2, 0, 1), 2, 3, 2)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_mesh_dofmap_pattern_bit_exact(name):
    from flow_b200 import _lib, dolfin as d
    from flow_b200._lib import lib

    mk, omk = CASES[name]
    m = mk(d)
    om = fem.Mesh(*omk())
    assert np.array_equal(m.coordinates(), om.points)
    assert np.array_equal(m.cells(), om.cells)
    ne = _lib.i64()
    nb = _lib.i64()
    lib.fb_mesh_info(m.handle, None, None, C.byref(ne), C.byref(nb))
    pe = _lib.pi32()
    lib.fb_mesh_edges(m.handle, C.byref(pe))
    assert np.array_equal(np.ctypeslib.as_array(pe, shape=(ne.value, 2)), om.edges)
    pc, pl = _lib.pi32(), _lib.pi32()
    lib.fb_mesh_boundary_facets(m.handle, C.byref(pc), C.byref(pl))
    assert np.array_equal(np.ctypeslib.as_array(pc, shape=(nb.value,)), om.bfacet_cell)
    assert np.array_equal(np.ctypeslib.as_array(pl, shape=(nb.value,)), om.bfacet_local)
    for deg in (1, 2):
        ns = m.node_space(deg)
        osx = fem.Space(om, deg)
        assert np.array_equal(ns.cell_nodes, osx.cell_nodes)
        assert np.array_equal(ns.coords, osx.node_coords)
        assert np.array_equal(ns.on_boundary, osx.boundary_node)
        nnz = _lib.i64()
        ip, ix = _lib.pi64(), _lib.pi32()
        lib.fb_space_pattern(ns.handle, C.byref(nnz), C.byref(ip), C.byref(ix))
        oip, oix = osx.pattern()
        assert np.array_equal(np.ctypeslib.as_array(ip, shape=(ns.nnodes + 1,)), oip)
        assert np.array_equal(np.ctypeslib.as_array(ix, shape=(nnz.value,)), oix)


def test_survey_sizes():
    """Counts quoted in SURVEY.md section 8 for UnitSquareMesh(32,32,'crossed')."""
    from flow_b200 import _lib, dolfin as d
    from flow_b200._lib import lib

    m = d.UnitSquareMesh(32, 32, "crossed")
    assert m.num_cells() == 4096
    p1, p2 = m.node_space(1), m.node_space(2)
    assert p1.nnodes == 2113 and 2 * p2.nnodes == 16642
    nnz = _lib.i64()
    lib.fb_space_pattern(p1.handle, C.byref(nnz), None, None)
    assert nnz.value == 14529
    lib.fb_space_pattern(p2.handle, C.byref(nnz), None, None)
    assert nnz.value == 94721


def test_unsorted_cells_are_canonicalised():
    from flow_b200 import dolfin as d

    pts, cells = fem.unit_square_mesh(3, 3, "crossed")
    rng = np.random.default_rng(0)
    shuffled = np.stack([rng.permutation(c) for c in cells])
    m = d.Mesh(pts, shuffled)
    assert np.array_equal(m.cells(), cells)


@pytest.mark.parametrize("dim", [2, 3])
def test_coordinate_node_order_is_a_consistent_renumbering(dim):
    """mesh.node_order = "lexicographic" (experimental locality numbering): dof map, coordinates, boundary flags and
    sparsity pattern are those of the canonical space under one permutation."""
    from flow_b200 import _lib
    from flow_b200 import dolfin as d
    from flow_b200._lib import lib

    def pattern(ns):
        nnz, ip, ix = _lib.i64(), _lib.pi64(), _lib.pi32()
        lib.fb_space_pattern(ns.handle, C.byref(nnz), C.byref(ip), C.byref(ix))
        indptr = np.ctypeslib.as_array(ip, shape=(ns.nnodes + 1,)).copy()
        return indptr, np.ctypeslib.as_array(ix, shape=(nnz.value,)).copy()

    make = (lambda: d.UnitSquareMesh(5, 4, "crossed")) if dim == 2 else (lambda: d.UnitCubeMesh(3, 2, 4))
    m0, m1 = make(), make()
    m1.node_order = "lexicographic"
    for degree in (1, 2):
        a, b = m0.node_space(degree), m1.node_space(degree)
        perm = b.perm
        assert a.perm is None and sorted(perm) == list(range(a.nnodes))
        assert np.array_equal(b.cell_nodes, perm[a.cell_nodes])
        assert np.array_equal(b.coords[perm], a.coords) and np.array_equal(b.on_boundary[perm], a.on_boundary)
        keys = [tuple(x[::-1]) for x in b.coords]
        assert keys == sorted(keys)  # numbered by (z, y, x)
        ipa, ixa = pattern(a)
        ipb, ixb = pattern(b)
        for i in range(a.nnodes):
            assert np.array_equal(np.sort(perm[ixa[ipa[i]:ipa[i + 1]]]), ixb[ipb[perm[i]]:ipb[perm[i] + 1]])


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_tile_format_reproduces_csr_pattern(dim, degree):
    """Tile-CSR format of csrc/fb_tile.cu (host build, no device): every owned row exactly once, union-local column
    indices map back to the CSR columns, value slots to the CSR slots, TMA alignment and shared-memory caps hold."""
    import ctypes as C

    import numpy as np

    from flow_b200 import _lib
    from flow_b200 import dolfin as d

    mesh = d.UnitSquareMesh(37, 29, "crossed") if dim == 2 else d.UnitCubeMesh(11, 9, 13)
    V = d.FunctionSpace(mesh, "CG", degree)
    stats = np.zeros(9, dtype=np.int64)
    _lib.check(_lib.lib.fb_space_tile_check(V.handle(), _lib.as_pi64(stats)), mesh.ctx, "fb_space_tile_check")
    nt, nent, usum, max_r, max_e, max_u, cap_r, cap_e, cap_u = [int(x) for x in stats]
    nnz = _lib.i64()
    _lib.lib.fb_space_pattern(V.handle(), C.byref(nnz), None, None)
    assert nt >= 1 and nnz.value <= nent < 1.6 * nnz.value + 32 * nt  # rows padded to the longest of their group of 8
    assert max_r <= cap_r and max_e <= cap_e and max_u <= cap_u
    # locality: a tile's column union stays far below one column per entry (this is what the format buys)
    assert usum < 0.5 * nnz.value or nt == 1
