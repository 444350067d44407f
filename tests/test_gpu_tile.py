"""Tile-CSR SpMM (csrc/fb_tile.cu: TMA-streamed entries, shared-memory staged x) against the row-wise CSR kernels and
scipy on the same matrices, and a full IPCS step that runs every scalar P2 product (mass products, velocity-correction
CG with its Dirichlet mask and fused dots, inner CG of the momentum preconditioner) through the tile kernel."""
import ctypes as C
import os

import numpy as np
import pytest

from common import mat_to_csr

pytestmark = pytest.mark.gpu


def _format(mat):
    from flow_b200._lib import lib
    import flow_b200._lib as _lib

    f, nt, ne, nu = C.c_int(), _lib.i64(), _lib.i64(), _lib.i64()
    lib.fb_mat_format_info(mat, C.byref(f), C.byref(nt), C.byref(ne), C.byref(nu))
    return f.value, nt.value, ne.value, nu.value


@pytest.mark.parametrize("dim,degree,kind", [(2, 2, "mass"), (3, 2, "mass"), (3, 2, "stiffness"), (3, 1, "stiffness"), (2, 1, "mass")])
@pytest.mark.parametrize("ncomp", [1, 2, 3])
def test_tile_spmm_matches_csr_and_scipy(gpu_ctx, dim, degree, kind, ncomp):
    from flow_b200 import _lib
    from flow_b200 import dolfin as d
    from flow_b200._lib import lib

    mesh = d.UnitSquareMesh(61, 47, "crossed") if dim == 2 else d.UnitCubeMesh(13, 11, 17)
    V = d.FunctionSpace(mesh, "CG", degree)
    h = _lib.vp()
    fn = lib.fb_assemble_mass if kind == "mass" else lib.fb_assemble_stiffness
    _lib.check(fn(V.handle(), C.byref(h)), mesh.ctx, "assemble")
    try:
        A = mat_to_csr(h, V.nodes, 1)
        n = A.shape[0]
        rng = np.random.default_rng(7)
        x = rng.standard_normal(n * ncomp)
        y_ref = (A @ x.reshape(n, ncomp)).reshape(-1)
        y_csr, y_tile = np.zeros_like(x), np.zeros_like(x)
        _lib.check(lib.fb_mat_spmv(h, ncomp, _lib.as_pd(x), _lib.as_pd(y_csr)), mesh.ctx, "spmv csr")
        assert _format(h)[0] == _lib.FORMAT_CSR
        _lib.check(lib.fb_mat_set_format(h, _lib.FORMAT_TILE), mesh.ctx, "set_format")
        fmt, nt, ne, nu = _format(h)
        assert fmt == _lib.FORMAT_TILE and nt >= 1 and ne >= A.nnz
        _lib.check(lib.fb_mat_spmv(h, ncomp, _lib.as_pd(x), _lib.as_pd(y_tile)), mesh.ctx, "spmv tile")
        scale = np.abs(y_ref).max()
        assert np.abs(y_csr - y_ref).max() < 1e-13 * scale
        assert np.abs(y_tile - y_ref).max() < 1e-13 * scale
        # repeated launches reuse the pipeline state correctly (mbarrier phases, stages)
        y2 = np.zeros_like(x)
        _lib.check(lib.fb_mat_spmv(h, ncomp, _lib.as_pd(x), _lib.as_pd(y2)), mesh.ctx, "spmv tile again")
        assert np.array_equal(y2, y_tile)
        _lib.check(lib.fb_mat_set_format(h, _lib.FORMAT_CSR), mesh.ctx, "set_format back")
        assert _format(h)[0] == _lib.FORMAT_CSR
    finally:
        lib.fb_mat_destroy(h)


def test_tile_cg_solve_with_dirichlet_mask(gpu_ctx):
    """fb_mat_solve_cg (masked Jacobi-PCG: fused dot products and the identity-row mask in the SpMM epilogue) gives the
    same solution in both formats."""
    from flow_b200 import _lib
    from flow_b200 import dolfin as d
    from flow_b200._lib import lib

    mesh = d.UnitCubeMesh(12, 12, 12)
    V = d.FunctionSpace(mesh, "CG", 2)
    h = _lib.vp()
    _lib.check(lib.fb_assemble_mass(V.handle(), C.byref(h)), mesh.ctx, "mass")
    try:
        n, nc = V.dim(), 3
        rng = np.random.default_rng(3)
        b = rng.standard_normal(n * nc) * 1e-3
        X = V.nodes.coords
        bnodes = np.nonzero((np.abs(X - 0.5).max(axis=1) > 0.5 - 1e-12))[0]
        dofs = np.ascontiguousarray((bnodes[:, None] * nc + np.array([0, 2])[None, :]).reshape(-1), dtype=np.int64)  # comps 0 and 2 only
        vals = rng.standard_normal(dofs.size)
        sols = {}
        for fmt in (_lib.FORMAT_CSR, _lib.FORMAT_TILE):
            _lib.check(lib.fb_mat_set_format(h, fmt), mesh.ctx, "set_format")
            x = np.zeros(n * nc)
            its = C.c_int()
            _lib.check(lib.fb_mat_solve_cg(h, nc, _lib.as_pd(b), _lib.as_pd(x), dofs.size, _lib.as_pi64(dofs), _lib.as_pd(vals),
                                           1e-12, 500, C.byref(its)), mesh.ctx, "solve_cg")
            sols[fmt] = (x, its.value)
        xa, ia = sols[_lib.FORMAT_CSR]
        xb, ib = sols[_lib.FORMAT_TILE]
        assert abs(ia - ib) <= 1
        assert np.allclose(xa[dofs], vals) and np.allclose(xb[dofs], vals)
        assert np.linalg.norm(xa - xb) < 1e-10 * np.linalg.norm(xa)
    finally:
        lib.fb_mat_destroy(h)


def test_ipcs_steps_through_tile_kernels_match_oracle(gpu_ctx, monkeypatch):
    """Three IPCS steps on a mesh below the automatic threshold with FB_TILE_MIN_ROWS=0: all scalar P2 products of the
    step run from the tile format; result vs the oracle as in the default-options parity tests."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav
    from flow_b200._lib import lib
    import flow_b200._lib as _lib
    from flow_b200.navier_stokes.pressure_correction import _engine
    from oracle import fem, navier_stokes as ons

    monkeypatch.setenv("FB_TILE_MIN_ROWS", "0")
    nav.reset_options()
    n, dt, rho, mu = 7, 1e-2, 1.0, 1e-2
    om = fem.Mesh(*fem.unit_cube_mesh(n, n, n))
    ost = ons.IPCS(om)
    bd = ost.W.boundary_dofs()
    g = np.zeros((ost.W.nnodes, 3))
    g[ost.W.node_coords[:, 2] > 1 - 1e-12, 0] = 1.0
    g = g.reshape(-1)
    mesh = d.UnitCubeMesh(n, n, n)
    W = d.VectorFunctionSpace(mesh, "CG", 2)
    P = d.FunctionSpace(mesh, "CG", 1)
    ns = _engine(W, P)
    hm = _lib.vp()
    lib.fb_ns_matrix(ns, 1, C.byref(hm))
    assert _format(hm)[0] == _lib.FORMAT_TILE
    bcs = [d.DirichletBC(W, (0.0, 0.0, 0.0), "on_boundary"), d.DirichletBC(W, (1.0, 0.0, 0.0), lambda x, on: x[2] > 1 - 1e-12)]
    zero = d.Constant((0.0, 0.0, 0.0))
    uo, po = np.zeros(ost.W.ndofs), np.zeros(ost.P.nnodes)
    u, p = d.Function(W), d.Function(P)
    for k in range(3):
        uo, po = ost.step(dt, uo, po, (bd, g[bd]), None, rho, mu, None, None, tol=1e-10)
        u, p = nav.IPCS().step(d.Constant(dt), {0: u}, p, bcs, [], d.Constant(rho), d.Constant(mu), {0: zero, 1: zero},
                               verbose=False, tol=1e-10)
        eu = np.linalg.norm(u._vec - uo) / np.linalg.norm(uo)
        dp = (p._vec - p._vec.mean()) - (po - po.mean())
        ep = np.linalg.norm(dp) / np.linalg.norm(po - po.mean())
        assert eu < 1e-8 and ep < 1e-7, (k, eu, ep)
