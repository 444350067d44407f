"""GPU parity for flow.heat.Heat and flow.stokes.solve against the oracle, plus the reference's
Stokes order test (/root/reference/tests/test_stokes.py:102-159) on the facade."""
import ctypes as C

import numpy as np
import pytest

import mms_problems as mp
from common import facade_mesh, mat_to_csr, oracle_mesh, rel
from oracle import fem, forms, heat as oheat, stokes as ostokes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,degree", [("tri_crossed", 2), ("tri_right", 1), ("tet", 2), ("tet_box", 1)])
def test_heat_operator_eval_solve(gpu_ctx, name, degree):
    from flow_b200 import _lib, dolfin as d, heat
    from flow_b200._lib import lib

    om = oracle_mesh(name)
    dim = om.dim
    m = facade_mesh(om)
    Q = d.FunctionSpace(m, "Lagrange", degree)
    W = d.VectorFunctionSpace(m, "CG", 2)
    Wo = fem.Space(om, 2, dim)
    rng = np.random.default_rng(3)
    conv = 0.3 * rng.standard_normal(Wo.ndofs)
    kappa, rho, cp = 0.6, 2.0, 1.5
    Vo = fem.Space(om, degree, 1)
    bd = Vo.boundary_dofs()
    hot = bd[Vo.node_coords[bd, 0] < om.points[:, 0].min() + 1e-12]
    vals = np.full(hot.size, 320.0)
    oh = oheat.Heat(om, degree, conv, kappa, rho, cp, (hot, vals))
    bc = d.DirichletBC(Q, 320.0, lambda x, on: on and x[0] < om.points[:, 0].min() + 1e-12)
    h = heat.Heat(Q, d.Function(W, conv.copy()), kappa, rho, cp, [bc], d.Constant(0.0))
    # operator parity
    mh = _lib.vp()
    lib.fb_heat_matrix(h._h, 0, C.byref(mh))
    A = mat_to_csr(mh, Q.nodes, 1)
    assert abs(A - oh.A).max() / abs(oh.A).max() < 1e-12
    # eval (heat.py:92-101)
    theta0 = 293.0 + rng.standard_normal(Vo.nnodes)
    out = h.eval_alpha_M_beta_F(1.3, -0.4, d.Function(Q, theta0.copy()), 0.0)
    assert rel(out.a, oh.eval_alpha_M_beta_F(1.3, -0.4, theta0)) < 1e-12
    # implicit Euler step (parabolic.ImplicitEuler): (M - dt A) theta1 = M theta0 with Dirichlet rows
    dt = 0.05
    th1 = heat.ImplicitEuler(h).step(d.Function(Q, theta0.copy()), 0.0, dt)
    th1o = oheat.implicit_euler_step(oh, theta0, 0.0, dt)
    assert np.linalg.norm(th1._vec - th1o) / np.linalg.norm(th1o) < 1e-9
    assert np.allclose(th1._vec[hot], 320.0)


@pytest.mark.parametrize("amg", [False, True])
@pytest.mark.parametrize("name", ["tri_leftright", "tet"])
def test_stokes_matches_oracle(gpu_ctx, name, amg, monkeypatch):
    """amg = True forces the AMG-preconditioned velocity-block solves (default from 4096 P2 nodes) on the small mesh."""
    from flow_b200 import dolfin as d, stokes

    monkeypatch.setenv("FB_STOKES_AMG_MIN", "0" if amg else "1000000000")

    om = oracle_mesh(name)
    dim = om.dim
    m = facade_mesh(om)
    cell = m.ufl_cell()
    WP = d.FunctionSpace(m, d.VectorElement("Lagrange", cell, 2) * d.FiniteElement("Lagrange", cell, 1))
    Wo, Po = fem.Space(om, 2, dim), fem.Space(om, 1, 1)
    X = Wo.node_coords
    g = np.stack([np.sin(X[:, (i + 1) % dim]) for i in range(dim)], 1)
    gp = np.cos(Po.node_coords[:, 0])
    ubd, pbd = Wo.boundary_dofs(), Po.boundary_dofs()
    mu = 0.7
    fvec = (0.3, -1.0) if dim == 2 else (0.3, -1.0, 0.2)
    Mv = __import__("scipy.sparse", fromlist=["kron"]).kron(forms.mass_matrix(fem.Space(om, 2, 1)), np.eye(dim))
    load = Mv @ np.tile(fvec, Wo.nnodes)
    uo, po = ostokes.solve(om, mu, load, (ubd, g.reshape(-1)[ubd]), (pbd, gp[pbd]))
    bcs = [d.DirichletBC(WP.sub(0), d.Function(WP.sub(0).collapse(), g.reshape(-1)), "on_boundary"),
           d.DirichletBC(WP.sub(1), d.Function(WP.sub(1).collapse(), gp), "on_boundary")]
    u, p = stokes.solve(WP, bcs, mu, d.Constant(fvec), verbose=False, tol=1e-12, max_iter=500)
    assert np.linalg.norm(u._vec - uo) / np.linalg.norm(uo) < 1e-8
    assert np.linalg.norm(p._vec - po) / np.linalg.norm(po) < 1e-8


def test_stokes_component_bcs_amg(gpu_ctx, monkeypatch):
    """Velocity conditions that differ per component (y-component on the whole boundary, x-component everywhere but on the
    right edge, whose natural condition fixes the pressure level; per-component conditions as in
    tests/test_karman_vortex_street.py:139-150) with the AMG-preconditioned velocity-block solves forced on: one
    hierarchy per constraint set."""
    from flow_b200 import dolfin as d, stokes

    monkeypatch.setenv("FB_STOKES_AMG_MIN", "0")
    om = oracle_mesh("tri_leftright")
    m = facade_mesh(om)
    cell = m.ufl_cell()
    WP = d.FunctionSpace(m, d.VectorElement("Lagrange", cell, 2) * d.FiniteElement("Lagrange", cell, 1))
    Wo = fem.Space(om, 2, 2)
    X = Wo.node_coords
    g = np.stack([np.sin(X[:, 1]), 0.3 * np.cos(X[:, 0])], 1)
    xmax = X[:, 0].max()
    bnodes = np.unique(Wo.boundary_dofs() // 2)
    right = bnodes[X[bnodes, 0] > xmax - 1e-12]
    ubd = np.sort(np.concatenate([2 * bnodes + 1, 2 * np.setdiff1d(bnodes, right)]))
    mu = 0.7
    Mv = __import__("scipy.sparse", fromlist=["kron"]).kron(forms.mass_matrix(fem.Space(om, 2, 1)), np.eye(2))
    load = Mv @ np.tile((0.3, -1.0), Wo.nnodes)
    uo, po = ostokes.solve(om, mu, load, (ubd, g.reshape(-1)[ubd]), None)
    V = WP.sub(0)
    gx = d.Function(V.sub(0).collapse(), g[:, 0].copy())
    gy = d.Function(V.sub(1).collapse(), g[:, 1].copy())
    bcs = [d.DirichletBC(V.sub(1), gy, "on_boundary"),
           d.DirichletBC(V.sub(0), gx, lambda x, on: on and x[0] < xmax - 1e-12)]
    u, p = stokes.solve(WP, bcs, mu, d.Constant((0.3, -1.0)), verbose=False, tol=1e-12, max_iter=500)
    assert np.linalg.norm(u._vec - uo) / np.linalg.norm(uo) < 1e-8
    assert np.linalg.norm(p._vec - po) / np.linalg.norm(po) < 1e-8


def test_stokes_order(gpu_ctx):
    """tests/test_stokes.py:102-117: Guermond1 on UnitSquareMesh(n,n,'left/right'), n = 8, 16;
    observed spatial orders of u and p must exceed 1.9."""
    from flow_b200 import dolfin as d, stokes

    pr = mp.stokes_guermond1()
    hmax, ue, pe = [], [], []
    for n in (8, 16):
        mesh = d.UnitSquareMesh(n, n, pr["mesh"][1])
        u_sol = d.Expression(pr["u"], degree=mp.MAX_DEGREE)
        p_sol = d.Expression(pr["p"], degree=mp.MAX_DEGREE)
        f = d.Expression(pr["f"], degree=mp.MAX_DEGREE)
        Wel = d.VectorElement("Lagrange", mesh.ufl_cell(), 2)
        Pel = d.FiniteElement("Lagrange", mesh.ufl_cell(), 1)
        WP = d.FunctionSpace(mesh, Wel * Pel)
        bcs = [d.DirichletBC(WP.sub(0), u_sol, "on_boundary"), d.DirichletBC(WP.sub(1), p_sol, "on_boundary")]
        u, p = stokes.solve(WP, bcs=bcs, mu=pr["mu"], f=f, verbose=False, tol=1.0e-12)
        hmax.append(mesh.hmax())
        ue.append(d.errornorm(u_sol, u))
        pe.append(d.errornorm(p_sol, p))
    assert mp.compute_numerical_order_of_convergence(hmax, ue)[0] > 1.9
    assert mp.compute_numerical_order_of_convergence(hmax, pe)[0] > 1.9


@pytest.mark.parametrize("degree", [1, 2])
def test_heat_supg(gpu_ctx, degree):
    """supg_stabilization=True (heat.py:60-86 with the tau of stabilization.py): operator, SUPG mass part,
    eval and an implicit Euler step against the oracle."""
    from flow_b200 import _lib, dolfin as d, heat, stabilization
    from flow_b200._lib import lib

    om = oracle_mesh("tri_crossed")
    m = facade_mesh(om)
    Q = d.FunctionSpace(m, "Lagrange", degree)
    W = d.VectorFunctionSpace(m, "CG", 2)
    Wo = fem.Space(om, 2, 2)
    X = Wo.node_coords
    conv = np.stack([2.0 + np.sin(3 * X[:, 1]), -1.0 + X[:, 0] ** 2], 1).reshape(-1)   # convection dominated
    kappa, rho, cp = 1e-3, 1.0, 1.0
    Vo = fem.Space(om, degree, 1)
    bd = Vo.boundary_dofs()
    vals = 300.0 + 20.0 * Vo.node_coords[bd, 1]
    oh = oheat.Heat(om, degree, conv, kappa, rho, cp, (bd, vals), supg=True, source=0.7,
                    source_load=forms.expression_load_vector(Vo, lambda Y: np.full((Y.shape[0], 1), 0.7), 0))
    bc = d.DirichletBC(Q, d.Expression("300.0 + 20.0*x[1]", degree=1), "on_boundary")
    h = heat.Heat(Q, d.Function(W, conv.copy()), kappa, rho, cp, [bc], d.Constant(0.7), supg_stabilization=True)
    mh = _lib.vp()
    lib.fb_heat_matrix(h._h, 0, C.byref(mh))
    A = mat_to_csr(mh, Q.nodes, 1)
    assert abs(A - oh.A).max() / abs(oh.A).max() < 1e-12
    ms = np.zeros(A.nnz)
    _lib.check(lib.fb_heat_supg_mass(h._h, _lib.as_pd(ms)), gpu_ctx, "supg mass")
    Ms = (oh.M - forms.lumped_vertex_mass(Vo)).tocsr()
    import scipy.sparse as sp
    got = sp.csr_matrix((ms, A.indices, A.indptr), shape=A.shape)
    assert abs(got - Ms).max() / abs(Ms).max() < 1e-12
    tau = stabilization.supg(m, d.Function(W, conv.copy()), kappa, degree).vertex_values()
    assert np.allclose(tau, oheat.supg_tau(om, Wo, conv, kappa, degree), rtol=1e-13)
    theta0 = 300.0 + np.random.default_rng(1).standard_normal(Vo.nnodes)
    out = h.eval_alpha_M_beta_F(1.0, -0.3, d.Function(Q, theta0.copy()), 0.0)
    assert rel(out.a, oh.eval_alpha_M_beta_F(1.0, -0.3, theta0)) < 1e-12
    th1 = heat.ImplicitEuler(h).step(d.Function(Q, theta0.copy()), 0.0, 0.02)
    th1o = oheat.implicit_euler_step(oh, theta0, 0.0, 0.02)
    assert np.linalg.norm(th1._vec - th1o) / np.linalg.norm(th1o) < 1e-9
