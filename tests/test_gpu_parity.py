"""GPU parity tests: CUDA path (through the C ABI) vs the oracle on the same seeded inputs.

Bit-exact for integer work (dof maps, patterns: checked on CPU in test_topology.py, here the
device pattern is re-used through fb_mat getters); fp64 assembly agrees to rounding
(atomics reorder sums): tolerance 1e-12 relative to the largest entry.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from common import MESHES, facade_mesh, mat_to_csr, oracle_mesh, rand_state, rel
from oracle import fem, forms, navier_stokes as ons, solvers

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("degree", [1, 2])
def test_constant_operators(gpu_ctx, name, degree):
    from flow_b200 import _lib
    from flow_b200._lib import lib

    om = oracle_mesh(name)
    m = facade_mesh(om)
    ns = m.node_space(degree)
    osx = fem.Space(om, degree, 1)
    for kind, fn, ofn in (("mass", lib.fb_assemble_mass, forms.mass_matrix), ("stiff", lib.fb_assemble_stiffness, forms.stiffness_matrix)):
        h = _lib.vp()
        _lib.check(fn(ns.handle, C.byref(h)), gpu_ctx, kind)
        A = mat_to_csr(h, ns, 1)
        Ao = ofn(osx)
        assert np.array_equal(A.indptr, Ao.indptr) and np.array_equal(A.indices, Ao.indices)
        assert rel(A.data, Ao.data) < TOL, kind
        # SpMV, 1..3 interleaved components
        rng = np.random.default_rng(1)
        for nc in (1, 2, 3):
            x = rng.standard_normal(ns.nnodes * nc)
            y = np.zeros_like(x)
            _lib.check(lib.fb_mat_spmv(h, nc, _lib.as_pd(x), _lib.as_pd(y)), gpu_ctx, "spmv")
            yo = (Ao @ x.reshape(-1, nc)).reshape(-1)
            assert rel(y, yo) < TOL
        lib.fb_mat_destroy(h)
    d = np.zeros(ns.nnodes)
    _lib.check(lib.fb_assemble_lumped_mass(ns.handle, _lib.as_pd(d)), gpu_ctx, "lumped")
    assert rel(d, forms.lumped_vertex_mass(osx).diagonal()) < TOL


def _engine(m):
    from flow_b200 import dolfin as d
    from flow_b200.navier_stokes.pressure_correction import _engine

    W = d.VectorFunctionSpace(m, "CG", 2)
    P = d.FunctionSpace(m, "CG", 1)
    return W, P, _engine(W, P)


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("theta", [1.0, 0.5, 0.0])
def test_momentum_residual_jacobian(gpu_ctx, name, theta):
    from flow_b200 import _lib
    from flow_b200._lib import lib

    om = oracle_mesh(name)
    Wo, Po, ui, u0, p0, _ = rand_state(om)
    m = facade_mesh(om)
    W, P, ns = _engine(m)
    dt, rho, mu = 0.37, 1.3, 0.71
    load = np.random.default_rng(5).standard_normal(Wo.ndofs)
    F = np.zeros(Wo.ndofs)
    _lib.check(lib.fb_ns_residual(ns, dt, rho, mu, theta, _lib.as_pd(ui), _lib.as_pd(u0), _lib.as_pd(p0), _lib.as_pd(load),
                                  _lib.as_pd(F), 1), gpu_ctx, "fb_ns_residual")
    Fo, Jo = forms.momentum_residual_jacobian(Wo, Po, ui, u0, p0, load, dt, rho, mu, theta)
    assert rel(F, Fo) < TOL
    h = _lib.vp()
    lib.fb_ns_matrix(ns, 2, C.byref(h))
    J = mat_to_csr(h, W.nodes, om.dim)
    # the device pattern is the full node-block pattern; the oracle drops nothing either
    D = (J - Jo).tocsr()
    assert abs(D).max() / abs(Jo).max() < TOL
    # blocked SpMV
    x = np.random.default_rng(2).standard_normal(Wo.ndofs)
    y = np.zeros_like(x)
    _lib.check(lib.fb_mat_spmv(h, 1, _lib.as_pd(x), _lib.as_pd(y)), gpu_ctx, "bspmv")
    assert rel(y, Jo @ x) < TOL


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("rotational", [0, 1])
def test_pressure_and_correction_rhs(gpu_ctx, name, rotational):
    from flow_b200 import _lib
    from flow_b200._lib import lib

    om = oracle_mesh(name)
    Wo, Po, ui, _, p0, p1 = rand_state(om, 3)
    m = facade_mesh(om)
    W, P, ns = _engine(m)
    dt, rho, mu = 0.21, 0.9, 1.7
    bp = np.zeros(Po.nnodes)
    _lib.check(lib.fb_ns_pressure_rhs(ns, dt, rho, mu, rotational, _lib.as_pd(ui), _lib.as_pd(p0), _lib.as_pd(bp)), gpu_ctx, "prhs")
    assert rel(bp, forms.pressure_rhs(Wo, Po, ui, p0, dt, rho, mu, bool(rotational))) < TOL
    bu = np.zeros(Wo.ndofs)
    _lib.check(lib.fb_ns_correction_rhs(ns, dt, rho, mu, rotational, _lib.as_pd(ui), _lib.as_pd(p1), _lib.as_pd(p0),
                                        _lib.as_pd(bu)), gpu_ctx, "crhs")
    assert rel(bu, forms.correction_rhs(Wo, Po, ui, p1, p0, dt, rho, mu, bool(rotational))) < TOL


@pytest.mark.parametrize("name", ["tri_crossed", "tet"])
def test_masked_cg_matches_symmetric_elimination(gpu_ctx, name):
    from flow_b200 import _lib
    from flow_b200._lib import lib

    om = oracle_mesh(name)
    m = facade_mesh(om)
    ns = m.node_space(2)
    osx = fem.Space(om, 2, om.dim)
    M = sp.kron(forms.mass_matrix(fem.Space(om, 2, 1)), sp.eye(om.dim), format="csr")
    rng = np.random.default_rng(7)
    b = rng.standard_normal(osx.ndofs)
    dofs = osx.boundary_dofs(comp=0).astype(np.int64)  # constrain one component only
    vals = rng.standard_normal(dofs.size)
    x = np.zeros_like(b)
    its = C.c_int()
    _lib.check(lib.fb_mat_solve_cg(ns.mass(), om.dim, _lib.as_pd(b), _lib.as_pd(x), dofs.size, _lib.as_pi64(dofs),
                                   _lib.as_pd(vals), 1e-13, 500, C.byref(its)), gpu_ctx, "cg")
    A, bb = forms.apply_bc_symmetric(M, b, dofs, vals)
    xo = solvers.lu_solve(A, bb)
    assert rel(x, xo) < 1e-10
    assert 0 < its.value < 100


CASES = [
    ("tri_crossed", "ipcs", "backward euler", 0.05),
    ("tri_leftright", "rotational", "crank-nicolson", 0.05),
    ("tri_right", "chorin", "backward euler", 0.05),
    ("tri_crossed", "ipcs", "forward euler", 2.0e-4),  # explicit: dt below the viscous stability limit
    ("tet", "ipcs", "backward euler", 0.05),
    ("tet_box", "rotational", "backward euler", 0.05),
]


@pytest.mark.parametrize("name,method,scheme,dt", CASES)
def test_step_matches_oracle(gpu_ctx, name, method, scheme, dt):
    """Three consecutive steps, nodal forcing given as a load vector, inhomogeneous Dirichlet data.
    Tolerance: 1e-8 relative L2 (BASELINE.json north_star), pressure compared modulo its mean."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    om = oracle_mesh(name)
    dim = om.dim
    m = facade_mesh(om)
    W = d.VectorFunctionSpace(m, "CG", 2)
    P = d.FunctionSpace(m, "CG", 1)
    if method == "chorin":
        ost, st = ons.Chorin(om), nav.Chorin()
    elif method == "ipcs":
        ost, st = ons.IPCS(om, scheme), nav.IPCS(scheme)
    else:
        ost, st = ons.Rotational(om, scheme), nav.Rotational(scheme)
    Wo, Po = ost.W, ost.P
    X = Wo.node_coords
    lo, hi = X.min(axis=0), X.max(axis=0)

    def field(X, t):  # smooth, time dependent; used for the initial state and the body force
        c = [np.sin(X[:, (i + 1) % dim] + t + 0.3 * i) * np.cos(X[:, i] - t) for i in range(dim)]
        return np.stack(c, 1)

    def lid(t):
        """Dirichlet data with zero normal component everywhere (tangential, time-dependent lid),
        so that the pure-Neumann pressure problem is consistent (pressure_correction.py:347-364)."""
        g = np.zeros((Wo.nnodes, dim))
        top = X[:, -1] > hi[-1] - 1e-12
        s = (X[:, 0] - lo[0]) * (hi[0] - X[:, 0]) * 4.0 / (hi[0] - lo[0]) ** 2
        g[top, 0] = (1.0 + t) * s[top]
        return g.reshape(-1)

    rho, mu = 1.2, 0.3
    bd = Wo.boundary_dofs()
    u0 = 0.2 * field(X, 0.0).reshape(-1)
    u0[bd] = lid(0.0)[bd]
    p0 = np.sin(Po.node_coords[:, 0] - Po.node_coords[:, 1])
    # load vectors int f.v dx of a smooth body force (nodal f, exact P2 mass)
    loads = [ost.M_vec @ (0.5 * field(X, 0.7 + k * dt)).reshape(-1) for k in range(4)]
    uo, po = u0.copy(), p0.copy()
    ug, pg = d.Function(W, u0.copy()), d.Function(P, p0.copy())

    class Load(object):
        def __init__(self, v):
            self.v = v

        def load_vector(self, W):
            return self.v

    for k in range(3):
        t1 = (k + 1) * dt
        gall = lid(t1)
        uo, po = ost.step(dt, uo, po, (bd, gall[bd]), None, rho, mu, loads[k], loads[k + 1], tol=1e-11)
        bcs = [d.DirichletBC(W, d.Function(W, gall), "on_boundary")]
        ug, pg = st.step(d.Constant(dt), {0: ug}, pg, bcs, [], d.Constant(rho), d.Constant(mu),
                         {0: Load(loads[k]), 1: Load(loads[k + 1])}, verbose=False, tol=1e-11)
        stats = nav.last_stats()
        assert stats["newton_its"] <= 10
    eu = np.linalg.norm(ug._vec - uo) / np.linalg.norm(uo)
    dp = (pg._vec - pg._vec.mean()) - (po - po.mean())
    ep = np.linalg.norm(dp) / np.linalg.norm(po - po.mean())
    assert eu < 1e-8, eu
    assert ep < 1e-7, ep


def test_step_with_pressure_dirichlet_and_component_bc(gpu_ctx):
    """p_bcs non-empty (pressure_correction.py:325-339) and a velocity BC on one component only
    (W.sub(0), as in tests/test_karman_vortex_street.py:138-145): the ds terms are live."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    om = oracle_mesh("tri_leftright")
    m = facade_mesh(om)
    W = d.VectorFunctionSpace(m, "CG", 2)
    P = d.FunctionSpace(m, "CG", 1)
    ost = ons.Rotational(om)
    Wo, Po = ost.W, ost.P
    X = Wo.node_coords
    rng = np.random.default_rng(4)
    u0 = 0.3 * rng.standard_normal(Wo.ndofs)
    p0 = rng.standard_normal(Po.nnodes)
    left = lambda x, on: on and x[0] < -1.0 + 1e-12
    walls = lambda x, on: on and (x[1] < 1e-12 or x[1] > 0.7 - 1e-12)
    right = lambda x, on: on and x[0] > 1.0 - 1e-12
    bcs = [d.DirichletBC(W.sub(0), d.Expression("x[1]*(0.7-x[1])", degree=2), left),
           d.DirichletBC(W, (0.0, 0.0), walls),
           d.DirichletBC(W.sub(0), d.Expression("x[1]*(0.7-x[1])", degree=2), right)]
    pbcs = [d.DirichletBC(P, 0.0, right)]
    ud, uv = d.collect_bcs(bcs, W)
    pd_, pv = d.collect_bcs(pbcs, P)
    dt, rho, mu = 0.02, 1.0, 0.05
    uo, po = ost.step(dt, u0, p0, (ud, uv), (pd_, pv), rho, mu, None, None, tol=1e-12)
    ug, pg = nav.Rotational().step(d.Constant(dt), {0: d.Function(W, u0.copy())}, d.Function(P, p0.copy()), bcs, pbcs,
                                   d.Constant(rho), d.Constant(mu), {0: d.Constant((0.0, 0.0)), 1: d.Constant((0.0, 0.0))},
                                   verbose=False, tol=1e-12)
    assert np.linalg.norm(ug._vec - uo) / np.linalg.norm(uo) < 1e-8
    assert np.linalg.norm(pg._vec - po) / np.linalg.norm(po) < 1e-8


def test_error_contract(gpu_ctx):
    """asserts of pressure_correction.py:488-489 and RuntimeError on non-convergence (:236)."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    m = d.UnitSquareMesh(4, 4)
    W = d.VectorFunctionSpace(m, "CG", 2)
    P = d.FunctionSpace(m, "CG", 1)
    u0, p0 = d.Function(W), d.Function(P)
    f = {0: d.Constant((0.0, 0.0)), 1: d.Constant((0.0, 0.0))}
    with pytest.raises(AssertionError):
        nav.IPCS().step(d.Constant(0.0), {0: u0}, p0, [], [], d.Constant(1.0), d.Constant(1.0), f, verbose=False)
    with pytest.raises(AssertionError):
        nav.IPCS().step(d.Constant(0.1), {0: u0}, p0, [], [], d.Constant(1.0), d.Constant(-1.0), f, verbose=False)
    # NaN input -> RuntimeError, like error_on_nonconvergence
    u0._vec[:] = np.nan
    with pytest.raises(RuntimeError):
        nav.IPCS().step(d.Constant(0.1), {0: u0}, p0, [], [], d.Constant(1.0), d.Constant(1.0), f, verbose=False)


@pytest.mark.parametrize("mesh_name", ["tri_crossed", "tet"])
def test_two_pass_jacobian_assembly_is_deterministic_and_equals_the_atomic_one(gpu_ctx, mesh_name, monkeypatch):
    """The default assembly of J stores the element blocks and sums them per matrix block in a fixed order
    (k_momentum_J_cf<D,1> + k_jac_gather): bit-identical from call to call, and equal to the one-pass fp64-atomic
    assembly (FB_J_TWO_PASS=0) up to the order of the sums."""
    from flow_b200 import _lib
    from flow_b200 import dolfin as d
    from flow_b200._lib import lib
    from flow_b200.navier_stokes.pressure_correction import _engine

    om = oracle_mesh(mesh_name)
    Wo, Po, ui, u0, p0, _ = rand_state(om, 11)
    m = facade_mesh(om)
    W = d.VectorFunctionSpace(m, "CG", 2)
    P = d.FunctionSpace(m, "CG", 1)
    ns = _engine(W, P)
    F = np.zeros(W.dim())
    h = _lib.vp()
    lib.fb_ns_matrix(ns, 2, C.byref(h))
    nrows, nnzb, blk = _lib.i64(), _lib.i64(), C.c_int()
    lib.fb_mat_info(h, C.byref(nrows), C.byref(nnzb), C.byref(blk))
    vals = {}
    for mode, reps in (("1", 3), ("0", 1)):
        monkeypatch.setenv("FB_J_TWO_PASS", mode)
        for r in range(reps):
            _lib.check(lib.fb_ns_residual(ns, 0.05, 1.1, 0.3, 1.0, _lib.as_pd(ui), _lib.as_pd(u0), _lib.as_pd(p0), None, _lib.as_pd(F), 1),
                       m.ctx, "fb_ns_residual")
            v = np.zeros(nnzb.value * blk.value ** 2)
            _lib.check(lib.fb_mat_values(h, _lib.as_pd(v)), m.ctx, "fb_mat_values")
            vals[(mode, r)] = v
    assert np.array_equal(vals[("1", 0)], vals[("1", 1)]) and np.array_equal(vals[("1", 0)], vals[("1", 2)])
    scale = np.abs(vals[("0", 0)]).max()
    assert np.abs(vals[("1", 0)] - vals[("0", 0)]).max() < 1e-14 * scale
