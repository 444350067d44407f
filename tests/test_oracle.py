"""Pin the oracle: the reference's own acceptance thresholds (tests/test_navier_stokes.py:386-445,
tests/test_sealed_box.py:134-141), closed-form integrals, Jacobian consistency, golden fixtures."""
import json
import os

import numpy as np
import pytest
import sympy

import mms_problems as mp
from oracle import fem, forms, navier_stokes as ons, solvers, util

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("dim,degree", [(1, 5), (2, 2), (2, 5), (2, 8), (3, 2), (3, 5), (3, 7)])
def test_quadrature_exact(dim, degree):
    import itertools
    import math

    lam, w = fem.simplex_quadrature(dim, degree)
    assert (w > 0).all() and abs(w.sum() - 1) < 1e-14
    for exps in itertools.product(range(degree + 1), repeat=dim + 1):
        if sum(exps) > degree:
            continue
        exact = math.prod(math.factorial(e) for e in exps) * math.factorial(dim) / math.factorial(sum(exps) + dim)
        assert abs(np.sum(w * np.prod(lam ** np.array(exps), axis=1)) - exact) < 1e-14


def test_p2_mass_closed_form():
    """int phi_i phi_j over the reference triangle, by sympy."""
    x, y = sympy.symbols("x y")
    lam = [1 - x - y, x, y]
    phi = [l * (2 * l - 1) for l in lam] + [4 * lam[a] * lam[b] for a, b in fem.TRI_EDGES]
    M = np.array([[float(sympy.integrate(sympy.integrate(pi * pj, (y, 0, 1 - x)), (x, 0, 1))) for pj in phi] for pi in phi])
    mesh = fem.Mesh(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]]))
    sp2 = fem.Space(mesh, 2)
    Mo = forms.mass_matrix(sp2).toarray()
    # edge nodes are numbered lexicographically (0,1),(0,2),(1,2); local edges are (1,2),(0,2),(0,1)
    perm = list(sp2.cell_nodes[0])
    assert np.allclose(Mo[np.ix_(perm, perm)], M, atol=1e-15)


@pytest.mark.parametrize("name", ["tri", "tet"])
@pytest.mark.parametrize("theta", [1.0, 0.5])
def test_jacobian_is_derivative_of_residual(name, theta):
    mesh = fem.Mesh(*(fem.unit_square_mesh(3, 2, "crossed") if name == "tri" else fem.unit_cube_mesh(2, 1, 1)))
    W, P = fem.Space(mesh, 2, mesh.dim), fem.Space(mesh, 1, 1)
    rng = np.random.default_rng(0)
    ui, u0, p0 = rng.standard_normal(W.ndofs), rng.standard_normal(W.ndofs), rng.standard_normal(P.nnodes)
    z = np.zeros(W.ndofs)
    F, J = forms.momentum_residual_jacobian(W, P, ui, u0, p0, z, 0.3, 1.2, 0.7, theta)
    v = rng.standard_normal(W.ndofs)
    h = 1e-6
    Fp, _ = forms.momentum_residual_jacobian(W, P, ui + h * v, u0, p0, z, 0.3, 1.2, 0.7, theta, want_J=False)
    Fm, _ = forms.momentum_residual_jacobian(W, P, ui - h * v, u0, p0, z, 0.3, 1.2, 0.7, theta, want_J=False)
    fd = (Fp - Fm) / (2 * h)
    assert np.abs(fd - J @ v).max() / np.abs(fd).max() < 1e-8


def test_mms_ipcs_order_and_golden():
    """test_ipcs of the reference (meshes 8,16,32; Dt 1,0.5; guermond2): u order > 1.9, p order > 0.9."""
    from golden.make_golden import time_errors

    Dt = [1.0, 0.5]
    e = time_errors(mp.problem_guermond2, ons.IPCS, [8, 16, 32], Dt)
    ou = mp.compute_numerical_order_of_convergence(Dt, e["u"].T).T
    op = mp.compute_numerical_order_of_convergence(Dt, e["p"].T).T
    assert (ou[:, 0] > 2.0 - 0.1).all() and (op[:, 0] > 1.0 - 0.1).all()
    gold = json.load(open(os.path.join(HERE, "golden", "mms_ipcs_guermond2.json")))
    assert np.allclose(e["u"], gold["u"], rtol=1e-8) and np.allclose(e["p"], gold["p"], rtol=1e-8)


def test_mms_chorin_order_flat():
    """test_chorin[problem_flat] of the reference: u order > 0.9, p order > 0.4."""
    from golden.make_golden import time_errors

    Dt = [1.0e-3, 0.5e-3]
    e = time_errors(mp.problem_flat, ons.Chorin, [16], Dt)
    assert mp.compute_numerical_order_of_convergence(Dt, e["u"].T)[0, 0] > 0.9
    assert mp.compute_numerical_order_of_convergence(Dt, e["p"].T)[0, 0] > 0.4


def test_mms_rotational_order():
    """test_rotational of the reference on the coarser of its two meshes (n=32): u > 1.9, p > 1.4."""
    from golden.make_golden import time_errors

    Dt = [1.0e-2, 0.5e-2]
    e = time_errors(mp.problem_guermond1, ons.Rotational, [32], Dt)
    assert mp.compute_numerical_order_of_convergence(Dt, e["u"].T)[0, 0] > 1.9
    assert mp.compute_numerical_order_of_convergence(Dt, e["p"].T)[0, 0] > 1.4


def test_sealed_box_invariant():
    """tests/test_sealed_box.py:85-141 on a structured mesh: |u|_inf < 1e-13 after two IPCS steps."""
    mesh = fem.Mesh(*fem.rectangle_mesh((0.0, 0.0), (0.1, 0.2), 6, 12, "left/right"))
    st = ons.IPCS(mesh)
    W, P = st.W, st.P
    g, rho, mu, dt = -9.81, 998.21, 1.002e-3, 1e-2
    u, p = np.zeros(W.ndofs), g * P.node_coords[:, 1]
    bd = W.boundary_dofs()
    load = forms.expression_load_vector(W, lambda X: np.tile([0.0, g], (X.shape[0], 1)), 0)
    for _ in range(2):
        u, p = st.step(dt, u, p, (bd, np.zeros(bd.size)), None, rho, mu, load, load, tol=1e-10)
    assert np.sqrt((u.reshape(-1, 2) ** 2).sum(1)).max() < 1e-13


def test_small_step_golden():
    from golden.make_golden import small_step_fixture

    gold = json.load(open(os.path.join(HERE, "golden", "rotational_small_step.json")))
    now = small_step_fixture()
    for k in gold:
        for f in ("u1", "p1"):
            assert np.allclose(now[k][f], gold[k][f], rtol=1e-9, atol=1e-12)


def test_krylov_mode_matches_lu():
    mesh = fem.Mesh(*fem.unit_cube_mesh(3, 3, 3))
    out = {}
    for lin in ("lu", "krylov"):
        st = ons.IPCS(mesh, linear=lin)
        W = st.W
        bd = W.boundary_dofs()
        g = np.zeros((W.nnodes, 3))
        g[W.node_coords[:, 2] > 1 - 1e-12, 0] = 1.0
        out[lin] = st.step(1e-2, np.zeros(W.ndofs), np.zeros(st.P.nnodes), (bd, g.reshape(-1)[bd]), None, 1.0, 1e-2, tol=1e-10)
    assert np.linalg.norm(out["lu"][0] - out["krylov"][0]) / np.linalg.norm(out["lu"][0]) < 1e-8


def test_stokes_order_oracle():
    """tests/test_stokes.py:102-117 on the oracle: spatial orders of u and p exceed 1.9."""
    from oracle import stokes as ostokes

    pr = mp.stokes_guermond1()
    hmax, ue, pe = [], [], []
    for n in (8, 16):
        mesh = fem.Mesh(*fem.unit_square_mesh(n, n, "left/right"))
        W, P = fem.Space(mesh, 2, 2), fem.Space(mesh, 1, 1)
        load = forms.expression_load_vector(W, pr["f"], mp.MAX_DEGREE)
        ubd, pbd = W.boundary_dofs(), P.boundary_dofs()
        u, p = ostokes.solve(mesh, pr["mu"], load, (ubd, util.interpolate(W, pr["u"])[ubd]), (pbd, util.interpolate(P, pr["p"])[pbd]))
        hmax.append(mesh.hmax())
        ue.append(util.errornorm(W, pr["u"], u))
        pe.append(util.errornorm(P, pr["p"], p))
    assert mp.compute_numerical_order_of_convergence(hmax, ue)[0] > 1.9
    assert mp.compute_numerical_order_of_convergence(hmax, pe)[0] > 1.9


def test_heat_lumped_mass_and_maximum_principle():
    """heat.py:39-45: vertex-lumped mass is diagonal with zero rows at P2 edge nodes; pure diffusion
    with Dirichlet data between 293 and 320 stays inside that range for P1."""
    from oracle import heat as oheat

    mesh = fem.Mesh(*fem.unit_square_mesh(6, 6, "crossed"))
    V2 = fem.Space(mesh, 2, 1)
    M = forms.lumped_vertex_mass(V2).diagonal()
    assert (M[: mesh.nv] > 0).all() and (M[mesh.nv:] == 0).all() and abs(M.sum() - 1.0) < 1e-14
    V1 = fem.Space(mesh, 1, 1)
    bd = V1.boundary_dofs()
    vals = np.where(V1.node_coords[bd, 0] < 1e-12, 320.0, 293.0)
    h = oheat.Heat(mesh, 1, None, 0.6, 1000.0, 4.0, (bd, vals))
    th = np.full(V1.nnodes, 293.0)
    for _ in range(3):
        th = oheat.implicit_euler_step(h, th, 0.0, 10.0)
    assert th.min() > 293.0 - 1e-9 and th.max() < 320.0 + 1e-9


def test_compiled_cpu_baseline_matches_numpy_oracle():
    """oracle/_cstep.so (bench.py's CPU arm) vs the numpy oracle: two cavity steps, 1e-8."""
    from oracle import cstep

    cv = cstep.CavityCPU(4)
    ost = ons.IPCS(cv.mesh)
    uo, po = np.zeros(cv.W.ndofs), np.zeros(cv.P.nnodes)
    uc, pc = uo.copy(), po.copy()
    for _ in range(2):
        uo, po = ost.step(1e-2, uo, po, cv.bc, None, 1.0, 1e-2, None, None, tol=1e-10)
        uc, pc, stats = cv.step(uc, pc)
    assert np.linalg.norm(uc - uo) / np.linalg.norm(uo) < 1e-8
    assert np.linalg.norm((pc - pc.mean()) - (po - po.mean())) / np.linalg.norm(po - po.mean()) < 1e-7
    assert stats[0] <= 3


def test_scalar_operator_preconditioner_clusters_the_momentum_jacobian():
    """Design rationale of the FGMRES momentum solver, checked on the oracle's matrices at the
    benchmark's cell CFL (0.74) and diffusion number (0.55): the Jacobian J of pressure_correction.py:202 is
    S (x) I + (viscous coupling + linearised convection) with S = M + theta dt nu K; with a few CG iterations on S as
    the (variable) preconditioner, flexible GMRES needs ~4x fewer products with J than block-Jacobi BiCGStab."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla

    n = 6
    h = 1.0 / n
    dt = 0.74 * h           # |u| ~ 1 at the lid
    mu = 0.55 * h * h / dt  # rho = 1
    om = fem.Mesh(*fem.unit_cube_mesh(n, n, n))
    W, P, Wn = fem.Space(om, 2, 3), fem.Space(om, 1, 1), fem.Space(om, 2, 1)
    X = W.node_coords
    bd = W.boundary_dofs()
    g = np.zeros((W.nnodes, 3))
    g[X[:, 2] > 1 - 1e-12, 0] = 1.0
    g = g.reshape(-1)
    rng = np.random.default_rng(0)
    u0 = 0.3 * np.stack([np.sin(3 * X[:, 1] + X[:, 2]), np.cos(2 * X[:, 0]) * X[:, 2], np.sin(X[:, 0] + 2 * X[:, 1])], 1).reshape(-1)
    u0[bd] = g[bd]
    F, J = forms.momentum_residual_jacobian(W, P, u0, u0, rng.standard_normal(P.nnodes), np.zeros(W.ndofs), dt, 1.0, mu, 1.0)
    J, b = forms.apply_bc_rows(J, F, bd, np.zeros(bd.size))
    J = J.tocsr()
    b = b.copy()
    b[bd] = 0.0  # lifted right-hand side: the Krylov space lives on the free dofs
    S = (forms.mass_matrix(Wn) + dt * mu * forms.stiffness_matrix(Wn)).tocsr()
    free = np.ones(W.ndofs)
    free[bd] = 0.0
    Dm = sp.diags(free)
    Sv = (Dm @ sp.kron(S, sp.eye(3), format="csr") @ Dm + sp.diags(1.0 - free)).tocsr()
    dS = Sv.diagonal()

    def inner_cg(r, its=4):
        x = np.zeros_like(r)
        rr = r.copy()
        z = rr / dS
        p = z.copy()
        rz = rr @ z
        for _ in range(its):
            if rz == 0.0:
                break
            Ap = Sv @ p
            a = rz / (p @ Ap)
            x += a * p
            rr -= a * Ap
            z = rr / dS
            rz2 = rr @ z
            p = z + (rz2 / rz) * p
            rz = rz2
        return x

    # flexible GMRES (Arnoldi on J Z, Z_j = inner_cg(V_j))
    m = 40
    nb = np.linalg.norm(b)
    V = np.zeros((m + 1, b.size))
    Z = np.zeros((m, b.size))
    H = np.zeros((m + 1, m))
    V[0] = b / nb
    e1 = np.zeros(m + 1)
    e1[0] = nb
    outer = None
    for j in range(m):
        Z[j] = inner_cg(V[j])
        w = J @ Z[j]
        for i in range(j + 1):
            H[i, j] = w @ V[i]
            w -= H[i, j] * V[i]
        H[j + 1, j] = np.linalg.norm(w)
        V[j + 1] = w / H[j + 1, j]
        y = np.linalg.lstsq(H[:j + 2, :j + 1], e1[:j + 2], rcond=None)[0]
        if np.linalg.norm(H[:j + 2, :j + 1] @ y - e1[:j + 2]) <= 1e-6 * nb:
            outer = j + 1
            break
    assert outer is not None and outer <= 16, outer
    x = Z[:outer].T @ y
    assert np.linalg.norm(b - J @ x) <= 2e-6 * nb
    # block-Jacobi BiCGStab on the same system
    N = W.ndofs // 3
    Jb = J.tobsr((3, 3))
    D = np.zeros((N, 3, 3))
    for i in range(N):
        for k in range(Jb.indptr[i], Jb.indptr[i + 1]):
            if Jb.indices[k] == i:
                D[i] = Jb.data[k]
    Dinv = np.linalg.inv(D)
    count = [0]

    def op(v):
        count[0] += 1
        return J @ np.einsum("nij,nj->ni", Dinv, v.reshape(N, 3)).ravel()

    _, info = sla.bicgstab(sla.LinearOperator(J.shape, matvec=op), b, rtol=1e-6, atol=0.0, maxiter=400)
    assert info == 0
    assert count[0] >= 3 * outer, (count[0], outer)  # measured on B200 at n = 74: 78 -> 22 Jacobian products per step


def test_semi_implicit_linearisation_is_linear_and_consistent():
    """Oracle variant of opts.semi_implicit ((u0 . grad) ui, pressure_correction.py:96-101): F1 is affine in ui (its
    Jacobian does not depend on ui and one Newton update lands on the root), J = dF/dui by finite differences, and the
    form coincides with the reference's at ui = u0."""
    om = fem.Mesh(*fem.unit_cube_mesh(2, 2, 2))
    W, P = fem.Space(om, 2, 3), fem.Space(om, 1, 1)
    rng = np.random.default_rng(5)
    u0, ua, ub = (rng.standard_normal(W.ndofs) for _ in range(3))
    p0 = rng.standard_normal(P.nnodes)
    load = np.zeros(W.ndofs)
    args = (0.05, 1.3, 0.2, 1.0)
    Fa, Ja = forms.momentum_residual_jacobian(W, P, ua, u0, p0, load, *args, semi_implicit=True)
    Fb, Jb = forms.momentum_residual_jacobian(W, P, ub, u0, p0, load, *args, semi_implicit=True)
    assert abs(Ja - Jb).max() < 1e-14                                    # J independent of ui
    assert np.abs(Fb - Fa - Ja @ (ub - ua)).max() < 1e-12 * np.abs(Fb).max()  # affine
    F0s, _ = forms.momentum_residual_jacobian(W, P, u0, u0, p0, load, *args, want_J=False, semi_implicit=True)
    F0r, _ = forms.momentum_residual_jacobian(W, P, u0, u0, p0, load, *args, want_J=False)
    assert np.abs(F0s - F0r).max() < 1e-13 * np.abs(F0r).max()
    # and one step of the stepper needs exactly one Newton update
    st = ons.IPCS(om, semi_implicit=True)
    bd = W.boundary_dofs()
    g = np.zeros(W.ndofs)
    u1, p1 = st.step(0.05, 0.1 * u0, 0.1 * p0, (bd, g[bd]), None, 1.3, 0.2, None, None, tol=1e-12)
    assert st.info["newton_its"] == 1 and np.isfinite(u1).all()
