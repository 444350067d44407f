"""Solver variants of the IPCS step against the oracle on a mesh large enough for the AMG hierarchy
(>= 4096 pressure unknowns): smoothed-aggregation AMG vs Jacobi for the Poisson solve
(pressure_correction.py:331, :414-419 use hypre BoomerAMG), warm-started CG, the chord Jacobian
carried across steps, and the fp32-stored chord Jacobian.  Every variant has to land on the
oracle's Newton+LU result within the north-star tolerance (1e-8 relative L2 after several steps)."""
import ctypes as C

import numpy as np
import pytest

from oracle import fem, forms
from oracle import navier_stokes as ons

pytestmark = pytest.mark.gpu

CHORD = dict(jacobian_reuse=1, jacobian_across_steps=1, adaptive_forcing=1)  # the round-1 default path, now opt-in

VARIANTS = {
    "defaults": {},
    "jacobi_cold": dict(pressure_precond="jacobi", warm_start=0),
    "jacobi_warm": dict(pressure_precond="jacobi", warm_start=1),
    "amg_cold": dict(pressure_precond="amg", warm_start=0),
    "fp32_jacobian": dict(jacobian_fp32=1),
    "bicgstab": dict(momentum_solver="bicgstab"),
    "fgmres_inner8": dict(momentum_solver="fgmres", momentum_inner_its=8),
    "fgmres_inner_fp32": dict(momentum_solver="fgmres", inner_fp32=1),
    "inner_cg": dict(inner_chebyshev=0),
    "momentum_amg_on_S": dict(momentum_amg_kappa=1.0),   # force the AMG V-cycle on S as the FGMRES preconditioner
    "momentum_no_amg": dict(momentum_amg_kappa=0.0),
    "chebyshev_degree6": dict(chebyshev_degree=6),
    "chord": dict(CHORD),
    "chord_within_step_only": dict(jacobian_reuse=1, jacobian_across_steps=0, adaptive_forcing=0),
    "chord_extrapolated_start": dict(CHORD, extrapolate_guess=1),
    "chord_bicgstab": dict(CHORD, momentum_solver="bicgstab"),
    "chord_fp32_jacobian": dict(CHORD, jacobian_fp32=1),
}


@pytest.fixture(scope="module")
def cavity2d():
    """Regularised lid-driven cavity on UnitSquareMesh(70, 70): 5041 pressure dofs, 39762 velocity dofs;
    4 oracle steps (Newton + LU, CG to 1e-12)."""
    n = 70
    om = fem.Mesh(*fem.unit_square_mesh(n, n, "right"))
    ost = ons.IPCS(om)
    X = ost.W.node_coords
    bd = ost.W.boundary_dofs()
    g = np.zeros((ost.W.nnodes, 2))
    top = X[:, 1] > 1.0 - 1e-12
    g[top, 0] = (1.0 - (2.0 * X[top, 0] - 1.0) ** 4)
    g = g.reshape(-1)
    dt, rho, mu = 0.02, 1.0, 0.01
    # "reference": _step (:468-518) exactly as the reference runs it (Newton |F| < 1e-10 hard-wired at :499, LU updates);
    # "root": Newton driven to 1e-14, i.e. the root of F1 -- what the chord variants converge to
    states = {}
    for kind, newton_tol in (("reference", 1e-10), ("root", 1e-14)):
        u, p = np.zeros(ost.W.ndofs), np.zeros(ost.P.nnodes)
        states[kind] = []
        for _ in range(4):
            ui = ost.tentative_velocity(u, p, None, None, (bd, g[bd]), rho, mu, dt, tol=newton_tol)
            p1 = ost.pressure(ui, p, None, rho, mu, dt, 1e-12)
            u = ost.velocity_correction(ui, p1, p, (bd, g[bd]), rho, mu, dt, 1e-12)
            p = p1
            states[kind].append((u.copy(), p.copy()))
    return om, g, (dt, rho, mu), states


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_matches_oracle(gpu_ctx, cavity2d, name):
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    om, g, (dt, rho, mu), states = cavity2d
    nav.reset_options()
    # Solver options as they ship (no newton_atol override).  The default path reproduces the reference's Newton iterates
    # and is compared with the oracle at the reference's settings; the chord variants converge to the root of F1 and are
    # compared with the oracle driven to the root (the two oracle runs differ by up to 3e-8 on this mesh: that is the
    # reference's own distance from the root when its last exact update lands just below 1e-10).
    nav.set_options(**VARIANTS[name])
    kind = "root" if VARIANTS[name].get("jacobian_reuse", 0) else "reference"
    try:
        m = d.Mesh(om.points, om.cells)
        W = d.VectorFunctionSpace(m, "CG", 2)
        P = d.FunctionSpace(m, "CG", 1)
        bcs = [d.DirichletBC(W, d.Function(W, g.copy()), "on_boundary")]
        zero = d.Constant((0.0, 0.0))
        u, p = d.Function(W), d.Function(P)
        st = nav.IPCS()
        assemblies, pits = [], []
        for k in range(4):
            u, p = st.step(d.Constant(dt), {0: u}, p, bcs, [], d.Constant(rho), d.Constant(mu), {0: zero, 1: zero},
                           verbose=False, tol=1e-12)
            s = nav.last_stats()
            assemblies.append(s["jacobian_assemblies"])
            pits.append(s["pressure_its"])
            uo, po = states[kind][k]
            eu = np.linalg.norm(u._vec - uo) / np.linalg.norm(uo)
            dp = (p._vec - p._vec.mean()) - (po - po.mean())
            ep = np.linalg.norm(dp) / np.linalg.norm(po - po.mean())
            assert eu < 1e-8, (name, k, eu)
            assert ep < 1e-7, (name, k, ep)
        if VARIANTS[name].get("pressure_precond", "amg") == "amg":
            assert max(pits) < 60, pits  # mesh-independent iteration counts (Jacobi needs several hundred here)
        if kind == "reference":
            assert min(assemblies) >= 2  # the reference's Newton assembles at every iteration
        if name == "chord":
            assert sum(assemblies) < 6   # ... the chord variant keeps its operator
    finally:
        nav.reset_options()


def test_amg_hierarchy_and_dirichlet_variant(gpu_ctx):
    """Hierarchy statistics through the C ABI, and the p_bcs != [] branch (:325-339) with AMG on the
    symmetrically eliminated matrix: same pressure as Jacobi-CG to solver tolerance."""
    from flow_b200 import _lib
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav
    from flow_b200._lib import lib
    from flow_b200.navier_stokes.pressure_correction import _engine

    n = 70
    res = {}
    for pc in ("amg", "jacobi"):
        nav.reset_options()
        nav.set_options(pressure_precond=pc)
        try:
            m = d.UnitSquareMesh(n, n)
            W = d.VectorFunctionSpace(m, "CG", 2)
            P = d.FunctionSpace(m, "CG", 1)
            if pc == "amg":
                ns = _engine(W, P)
                lv, cx, sz = C.c_int(), C.c_double(), (C.c_int * 12)()
                _lib.check(lib.fb_ns_amg_info(ns, C.byref(lv), C.byref(cx), sz, 12), W.mesh().ctx, "amg_info")
                sizes = list(sz)[:lv.value]
                assert lv.value >= 2 and sizes[0] == (n + 1) ** 2 and sizes[-1] <= 400
                assert all(a > 2 * b for a, b in zip(sizes, sizes[1:]))  # every level coarsens by more than 2x
                assert 1.0 < cx.value < 2.5
            inflow = d.DirichletBC(W, d.Expression(("x[1]*(1-x[1])", "0.0"), degree=2), lambda x, on: on and x[0] < 1e-12)
            walls = d.DirichletBC(W, (0.0, 0.0), lambda x, on: on and (x[1] < 1e-12 or x[1] > 1 - 1e-12))
            pout = d.DirichletBC(P, 0.0, lambda x, on: on and x[0] > 1 - 1e-12)
            zero = d.Constant((0.0, 0.0))
            u, p = d.Function(W), d.Function(P)
            for _ in range(2):
                u, p = nav.Rotational().step(d.Constant(0.05), {0: u}, p, [inflow, walls], [pout], d.Constant(1.0),
                                             d.Constant(0.02), {0: zero, 1: zero}, verbose=False, tol=1e-12)
            res[pc] = (u._vec.copy(), p._vec.copy(), nav.last_stats()["pressure_its"])
        finally:
            nav.reset_options()
    ua, pa, ia = res["amg"]
    uj, pj, ij = res["jacobi"]
    assert np.linalg.norm(ua - uj) / np.linalg.norm(uj) < 1e-9
    assert np.linalg.norm(pa - pj) / np.linalg.norm(pj) < 1e-8
    assert ia < 40 and ia < ij / 4, (ia, ij)


def test_coordinate_node_order_same_solution(gpu_ctx):
    """mesh.node_order = "lexicographic" (experimental locality numbering, tests/test_topology.py) changes the
    numbering only: two IPCS steps give the canonical solution under the node permutation."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    res = {}
    for order in (None, "lexicographic"):
        mesh = d.UnitCubeMesh(6, 5, 7)
        mesh.node_order = order
        W = d.VectorFunctionSpace(mesh, "CG", 2)
        P = d.FunctionSpace(mesh, "CG", 1)
        bcs = [d.DirichletBC(W, (0.0, 0.0, 0.0), "on_boundary"),
               d.DirichletBC(W, d.Expression(("4*x[0]*(1-x[0])", "0.0", "0.0"), degree=2), lambda x, on: x[2] > 1 - 1e-12)]
        f = d.Constant((0.0, 0.1, -1.0))
        u, p = d.Function(W), d.Function(P)
        for _ in range(2):
            u, p = nav.IPCS().step(d.Constant(0.02), {0: u}, p, bcs, [], d.Constant(1.0), d.Constant(0.02), {0: f, 1: f},
                                   verbose=False, tol=1e-11)
        res[order] = (u._vec.copy(), p._vec.copy(), W.nodes.perm, P.nodes.perm)
    ua, pa, _, _ = res[None]
    ub, pb, permW, permP = res["lexicographic"]
    assert permW is not None and permP is not None
    assert np.linalg.norm(ub.reshape(-1, 3)[permW] - ua.reshape(-1, 3)) / np.linalg.norm(ua) < 1e-8
    da = pa - pa.mean()
    assert np.linalg.norm((pb - pb.mean())[permP] - da) / np.linalg.norm(da) < 1e-7


@pytest.mark.parametrize("dim", [2, 3])
def test_semi_implicit_option_matches_its_oracle_variant(gpu_ctx, dim):
    """opts.semi_implicit = 1 ((u0 . grad) ui, the linearisation of pressure_correction.py:96-101): three steps against the
    oracle's variant of the same form -- one Jacobian assembly per step, forcing and a component
    (W.sub(0)-like full vector) Dirichlet condition included."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    nav.reset_options()
    nav.set_options(semi_implicit=1)
    try:
        if dim == 2:
            om = fem.Mesh(*fem.unit_square_mesh(12, 10, "crossed"))
            fvec, top = (0.3, -1.0), 1
        else:
            om = fem.Mesh(*fem.unit_cube_mesh(5, 4, 6))
            fvec, top = (0.0, 0.2, -1.0), 2
        ost = ons.IPCS(om, semi_implicit=True)
        X = ost.W.node_coords
        bd = ost.W.boundary_dofs()
        g = np.zeros((ost.W.nnodes, dim))
        lid = X[:, top] > 1.0 - 1e-12
        g[lid, 0] = 4.0 * X[lid, 0] * (1.0 - X[lid, 0])
        g = g.reshape(-1)
        dt, rho, mu = 0.05, 1.2, 0.02
        load = forms.mass_matrix(fem.Space(om, 2, 1))
        import scipy.sparse as sp

        lvec = sp.kron(load, sp.eye(dim)) @ np.tile(np.array(fvec), ost.W.nnodes)
        m = d.Mesh(om.points, om.cells)
        W = d.VectorFunctionSpace(m, "CG", 2)
        P = d.FunctionSpace(m, "CG", 1)
        bcs = [d.DirichletBC(W, d.Function(W, g.copy()), "on_boundary")]
        f = d.Constant(fvec)
        uo, po = np.zeros(ost.W.ndofs), np.zeros(ost.P.nnodes)
        u, p = d.Function(W), d.Function(P)
        for k in range(3):
            uo, po = ost.step(dt, uo, po, (bd, g[bd]), None, rho, mu, lvec, lvec, tol=1e-11)
            u, p = nav.IPCS().step(d.Constant(dt), {0: u}, p, bcs, [], d.Constant(rho), d.Constant(mu), {0: f, 1: f},
                                   verbose=False, tol=1e-11)
            s = nav.last_stats()
            eu = np.linalg.norm(u._vec - uo) / np.linalg.norm(uo)
            dp = (p._vec - p._vec.mean()) - (po - po.mean())
            ep = np.linalg.norm(dp) / np.linalg.norm(po - po.mean())
            assert eu < 1e-8 and ep < 1e-7, (dim, k, eu, ep, s)
            # linear in ui: ONE assembly; a second update, if any, is iterative refinement with the same matrix
            assert s["jacobian_assemblies"] == 1 and s["newton_its"] <= 2, s
    finally:
        nav.reset_options()
