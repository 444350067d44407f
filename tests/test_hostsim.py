"""Element-level math of the CUDA kernels (flow_b200/csrc/fb_element.cuh, compiled for the host by
tests/hostsim) vs the oracle's independent einsum formulation.  Tolerance 1e-13 (pure rounding)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from common import MESHES, oracle_mesh, rand_state, rel
from oracle import fem, forms


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("theta", [1.0, 0.5, 0.0])
@pytest.mark.parametrize("closed_form", [0, 1])
def test_momentum_element_math(hostsim, name, theta, closed_form):
    """closed_form = 1: the cell part of J from fb_jac_pair (exact integrals through the vertex values of the affine
    gradients and the M3 table, as k_momentum_J_cf computes it) instead of the degree-5 quadrature of fb_jac_point."""
    hostsim.hs_set_closed_form(closed_form)
    om = oracle_mesh(name)
    W, P, ui, u0, p0, _ = rand_state(om)
    n = W.ndofs
    dt, rho, mu = 0.37, 1.3, 0.71
    F = np.zeros(n)
    J = np.zeros((n, n))
    hostsim.hs_momentum(om.dim, C.c_int64(om.nc), _pi(W.cell_nodes), _pd(om.points), C.c_int64(om.bfacet_cell.size),
                        _pi(om.bfacet_cell), _pi(om.bfacet_local), C.c_double(dt), C.c_double(rho), C.c_double(mu),
                        C.c_double(theta), _pd(ui), _pd(u0), _pd(p0), C.c_int64(n), _pd(F), _pd(J))
    Fo, Jo = forms.momentum_residual_jacobian(W, P, ui, u0, p0, np.zeros(n), dt, rho, mu, theta)
    hostsim.hs_set_closed_form(0)
    assert rel(F, Fo) < 1e-13
    assert rel(J, Jo.toarray()) < 1e-13


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("rotational", [0, 1])
def test_rhs_element_math(hostsim, name, rotational):
    om = oracle_mesh(name)
    W, P, ui, _, p0, p1 = rand_state(om, 3)
    dt, rho, mu = 0.21, 0.9, 1.7
    bp = np.zeros(P.nnodes)
    bu = np.zeros(W.ndofs)
    hostsim.hs_rhs(om.dim, C.c_int64(om.nc), _pi(W.cell_nodes), _pd(om.points), C.c_double(dt), C.c_double(rho),
                   C.c_double(mu), rotational, _pd(ui), _pd(p1), _pd(p0), _pd(bp), _pd(bu))
    assert rel(bp, forms.pressure_rhs(W, P, ui, p0, dt, rho, mu, bool(rotational))) < 1e-13
    Mv = sp.kron(forms.mass_matrix(fem.Space(om, 2, 1)), sp.eye(om.dim))
    assert rel(bu, forms.correction_rhs(W, P, ui, p1, p0, dt, rho, mu, bool(rotational)) - Mv @ ui) < 1e-12


def test_supg_tau_device_routine(hostsim):
    """fb_supg_tau (the routine k_heat_supg calls) vs the oracle's restatement of stabilization.py:50-143,
    including the |b| ~ 0 early return and the small-Peclet Taylor branch."""
    from oracle import heat as oheat

    hostsim.hs_supg_tau.restype = C.c_double
    om = oracle_mesh("tri_leftright")
    W = fem.Space(om, 2, 2)
    rng = np.random.default_rng(0)
    for scale, eps in ((1.0, 0.6), (1e-7, 0.6), (1e-12, 0.6), (50.0, 1e-3)):
        conv = scale * rng.standard_normal(W.ndofs)
        ref = oheat.supg_tau(om, W, conv, eps, 2)
        V = conv.reshape(-1, 2)[W.cell_nodes[:, :3]]
        X = np.ascontiguousarray(om.points[om.cells].reshape(om.nc, 6))
        got = np.array([[hostsim.hs_supg_tau(_pd(X[c]), _pd(np.ascontiguousarray(V[c, v])), C.c_double(eps), 2) for v in range(3)]
                        for c in range(om.nc)])
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("variant", ["neumann", "dirichlet"])
@pytest.mark.parametrize("meshname", ["cube14", "square70"])
def test_amg_hierarchy_and_pcg(hostsim, variant, meshname):
    """Host set-up of the smoothed-aggregation AMG (flow_b200/csrc/fb_amg_host.h, the code whose output the library
    uploads) driven by a host V-cycle + PCG: hierarchy statistics, mesh-independent iteration counts, and the solution
    of the pressure Poisson problem of pressure_correction.py:317-339 (Dirichlet) / :340-432 (pure Neumann, singular,
    consistent right-hand side) against scipy's direct solve."""
    import scipy.sparse.linalg as sla

    if meshname == "cube14":
        om = fem.Mesh(*fem.unit_cube_mesh(14, 14, 14))
    else:
        om = fem.Mesh(*fem.unit_square_mesh(70, 70, "right"))
    P = fem.Space(om, 1, 1)
    A = forms.stiffness_matrix(P).tocsr()
    A.sort_indices()
    n = A.shape[0]
    X = P.node_coords
    rng = np.random.default_rng(1)
    b = rng.standard_normal(n)
    if variant == "neumann":
        b -= b.mean()  # consistent: b orthogonal to the constants
    else:
        bd = np.nonzero(X[:, 0] > 1.0 - 1e-12)[0]
        A, b = forms.apply_bc_symmetric(A, b, bd, np.zeros(bd.size))
        A = A.tocsr()
        A.sort_indices()
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data)
    x = np.zeros(n)
    lev, cx, sing, dense = C.c_int(), C.c_double(), C.c_int(), C.c_int()
    sizes = (C.c_int * 12)()
    its = hostsim.hs_amg_pcg(n, _pi(rp), _pi(ci), _pd(va), _pd(b), _pd(x), C.c_double(1e-10), 200, C.byref(lev), sizes,
                             C.byref(cx), C.byref(sing), C.byref(dense))
    sz = list(sizes)[:lev.value]
    assert lev.value >= 2 and sz[0] == n and sz[-1] <= 400 and dense.value == 1
    assert all(a > 2 * c for a, c in zip(sz, sz[1:]))  # every level coarsens by more than 2x
    assert 1.0 < cx.value < 2.6
    assert sing.value == (1 if variant == "neumann" else 0)
    assert 0 < its < 40, its  # Jacobi-CG needs several hundred iterations on these meshes
    if variant == "neumann":
        ref = sla.lsqr(A, b, atol=1e-14, btol=1e-14, iter_lim=20000)[0]
        d = (x - x.mean()) - (ref - ref.mean())
        assert np.linalg.norm(d) / np.linalg.norm(ref - ref.mean()) < 1e-6
        assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-8
    else:
        ref = sla.spsolve(A.tocsc(), b)
        assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-8
