"""Size-independent properties at BASELINE.json's benchmark size (UnitCubeMesh(74): 10 345 722 dofs), where
no oracle result exists: operator identities (a checksum of checksums), linearity and symmetry of the
products, and the hydrostatic invariant of tests/test_sealed_box.py:134-141 through the full IPCS step."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 74


@pytest.fixture(scope="module")
def cube():
    from flow_b200 import dolfin as d
    from flow_b200.navier_stokes.pressure_correction import _engine

    mesh = d.UnitCubeMesh(N, N, N)
    W = d.VectorFunctionSpace(mesh, "CG", 2)
    P = d.FunctionSpace(mesh, "CG", 1)
    return mesh, W, P, _engine(W, P)


def _spmv(h, ncomp, x):
    from flow_b200 import _lib
    from flow_b200._lib import lib

    y = np.empty_like(x)
    _lib.check(lib.fb_mat_spmv(h, ncomp, _lib.as_pd(x), _lib.as_pd(y)), None, "fb_mat_spmv")
    return y


def test_sizes_and_operator_identities(gpu_ctx, cube):
    from flow_b200 import _lib
    from flow_b200._lib import lib

    mesh, W, P, ns = cube
    assert W.dim() == 9923847 and P.dim() == 421875  # config 5
    h = _lib.vp()
    rng = np.random.default_rng(0)
    # P1 stiffness: constants in the kernel, symmetric, energy of x = coordinate equals the volume
    lib.fb_ns_matrix(ns, 0, C.byref(h))
    one = np.ones(P.dim())
    assert np.abs(_spmv(h, 1, one)).max() < 1e-10
    x, y = rng.standard_normal(P.dim()), rng.standard_normal(P.dim())
    Ax, Ay = _spmv(h, 1, x), _spmv(h, 1, y)
    assert abs(y @ Ax - x @ Ay) < 1e-9 * abs(y @ Ax)
    assert np.abs(_spmv(h, 1, 2.0 * x - 3.0 * y) - (2.0 * Ax - 3.0 * Ay)).max() < 1e-9 * np.abs(Ax).max()
    cx = P.tabulate_dof_coordinates()[:, 0].copy()
    assert abs(cx @ _spmv(h, 1, cx) - 1.0) < 1e-10  # int |grad x|^2 = |Omega| = 1
    # P2 mass (shared by the three velocity components): total mass = volume, per component
    lib.fb_ns_matrix(ns, 1, C.byref(h))
    ones = np.ones(W.dim())
    M1 = _spmv(h, 3, ones)
    assert abs(M1.reshape(-1, 3)[:, 0].sum() - 1.0) < 1e-10 and abs(M1.sum() - 3.0) < 1e-10
    xv, yv = rng.standard_normal(W.dim()), rng.standard_normal(W.dim())
    Mx, My = _spmv(h, 3, xv), _spmv(h, 3, yv)
    assert abs(yv @ Mx - xv @ My) < 1e-9 * abs(yv @ Mx)
    assert xv @ Mx > 0.0
    # int x^2 over the unit cube = 1/3 (P2 represents x^2 exactly)
    X = W.tabulate_dof_coordinates()
    q = np.zeros(W.dim())
    q[0::3] = X[0::3, 0]
    assert abs(q @ _spmv(h, 3, q) - 1.0 / 3.0) < 1e-10


def test_sealed_box_full_size(gpu_ctx, cube):
    """f = (0, 0, g) balanced by p0 = g z: the velocity stays zero through two full IPCS steps (assembly of F and
    J, Newton, pressure Poisson with AMG, velocity correction) at 10.3 M dofs."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    mesh, W, P, _ = cube
    g = -9.81
    u0 = d.Function(W)
    p0 = d.interpolate(d.Expression("g*x[2]", degree=1, g=g), P)
    bcs = [d.DirichletBC(W, (0.0, 0.0, 0.0), "on_boundary")]
    f = d.Constant((0.0, 0.0, g))
    st = nav.IPCS()
    for _ in range(2):
        u0, p1 = st.step(d.Constant(1e-2), {0: u0}, p0, bcs, [], d.Constant(998.21), d.Constant(1.002e-3), {0: f, 1: f},
                         verbose=False, tol=1e-10)
        p0 = p1
    assert np.sqrt((u0.nodal() ** 2).sum(axis=1)).max() < 1e-13
    pz = p0.vector().get_local() - g * P.tabulate_dof_coordinates()[:, 2]
    assert np.abs(pz - pz.mean()).max() < 1e-9 * abs(g)


def test_cavity_steps_full_size(gpu_ctx, cube):
    """Three lid-driven cavity steps at the benchmark settings: Newton converges by the reference's test, the
    iteration counts stay in the range the benchmark reports, boundary values are reproduced exactly and the
    velocity stays bounded by the lid speed (up to the P2 overshoot at the lid corners)."""
    from flow_b200 import dolfin as d
    from flow_b200 import navier_stokes as nav

    mesh, W, P, _ = cube
    bcs = [d.DirichletBC(W, (0.0, 0.0, 0.0), "on_boundary"), d.DirichletBC(W, (1.0, 0.0, 0.0), lambda x, on: x[2] > 1 - 1e-12)]
    ud, uv = d.collect_bcs(bcs, W)
    zero = d.Constant((0.0, 0.0, 0.0))
    u, p = d.Function(W), d.Function(P)
    st = nav.IPCS()
    for _ in range(3):
        u, p = st.step(d.Constant(1e-2), {0: u}, p, bcs, [], d.Constant(1.0), d.Constant(1e-2), {0: zero, 1: zero}, verbose=False)
        s = nav.last_stats()
        assert s["newton_residual"] < 1e-10 and s["newton_its"] <= 6
        assert s["pressure_its"] < 60 and s["correction_its"] < 80 and s["momentum_its"] < 150
    uvec = u.vector().get_local()
    assert np.isfinite(uvec).all() and np.isfinite(p.vector().get_local()).all()
    assert np.abs(uvec[ud] - uv).max() == 0.0
    un = u.nodal()
    assert un[:, 0].max() <= 1.0 + 1e-9 and np.abs(un).max() <= 1.5
