"""At-size parity fixtures: consecutive IPCS steps of the ORACLE (reference settings: Newton |F|_2 < 1e-10 with
exact linear solves, CG rtol = tol; pressure_correction.py:228-236, :414-432, :451-464) on meshes too large for
the oracle to run inside the GPU test tier.  The reference itself (DOLFIN/PETSc) cannot run here.

  python tests/golden/make_parity_fixtures.py cube24 [cube32] [cavity2d_128] [cavity2d_333]

Each case writes tests/golden/parity_<case>.npz holding, for the stored steps, the values of a seeded random sample
of velocity and pressure dofs (canonical numbering of oracle/fem.py == numbering of the C ABI), the full-vector
norms |u|_2 and |p - mean p|_2, and the Newton residual histories.  tests/test_gpu_parity_default.py reproduces
them on the GPU at DEFAULT solver options."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import fem, navier_stokes as ons  # noqa: E402

DT, RHO, MU, TOL = 1.0e-2, 1.0, 1.0e-2, 1.0e-10

CASES = {
    # name: (dim, n, linear, steps, stored steps)
    "cube24": (3, 24, "tight-krylov", 10, (1, 2, 5, 10)),
    "cube32": (3, 32, "tight-krylov", 10, (1, 2, 5, 10)),
    "cavity2d_128": (2, 128, "lu", 10, (1, 2, 5, 10)),
    "cavity2d_333": (2, 333, "lu", 3, (1, 2, 3)),
}


def cavity(dim, n, linear):
    """Lid-driven cavity of bench.py (3D: lid u = (1,0,0) on z = 1) / SURVEY.md 8d config 2 (2D: u = (1,0) on y = 1,
    unregularised, 'right' diagonals); no-slip elsewhere, the lid wins on the shared edges; p_bcs = []."""
    if dim == 3:
        om = fem.Mesh(*fem.unit_cube_mesh(n, n, n))
    else:
        om = fem.Mesh(*fem.unit_square_mesh(n, n, "right"))
    st = ons.IPCS(om, linear=linear)
    g = np.zeros((st.W.nnodes, dim))
    g[st.W.node_coords[:, dim - 1] > 1.0 - 1e-12, 0] = 1.0
    bd = st.W.boundary_dofs()
    return st, (bd, g.reshape(-1)[bd])


def make(case):
    dim, n, linear, steps, stored = CASES[case]
    t0 = time.time()
    st, bc = cavity(dim, n, linear)
    rng = np.random.default_rng(20261018)
    iu = np.sort(rng.choice(st.W.ndofs, size=min(8000, st.W.ndofs), replace=False))
    ip = np.sort(rng.choice(st.P.nnodes, size=min(4000, st.P.nnodes), replace=False))
    out = {"n": n, "dim": dim, "dt": DT, "rho": RHO, "mu": MU, "tol": TOL, "steps": np.array(stored), "iu": iu, "ip": ip,
           "ndofs_u": st.W.ndofs, "ndofs_p": st.P.nnodes, "linear": linear}
    u, p = np.zeros(st.W.ndofs), np.zeros(st.P.nnodes)
    hist = []
    for k in range(1, steps + 1):
        u, p = st.step(DT, u, p, bc, None, RHO, MU, None, None, tol=TOL)
        hist.append(list(st.info["newton_residuals"]) + [np.nan] * (8 - len(st.info["newton_residuals"])))
        print("%s step %d: newton %s  (%.0f s)" % (case, k, ["%.1e" % r for r in st.info["newton_residuals"]], time.time() - t0),
              flush=True)
        if k in stored:
            out["u_%d" % k] = u[iu].copy()
            out["p_%d" % k] = (p - p.mean())[ip].copy()
            out["unorm_%d" % k] = np.linalg.norm(u)
            out["pnorm_%d" % k] = np.linalg.norm(p - p.mean())
    out["newton_residuals"] = np.array(hist)
    np.savez_compressed(os.path.join(HERE, "parity_%s.npz" % case), **out)
    print("wrote parity_%s.npz in %.0f s" % (case, time.time() - t0))


if __name__ == "__main__":
    for c in sys.argv[1:] or ["cube24"]:
        make(c)
