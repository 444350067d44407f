"""Generate the committed golden fixtures from the oracle (the reference itself cannot run
here: DOLFIN/PETSc are not installable).  Run: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import mms_problems as mp  # noqa: E402
from oracle import fem, forms, navier_stokes as ns, util  # noqa: E402


def make_mesh(spec, n):
    if spec[0] == "unit_square":
        return fem.Mesh(*fem.unit_square_mesh(n, n, spec[1]))
    return fem.Mesh(*fem.rectangle_mesh(spec[1], spec[2], n, n, spec[3]))


def time_errors(problem, factory, mesh_sizes, Dt):
    """Oracle run of compute_time_errors (tests/test_navier_stokes.py:232-383)."""
    pr = problem()
    errs = {"u": np.zeros((len(mesh_sizes), len(Dt))), "p": np.zeros((len(mesh_sizes), len(Dt)))}
    for k, n in enumerate(mesh_sizes):
        mesh = make_mesh(pr["mesh"], n)
        st = factory(mesh)
        W, P = st.W, st.P
        area = mesh.vol.sum()
        for j, dt in enumerate(Dt):
            u0 = util.project(W, pr["u"](0.0), pr["udeg"])
            p0 = util.project(P, pr["p"](0.0), pr["pdeg"])
            bd = W.boundary_dofs()
            g = util.interpolate(W, pr["u"](dt))[bd]
            l0 = forms.expression_load_vector(W, pr["f"](0.0), pr["fdeg"])
            l1 = forms.expression_load_vector(W, pr["f"](dt), pr["fdeg"])
            u1, p1 = st.step(dt, u0, p0, (bd, g), None, pr["rho"], pr["mu"], l0, l1, tol=1e-10)
            errs["u"][k, j] = util.errornorm(W, pr["u"](dt), u1)
            alpha = (util.integrate_expression(mesh, pr["p"](dt), pr["pdeg"]) - util.integrate_function(P, p1)) / area
            errs["p"][k, j] = util.errornorm(P, pr["p"](dt), p1 + alpha)
    return errs


def small_step_fixture():
    """One IPCS and one Rotational step on tiny meshes: full state vectors."""
    out = {}
    for name, mesh in (("tri", fem.Mesh(*fem.unit_square_mesh(3, 3, "crossed"))), ("tet", fem.Mesh(*fem.unit_cube_mesh(2, 2, 2)))):
        st = ns.Rotational(mesh)
        W, P = st.W, st.P
        X = W.node_coords
        u0 = np.stack([np.sin(X[:, (i + 1) % mesh.dim] + 0.3 * i) for i in range(mesh.dim)], 1).reshape(-1)
        p0 = np.cos(P.node_coords[:, 0])
        bd = W.boundary_dofs()
        u1, p1 = st.step(0.1, u0, p0, (bd, u0[bd]), None, 1.1, 0.4, None, None, tol=1e-12)
        out[name] = {"u0": u0.tolist(), "p0": p0.tolist(), "u1": u1.tolist(), "p1": (p1 - p1.mean()).tolist()}
    return out


if __name__ == "__main__":
    e = time_errors(mp.problem_guermond2, ns.IPCS, [8, 16, 32], [1.0, 0.5])
    json.dump({k: v.tolist() for k, v in e.items()}, open(os.path.join(HERE, "mms_ipcs_guermond2.json"), "w"), indent=1)
    json.dump(small_step_fixture(), open(os.path.join(HERE, "rotational_small_step.json"), "w"))
    print("wrote golden fixtures")
