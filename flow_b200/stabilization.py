"""flow.stabilization.supg (flow/stabilization.py:13-152): per-cell SUPG parameter

    tau = h^2 / (4 eps p) * (coth(Pe) - 1/Pe) / Pe,   Pe = |b| h / (2 p eps),

with h the element diameter in the direction of the convection b (triangles only, like the
reference's embedded C++ `SupgStab`).  Inside `flow_b200.heat.Heat(..., supg_stabilization=True)` the
parameter is evaluated by the CUDA kernel `k_heat_supg` (`fb_supg_tau` in csrc/fb_element.cuh); this
module exposes the same quantity as a host-side object for inspection, mirroring the reference's
`supg(mesh, convection, diffusion, element_degree)` signature."""
import numpy as np


class SupgStab(object):
    """Degree-1 expression: `vertex_values()` -> (ncells, 3) values of tau at the cell vertices."""

    def __init__(self, mesh, convection, epsilon, p):
        self.mesh, self.convection, self.epsilon, self.p = mesh, convection, float(epsilon), int(p)

    def degree(self):
        return 1

    def vertex_values(self):
        mesh = self.mesh
        assert mesh.dim == 2, "SUPG tau is defined for triangles only (stabilization.py:84-92)"
        W = self.convection.function_space()
        X = mesh.coordinates()[mesh.cells()]
        V = self.convection.nodal()[W.nodes.cell_nodes[:, :3]]
        nrm = np.sqrt((V ** 2).sum(axis=2))
        s = np.zeros_like(nrm)
        for i in range(3):
            for j in range(i + 1, 3):
                e = X[:, i, :] - X[:, j, :]
                s += np.abs(e[:, None, 1] * V[:, :, 0] - e[:, None, 0] * V[:, :, 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            h = 4.0 * nrm * mesh.volumes()[:, None] / s
            Pe = 0.5 * nrm * h / (self.p * self.epsilon)
            xi = np.where(Pe > 1.0e-5, (1.0 / np.tanh(Pe) - 1.0 / Pe) / Pe, 1.0 / 3.0 - Pe ** 2 / 45.0 + 2.0 / 945.0 * Pe ** 4)
            tau = h * h / 4.0 / self.epsilon / self.p * xi
        tau = np.where(nrm < 1.0e-10, 0.0, tau)
        if (tau > 1.0e3).any():
            raise RuntimeError("SUPG tau > 1e3 (stabilization.py:132-140)")
        return tau


def supg(mesh, convection, diffusion, element_degree):
    return SupgStab(mesh, convection, diffusion, element_degree)
