"""flow.stabilization.supg (flow/stabilization.py:13-152): per-cell SUPG parameter

    tau = h^2 / (4 eps p) * (coth(Pe) - 1/Pe) / Pe,   Pe = |b| h / (2 p eps),

with h the element diameter in the direction of the convection b (triangles only, like the
reference's embedded C++ `SupgStab`).  Inside `flow_b200.heat.Heat(..., supg_stabilization=True)` the
parameter is evaluated by the CUDA kernel `k_heat_supg` (`fb_supg_tau` in csrc/fb_element.cuh); this
module exposes the same quantity (evaluated by that device routine through `fb_supg_tau` of the C ABI), mirroring the reference's
`supg(mesh, convection, diffusion, element_degree)` signature."""
import numpy as np

from . import _lib
from ._lib import lib


class SupgStab(object):
    """Degree-1 expression: `vertex_values()` -> (ncells, 3) values of tau at the cell vertices."""

    def __init__(self, mesh, convection, epsilon, p):
        self.mesh, self.convection, self.epsilon, self.p = mesh, convection, float(epsilon), int(p)

    def degree(self):
        return 1

    def vertex_values(self):
        """Evaluated on the device by the routine the heat kernel uses (`fb_supg_tau`); tau > 1e3 raises like the
        reference's `throw 1` (stabilization.py:132-140)."""
        mesh = self.mesh
        assert mesh.dim == 2, "SUPG tau is defined for triangles only (stabilization.py:84-92)"
        W = self.convection.function_space()
        out = np.zeros((mesh.num_cells(), 3))
        _lib.check(lib.fb_supg_tau(W.handle(), _lib.as_pd(_lib.f64(self.convection._vec)), self.epsilon, self.p, _lib.as_pd(out)),
                   mesh.ctx, "supg")
        return out


def supg(mesh, convection, diffusion, element_degree):
    return SupgStab(mesh, convection, diffusion, element_degree)
