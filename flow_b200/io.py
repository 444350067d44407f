"""Result output for the driver loops: the reference's drivers write every accepted state with
``XDMFFile(mpi_comm_world(), 'x.xdmf').write(u, t)`` (tests/test_sealed_box.py:102-113,
tests/test_boussinesq.py:164, 307-309, tests/test_karman_vortex_street.py:214-222).  HDF5 is not available
here, so the same calls write a ParaView collection instead: ``<stem>.pvd`` + one ASCII ``<stem>_NNNNNN.vtu``
per time step, with quadratic (P2) or linear (P1) Lagrange cells.  Host-side I/O only -- no arithmetic.
"""
import os

import numpy as np

# UFC local P2 node order (vertices, then edges (1,2),(0,2),(0,1) / (2,3),(1,3),(1,2),(0,3),(0,2),(0,1)) ->
# VTK quadratic cell order (vertices, then mid-edge nodes (0,1),(1,2),(2,0)[,(0,3),(1,3),(2,3)])
_VTK = {(2, 1): (5, (0, 1, 2)), (3, 1): (10, (0, 1, 2, 3)),
        (2, 2): (22, (0, 1, 2, 5, 3, 4)), (3, 2): (24, (0, 1, 2, 3, 9, 6, 8, 7, 5, 4))}


def mpi_comm_world():
    """Placeholder for dolfin.mpi_comm_world() (only ever passed to XDMFFile by the reference's drivers)."""
    return None


def _fmt(a, per_line):
    a = np.asarray(a)
    flat = a.reshape(-1)
    rows = [" ".join(repr(x) if isinstance(x, float) else str(x) for x in flat[i:i + per_line].tolist())
            for i in range(0, flat.size, per_line)]
    return "\n".join(rows)


def write_vtu(path, functions):
    """One .vtu file with all `functions` (flow_b200.dolfin.Function objects on the same mesh and of the same
    Lagrange degree) as point data."""
    functions = list(functions) if isinstance(functions, (list, tuple)) else [functions]
    V = functions[0].function_space()
    ns, mesh = V.nodes, V.mesh()
    dim, degree = mesh.dim, ns.degree
    for f in functions:
        W = f.function_space()
        if W.mesh() is not mesh or W.nodes.degree != degree:
            raise ValueError("write_vtu: all functions must live on the same mesh with the same degree")
    ctype, order = _VTK[(dim, degree)]
    pts = np.zeros((ns.nnodes, 3))
    pts[:, :dim] = ns.coords
    conn = np.asarray(ns.cell_nodes)[:, list(order)]
    nc, npc = conn.shape
    out = ['<?xml version="1.0"?>',
           '<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">',
           "<UnstructuredGrid>",
           '<Piece NumberOfPoints="%d" NumberOfCells="%d">' % (ns.nnodes, nc),
           '<Points><DataArray type="Float64" NumberOfComponents="3" format="ascii">', _fmt(pts, 3), "</DataArray></Points>",
           "<Cells>",
           '<DataArray type="Int32" Name="connectivity" format="ascii">', _fmt(conn, npc), "</DataArray>",
           '<DataArray type="Int32" Name="offsets" format="ascii">', _fmt(np.arange(1, nc + 1) * npc, 16), "</DataArray>",
           '<DataArray type="UInt8" Name="types" format="ascii">', _fmt(np.full(nc, ctype), 32), "</DataArray>",
           "</Cells>", "<PointData>"]
    for f in functions:
        vals = f.nodal()
        ncomp = vals.shape[1]
        if ncomp > 1:  # VTK vectors have three components
            v3 = np.zeros((vals.shape[0], 3))
            v3[:, :ncomp] = vals
            vals, ncomp = v3, 3
        out += ['<DataArray type="Float64" Name="%s" NumberOfComponents="%d" format="ascii">' % (f.name(), ncomp),
                _fmt(vals, ncomp), "</DataArray>"]
    out += ["</PointData>", "</Piece>", "</UnstructuredGrid>", "</VTKFile>", ""]
    with open(path, "w") as fh:
        fh.write("\n".join(out))


class XDMFFile(object):
    """``XDMFFile(comm, filename).write(u, t)`` as in the reference's drivers; writes ``<stem>.pvd`` and
    ``<stem>_NNNNNN.vtu`` (see the module docstring).  ``parameters`` accepts the keys the drivers set
    (flush_output, rewrite_function_mesh) and ignores them."""

    def __init__(self, comm_or_filename, filename=None):
        name = filename if filename is not None else comm_or_filename
        self.stem = os.path.splitext(str(name))[0]
        self.parameters = {}
        self._steps = []

    def write(self, f, t=None):
        k = len(self._steps)
        path = "%s_%06d.vtu" % (self.stem, k)
        write_vtu(path, f)
        self._steps.append((float(k) if t is None else float(t), os.path.basename(path)))
        lines = ['<?xml version="1.0"?>', '<VTKFile type="Collection" version="0.1">', "<Collection>"]
        lines += ['<DataSet timestep="%r" part="0" file="%s"/>' % s for s in self._steps]
        lines += ["</Collection>", "</VTKFile>", ""]
        with open(self.stem + ".pvd", "w") as fh:
            fh.write("\n".join(lines))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass


class File(XDMFFile):
    """dolfin's ``File('u.pvd') << u``."""

    def __lshift__(self, f):
        if isinstance(f, tuple):
            self.write(f[0], f[1])
        else:
            self.write(f)
        return self
