"""XDMFFile / File stand-ins (flow_b200/io.py): the calls of the reference's drivers (tests/test_sealed_box.py:102-113)
produce a ParaView collection whose quadratic cells carry the P2 nodes in VTK order."""
import xml.etree.ElementTree as ET

import numpy as np
import pytest


def _arrays(path):
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    out = {"npoints": int(piece.get("NumberOfPoints")), "ncells": int(piece.get("NumberOfCells"))}
    out["points"] = np.array(piece.find("Points/DataArray").text.split(), dtype=float).reshape(-1, 3)
    for da in piece.findall("Cells/DataArray"):
        out[da.get("Name")] = np.array(da.text.split(), dtype=int)
    for da in piece.findall("PointData/DataArray"):
        out["data:" + da.get("Name")] = np.array(da.text.split(), dtype=float).reshape(out["npoints"], -1)
    return out


@pytest.mark.parametrize("dim", [2, 3])
def test_xdmf_standin_writes_quadratic_cells(tmp_path, dim):
    from flow_b200 import dolfin as d

    mesh = d.UnitSquareMesh(3, 2, "crossed") if dim == 2 else d.UnitCubeMesh(2, 1, 2)
    W = d.VectorFunctionSpace(mesh, "CG", 2)
    P = d.FunctionSpace(mesh, "CG", 1)
    u = d.interpolate(d.Expression(("x[0]*x[1]", "1.0-x[0]") if dim == 2 else ("x[0]*x[1]", "1.0-x[0]", "x[2]*x[2]"), degree=2), W)
    p = d.interpolate(d.Expression("2.0*x[0]-x[1]", degree=1), P)
    u.rename("u", "velocity")
    p.rename("p", "pressure")
    with d.XDMFFile(d.mpi_comm_world(), str(tmp_path / "u.xdmf")) as uf:
        uf.parameters["flush_output"] = True
        uf.write(u, 0.0)
        uf.write(u, 0.5)
    d.File(str(tmp_path / "p.pvd")) << p
    pvd = ET.parse(tmp_path / "u.pvd").getroot()
    sets = pvd.findall("Collection/DataSet")
    assert [s.get("timestep") for s in sets] == ["0.0", "0.5"] and sets[1].get("file") == "u_000001.vtu"
    a = _arrays(tmp_path / "u_000001.vtu")
    nl = 6 if dim == 2 else 10
    nv = dim + 1
    assert a["npoints"] == W.nodes.nnodes and a["ncells"] == mesh.num_cells()
    assert set(a["types"]) == {22 if dim == 2 else 24} and np.array_equal(a["offsets"], nl * np.arange(1, a["ncells"] + 1))
    conn = a["connectivity"].reshape(-1, nl)
    X = a["points"]
    edges = [(0, 1), (1, 2), (2, 0)] + ([(0, 3), (1, 3), (2, 3)] if dim == 3 else [])
    for k, (i, j) in enumerate(edges):  # VTK: mid-edge node nv+k sits between vertices i and j
        assert np.abs(X[conn[:, nv + k]] - 0.5 * (X[conn[:, i]] + X[conn[:, j]])).max() < 1e-14
    U = a["data:u"]
    assert U.shape == (a["npoints"], 3)
    assert np.abs(U[:, 0] - X[:, 0] * X[:, 1]).max() < 1e-14 and np.abs(U[:, 1] - (1.0 - X[:, 0])).max() < 1e-14
    if dim == 3:
        assert np.abs(U[:, 2] - X[:, 2] ** 2).max() < 1e-14
    b = _arrays(tmp_path / "p_000000.vtu")
    assert set(b["types"]) == {5 if dim == 2 else 10} and b["npoints"] == P.nodes.nnodes
    assert np.abs(b["data:p"][:, 0] - (2.0 * b["points"][:, 0] - b["points"][:, 1])).max() < 1e-14
