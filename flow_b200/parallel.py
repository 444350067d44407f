"""Mesh partitioning for multi-GPU runs (one process per GPU).

Recursive coordinate bisection of the cells, dof ownership = lowest rank among the sharing cells,
owner-computes rows: every rank keeps all cells that touch one of its owned nodes (one ghost-cell
layer), so that owned matrix rows assemble completely without exchanging matrix entries.  Local node
numbering: owned nodes first (ascending global id), then ghost nodes grouped by owner (ascending
global id inside a group) -- a ghost segment is then a contiguous receive buffer.

The reference has no partitioning code of its own (DOLFIN would do it implicitly under mpirun
[EXT]); this module is the replacement for that implicit layer, for the hot path only.  Everything
here is deterministic and computed redundantly on every rank from the global mesh arrays, so the
set-up needs no communication.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib


def rcb(centroids, nparts):
    """Recursive coordinate bisection: part id per cell (balanced to +-1 cell)."""
    part = np.zeros(centroids.shape[0], dtype=np.int32)

    def split(idx, p0, n):
        if n == 1:
            part[idx] = p0
            return
        c = centroids[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        nl = n // 2
        k = (idx.size * nl) // n
        # stable order: coordinate first, cell id as tie break -> deterministic on every rank
        order = np.lexsort((idx, c[:, axis]))
        split(idx[order[:k]], p0, nl)
        split(idx[order[k:]], p0 + nl, n - nl)

    split(np.arange(centroids.shape[0]), 0, nparts)
    return part


class SpacePlan(object):
    """Numbering and halo plan of one Lagrange node set on one rank."""

    def __init__(self, l2g, n_owned, perm, ranks, send_ptr, send_nodes, recv_ptr):
        self.l2g = l2g              # local (space numbering) -> global node id
        self.n_owned = n_owned
        self.perm = perm            # local canonical -> space numbering
        self.ranks, self.send_ptr, self.send_nodes, self.recv_ptr = ranks, send_ptr, send_nodes, recv_ptr


class Partition(object):
    """Everything rank `rank` of `nranks` needs to build its local problem."""

    def __init__(self, gmesh, rank, nranks, part=None, boundary_facet_filter=None):
        """`boundary_facet_filter(vertex coordinates (nf, dim, dim)) -> bool mask` keeps only the facets of gmesh's
        boundary that lie on the boundary of the DOMAIN; needed when gmesh is a cut-out of a larger mesh
        (structured_cube_mesh), whose artificial cut faces are not boundary."""
        pts, cells = gmesh.coordinates(), gmesh.cells()
        self.rank, self.nranks = rank, nranks
        if part is None:
            part = rcb(pts[cells].mean(axis=1), nranks)
        self.cell_part = part
        nv = pts.shape[0]
        g2 = gmesh.node_space(2)
        cn2 = g2.cell_nodes                       # global canonical P2 dof map (vertices, then edges)
        owner2 = np.full(g2.nnodes, nranks, dtype=np.int32)
        np.minimum.at(owner2, cn2.ravel(), np.repeat(part, cn2.shape[1]))
        self.node_owner = {2: owner2, 1: owner2[:nv]}
        own_in_cell = owner2[cn2]
        mask = (own_in_cell == rank).any(axis=1)
        self.local_cells = np.nonzero(mask)[0]
        lcn2 = cn2[self.local_cells]
        lown = own_in_cell[self.local_cells]
        nodes2 = np.unique(lcn2)                  # == local canonical numbering (vertices first)
        nodes1 = nodes2[nodes2 < nv]
        self.points = pts[nodes1]
        self.cells = np.searchsorted(nodes1, cells[self.local_cells]).astype(np.int32)
        # global boundary facets that lie in local cells
        nb = _lib.i64()
        lib.fb_mesh_info(gmesh.handle, None, None, None, C.byref(nb))
        pc, pl = _lib.pi32(), _lib.pi32()
        lib.fb_mesh_boundary_facets(gmesh.handle, C.byref(pc), C.byref(pl))
        bc_ = np.ctypeslib.as_array(pc, shape=(nb.value,))
        bl_ = np.ctypeslib.as_array(pl, shape=(nb.value,))
        g2l_cell = np.full(cells.shape[0], -1, dtype=np.int64)
        g2l_cell[self.local_cells] = np.arange(self.local_cells.size)
        keep = g2l_cell[bc_] >= 0
        if boundary_facet_filter is not None and nb.value:
            nvc = cells.shape[1]
            # facet l of a cell: all its (ascending) vertices but the l-th
            sel = np.arange(nvc)[None, :] != bl_[:, None]
            fverts = cells[bc_][sel].reshape(nb.value, nvc - 1)
            keep &= np.asarray(boundary_facet_filter(pts[fverts]), dtype=bool)
        self.bf_cell = g2l_cell[bc_[keep]].astype(np.int32)
        self.bf_local = bl_[keep].astype(np.int32).copy()
        self.plans = {2: self._plan(nodes2, owner2, lcn2, lown, lambda n: n),
                      1: self._plan(nodes1, owner2, lcn2[:, :cells.shape[1]], lown, lambda n: n)}

    def _plan(self, nodes, owner, lcn, lown_p2, _):
        """nodes: sorted global ids of the local nodes of this space; lcn: local cells -> global nodes of
        this space; lown_p2: owners of the P2 nodes of the local cells (defines which rank holds a cell)."""
        rank = self.rank
        own = owner[nodes]
        owned = nodes[own == rank]
        ghost = nodes[own != rank]
        gown = own[own != rank]
        order = np.lexsort((ghost, gown))
        ghost, gown = ghost[order], gown[order]
        l2g = np.concatenate([owned, ghost])
        # perm: canonical local index (position in `nodes`) -> space numbering
        perm = np.empty(nodes.size, dtype=np.int32)
        perm[np.searchsorted(nodes, l2g)] = np.arange(l2g.size, dtype=np.int32)
        # neighbours: every rank that shares a local cell with us (symmetric relation, taken from the P2
        # ownership so that both sides of a pair agree even when one direction carries no P1 node)
        ranks = np.unique(lown_p2)
        ranks = ranks[ranks != rank].astype(np.int32)
        recv_ptr = np.zeros(ranks.size + 1, dtype=np.int64)
        recv_ptr[1:] = np.cumsum([np.count_nonzero(gown == q) for q in ranks])
        send_ptr = np.zeros(ranks.size + 1, dtype=np.int64)
        send = []
        lcn_owner = owner[lcn]
        for k, q in enumerate(ranks):
            cells_q = (lown_p2 == q).any(axis=1)      # my local cells that rank q also holds
            cand = lcn[cells_q][lcn_owner[cells_q] == rank]
            gl = np.unique(cand)                      # my owned nodes that q sees as ghosts, ascending global id
            send.append(np.searchsorted(owned, gl).astype(np.int32))
            send_ptr[k + 1] = send_ptr[k] + gl.size
        send_nodes = np.concatenate(send) if send else np.zeros(0, np.int32)
        return SpacePlan(l2g, int(owned.size), perm, ranks, send_ptr, send_nodes.astype(np.int32), recv_ptr)


def distributed_mesh(gmesh, rank, nranks, device=None, part=None):
    """Rank-local `dolfin.Mesh` of a partitioned global mesh.  Its node spaces are numbered owned-first
    and carry their halo plans; `mesh.partition` holds the maps back to global ids."""
    from .dolfin import Mesh

    P = Partition(gmesh, rank, nranks, part)
    m = Mesh(P.points, P.cells, device=device)
    _lib.check(lib.fb_mesh_set_boundary_facets(m.handle, P.bf_cell.size, _lib.as_pi32(P.bf_cell), _lib.as_pi32(P.bf_local)),
               m.ctx, "fb_mesh_set_boundary_facets")
    m.partition = P
    m.global_mesh = gmesh
    return m


def cube_blocks(n, nparts):
    """Recursive bisection of the n^3 hexahedra of UnitCubeMesh(n) into `nparts` index boxes aligned with the
    hexahedra (all six tetrahedra of a hexahedron stay together): list of ((i0, j0, k0), (i1, j1, k1)), part p = p-th
    box.  Same recursion as `rcb` (longest extent first, parts split nl : n - nl), but cut on lattice planes, so that
    every rank can name the owner of any cell from its index alone -- no global arrays."""
    out = []

    def split(lo, hi, p0, m):
        if m == 1:
            out.append((p0, lo, hi))
            return
        ext = [hi[a] - lo[a] for a in range(3)]
        axis = int(np.argmax(ext))
        ml = m // 2
        cut = lo[axis] + (ext[axis] * ml + m // 2) // m
        cut = min(max(cut, lo[axis] + 1), hi[axis] - 1) if ext[axis] > 1 else lo[axis]
        hi_l, lo_r = list(hi), list(lo)
        hi_l[axis] = cut
        lo_r[axis] = cut
        split(lo, tuple(hi_l), p0, ml)
        split(tuple(lo_r), hi, p0 + ml, m - ml)

    split((0, 0, 0), (n, n, n), 0, nparts)
    out.sort()
    return [(lo, hi) for _, lo, hi in out]


def cube_cell_part(n, nparts, i, j, k):
    """Part of the hexahedra with lattice indices (i, j, k) under `cube_blocks`."""
    part = np.full(np.shape(i), -1, dtype=np.int32)
    for p, (lo, hi) in enumerate(cube_blocks(n, nparts)):
        inside = (i >= lo[0]) & (i < hi[0]) & (j >= lo[1]) & (j < hi[1]) & (k >= lo[2]) & (k < hi[2])
        part[inside] = p
    return part


def structured_cube_mesh(n, rank, nranks, device=None, layers=2):
    """Rank-local mesh of the partitioned UnitCubeMesh(n) built WITHOUT the global mesh (80 M dofs and beyond: the
    global P2 dof map alone is 0.8 GB per rank).  Every rank generates the cut-out of the lattice that covers its
    block of hexahedra plus `layers` = 2 layers around it -- enough to know every cell that touches one of its nodes
    (1 layer) and every cell that touches one of THOSE cells' nodes (2 layers, needed for the owners of its ghost
    nodes) -- and runs the generic `Partition` on that cut-out with the analytic cell -> part map.  Vertex and edge
    numbers of the cut-out are order-isomorphic to the global ones, so the halo lists of neighbouring ranks, each
    sorted by its own cut-out numbers, line up.  Same result as distributed_mesh(UnitCubeMesh(n), part = the same
    map), array for array (tests/test_parallel.py)."""
    from . import hostfem
    from .dolfin import Mesh

    lo, hi = cube_blocks(n, nranks)[rank]
    e0 = [max(0, lo[a] - layers) for a in range(3)]
    e1 = [min(n, hi[a] + layers) for a in range(3)]
    nx, ny, nz = (e1[a] - e0[a] for a in range(3))
    # the generator on the sub-lattice, with the coordinates of the global generator (0.0 + 1.0 * i / n, bit for bit)
    lat = [0.0 + (1.0 - 0.0) * np.arange(e0[a], e1[a] + 1) / n for a in range(3)]
    pts, cells = hostfem.structured_box_lattice(lat[0], lat[1], lat[2])
    idx = np.arange(nx * ny * nz)
    kk, rem = np.divmod(idx, nx * ny)
    jj, ii = np.divmod(rem, nx)
    part = np.repeat(cube_cell_part(n, nranks, ii + e0[0], jj + e0[1], kk + e0[2]), 6)
    sub = Mesh(pts, cells, device=device)

    def on_domain_boundary(X):  # X: (nf, 3 vertices, 3 coordinates)
        flat0 = (X == 0.0).all(axis=1).any(axis=1)
        flat1 = (X == 1.0).all(axis=1).any(axis=1)
        return flat0 | flat1

    P = Partition(sub, rank, nranks, part=part, boundary_facet_filter=on_domain_boundary)
    m = Mesh(P.points, P.cells, device=device)
    _lib.check(lib.fb_mesh_set_boundary_facets(m.handle, P.bf_cell.size, _lib.as_pi32(P.bf_cell), _lib.as_pi32(P.bf_local)),
               m.ctx, "fb_mesh_set_boundary_facets")
    m.partition = P
    m.global_mesh = None   # no replicated global operators: the pressure AMG is per rank (additive Schwarz)
    m.global_counts = {"cells": 6 * n ** 3, "vertices": (n + 1) ** 3}
    return m


def init_comm(ctx, rank, nranks, broadcast_bytes, allgather_bytes=None, staging_mb_per_peer=8):
    """Create the library's communicator.  `broadcast_bytes(buf)` must broadcast a 128-byte numpy uint8
    array from rank 0 in place (e.g. through torch.distributed).

    If `allgather_bytes(buf) -> (nranks, len(buf)) uint8 array` is given (default: torch.distributed when it is
    initialised) and FB_NO_P2P is unset, the ranks also exchange CUDA IPC handles of their peer-memory windows:
    halo exchanges and dot-product all-reduces then run as the library's own kernels over NVLink peer memory and
    NCCL is only the fallback.  Returns True if the peer-memory transport is active."""
    import os

    ident = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _lib.check(lib.fb_comm_unique_id(ident.ctypes.data_as(C.c_void_p)), ctx, "fb_comm_unique_id")
    ident = broadcast_bytes(ident)
    _lib.check(lib.fb_comm_init(ctx, rank, nranks, ident.ctypes.data_as(C.c_void_p)), ctx, "fb_comm_init")
    if nranks == 1 or nranks > 8 or os.environ.get("FB_NO_P2P", "0") not in ("", "0"):
        return False
    if allgather_bytes is None:
        allgather_bytes = torch_allgather()
        if allgather_bytes is None:
            return False
    handle = np.zeros(64, dtype=np.uint8)
    ok = lib.fb_comm_window_create(ctx, int(staging_mb_per_peer) << 20, handle.ctypes.data_as(C.c_void_p)) == 0
    # every rank takes part in both all-gathers, whatever happened locally: the transport must be the same everywhere
    handles = np.ascontiguousarray(allgather_bytes(handle))
    if ok:
        ok = lib.fb_comm_window_open(ctx, handles.ctypes.data_as(C.c_void_p)) == 0
    all_ok = bool(allgather_bytes(np.array([1 if ok else 0], dtype=np.uint8)).min())
    if not all_ok:
        lib.fb_comm_window_disable(ctx)
    return all_ok


def torch_allgather(device=None):
    """allgather_bytes implementation on torch.distributed; None if no process group is initialised."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return None
    if not dist.is_available() or not dist.is_initialized():
        return None

    def gather(buf):
        t = torch.from_numpy(np.ascontiguousarray(buf).copy())
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return torch.stack(out).cpu().numpy()

    return gather


def torch_broadcast(device=None):
    """broadcast_bytes implementation on torch.distributed (NCCL needs a CUDA tensor, gloo a CPU one)."""
    import torch
    import torch.distributed as dist

    def bcast(buf):
        t = torch.from_numpy(buf.copy())
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        dist.broadcast(t, 0)
        return t.cpu().numpy()

    return bcast
