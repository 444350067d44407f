"""Host-side multi-GPU logic on the CPU: RCB partitioning, ownership, owned-first numbering,
halo plans (consistency between ranks), owner-computes completeness of the local rows, and a
world_size-2 gloo run that moves ghost values with the same plans the NCCL path uses."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import fem, forms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plans(g, R):
    from flow_b200 import parallel

    return [parallel.distributed_mesh(g, r, R) for r in range(R)]


def test_rcb_balanced_and_deterministic():
    from flow_b200 import parallel

    rng = np.random.default_rng(0)
    c = rng.random((1001, 3))
    for R in (2, 3, 4, 8):
        p = parallel.rcb(c, R)
        cnt = np.bincount(p, minlength=R)
        assert cnt.max() - cnt.min() <= 2 and cnt.sum() == 1001
        assert np.array_equal(p, parallel.rcb(c, R))


@pytest.mark.parametrize("kind,R", [("cube", 2), ("cube", 4), ("cube", 8), ("square", 2), ("cube", 3)])
def test_partition_plans_are_consistent(kind, R):
    from flow_b200 import dolfin as d

    g = d.UnitCubeMesh(5, 4, 6) if kind == "cube" else d.UnitSquareMesh(9, 7, "crossed")
    meshes = _plans(g, R)
    for deg in (1, 2):
        gs = g.node_space(deg)
        owned_total = 0
        seen = np.zeros(gs.nnodes, dtype=int)
        for m in meshes:
            ns, pl = m.node_space(deg), m.node_space(deg).plan
            assert np.allclose(ns.coords, gs.coords[pl.l2g])
            # the true domain boundary, not the cut faces of the sub-mesh
            assert np.array_equal(ns.on_boundary, gs.on_boundary[pl.l2g])
            owned_total += pl.n_owned
            seen[pl.l2g[: pl.n_owned]] += 1
            # ghosts are grouped by owner in ascending rank order
            own = m.partition.node_owner[deg][pl.l2g[pl.n_owned:]]
            assert (np.diff(own) >= 0).all() and (own != m.partition.rank).all()
        assert owned_total == gs.nnodes and (seen == 1).all()
        for r, m in enumerate(meshes):
            pr = m.node_space(deg).plan
            for k, q in enumerate(pr.ranks):
                seg = pr.l2g[pr.n_owned + pr.recv_ptr[k]: pr.n_owned + pr.recv_ptr[k + 1]]
                pq = meshes[q].node_space(deg).plan
                kk = list(pq.ranks).index(r)
                sent = pq.l2g[pq.send_nodes[pq.send_ptr[kk]: pq.send_ptr[kk + 1]]]
                assert np.array_equal(seg, sent)


def test_owner_computes_rows_are_complete():
    """Rows of owned nodes assembled from the rank-local cells equal the global rows."""
    from flow_b200 import dolfin as d

    g = d.UnitCubeMesh(4, 3, 3)
    og = fem.Mesh(g.coordinates(), g.cells())
    for deg, form in ((1, forms.stiffness_matrix), (2, forms.mass_matrix)):
        Ag = form(fem.Space(og, deg, 1)).tocsr()
        for m in _plans(g, 3):
            pl = m.node_space(deg).plan
            ol = fem.Mesh(m.coordinates(), m.cells())
            Al = form(fem.Space(ol, deg, 1)).tocsr()  # canonical local numbering
            perm = pl.perm
            # space numbering -> canonical local
            inv = np.empty_like(perm)
            inv[perm] = np.arange(perm.size)
            Al = Al[inv][:, inv]
            owned = np.arange(pl.n_owned)
            sub = Ag[pl.l2g[owned]][:, pl.l2g]
            assert abs(Al[owned] - sub).max() < 1e-14
            # and every column an owned row touches is present locally
            assert abs(Ag[pl.l2g[owned]]).sum() == pytest.approx(abs(sub).sum(), rel=1e-14)


GLOO_SCRIPT = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
from flow_b200 import dolfin as d, parallel
g = d.UnitCubeMesh(4, 4, 3)
m = parallel.distributed_mesh(g, rank, world)
for deg, ncomp in ((1, 1), (2, 3)):
    pl = m.node_space(deg).plan
    n = pl.l2g.size
    truth = np.stack([np.sin(pl.l2g * (c + 1.0)) for c in range(ncomp)], 1)
    x = truth.copy()
    x[pl.n_owned:] = -777.0                       # stale ghosts
    reqs, bufs = [], []
    for k, q in enumerate(pl.ranks):
        s = torch.from_numpy(np.ascontiguousarray(x[pl.send_nodes[pl.send_ptr[k]:pl.send_ptr[k + 1]]]))
        if s.numel():
            reqs.append(dist.isend(s, int(q)))
        r = torch.empty((int(pl.recv_ptr[k + 1] - pl.recv_ptr[k]), ncomp), dtype=torch.float64)
        if r.numel():
            reqs.append(dist.irecv(r, int(q)))
        bufs.append(r)
    for rq in reqs:
        rq.wait()
    for k, r in enumerate(bufs):
        x[pl.n_owned + pl.recv_ptr[k]: pl.n_owned + pl.recv_ptr[k + 1]] = r.numpy()
    assert np.array_equal(x, truth), (rank, deg)
    # global dot product = all-reduced sum over owned entries
    t = torch.tensor([float((truth[:pl.n_owned] ** 2).sum())], dtype=torch.float64)
    dist.all_reduce(t)
    ng = g.node_space(deg).nnodes
    ref = sum(float((np.sin(np.arange(ng) * (c + 1.0)) ** 2).sum()) for c in range(ncomp))
    assert abs(t.item() - ref) < 1e-9 * ref
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_halo_plans_move_ghosts_gloo_world2(tmp_path):
    script = tmp_path / "gloo_halo.py"
    script.write_text(GLOO_SCRIPT % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", FLOW_B200_DEVICE="-1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


@pytest.mark.parametrize("n,R", [(5, 2), (6, 4), (7, 8), (5, 3)])
def test_structured_cube_mesh_equals_generic_partition(n, R):
    """parallel.structured_cube_mesh builds every rank's local problem from a cut-out of the lattice (no global
    mesh); it must be, array for array, what the generic Partition makes of the global mesh with the same cell -> part
    map: same local mesh, same numbering, same halo plans, same boundary facets."""
    from flow_b200 import dolfin as d
    from flow_b200 import parallel

    blocks = parallel.cube_blocks(n, R)
    cover = np.zeros((n, n, n), dtype=int)
    for lo, hi in blocks:
        cover[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
    assert (cover == 1).all() and len(blocks) == R
    sizes = [np.prod([hi[a] - lo[a] for a in range(3)]) for lo, hi in blocks]
    assert max(sizes) <= 1.01 * (-(-n // 2) / (n // 2)) ** 3 * min(sizes) or R == 3  # cuts fall on lattice planes

    g = d.UnitCubeMesh(n, n, n)
    idx = np.arange(n ** 3)
    kk, rem = np.divmod(idx, n * n)
    jj, ii = np.divmod(rem, n)
    part = np.repeat(parallel.cube_cell_part(n, R, ii, jj, kk), 6)
    for r in range(R):
        a = parallel.distributed_mesh(g, r, R, part=part)
        b = parallel.structured_cube_mesh(n, r, R)
        assert np.array_equal(a.coordinates(), b.coordinates()) and np.array_equal(a.cells(), b.cells())
        assert np.array_equal(a.partition.bf_cell, b.partition.bf_cell) and np.array_equal(a.partition.bf_local, b.partition.bf_local)
        for deg in (1, 2):
            pa, pb = a.node_space(deg).plan, b.node_space(deg).plan
            assert pa.n_owned == pb.n_owned
            for name in ("perm", "ranks", "send_ptr", "send_nodes", "recv_ptr"):
                assert np.array_equal(getattr(pa, name), getattr(pb, name)), (r, deg, name)
            assert np.array_equal(a.node_space(deg).on_boundary, b.node_space(deg).on_boundary)
