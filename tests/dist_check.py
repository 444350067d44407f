"""Multi-GPU parity check, launched with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 tests/dist_check.py [n]

Every rank first runs the lid-driven cavity on the GLOBAL mesh on its own GPU (single-rank path),
then the partitioned run (RCB + halo exchange + all-reduced dots over NCCL) and compares its owned
dofs with the global solution: 1e-8 relative L2 (pressure modulo its mean), same Newton counts."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["FLOW_B200_DEVICE"] = str(local)
    from flow_b200 import _lib, dolfin as d, navier_stokes as nav, parallel

    # solver options as they ship: both runs follow the reference's Newton iterates (include/flowb200.h, jacobian_reuse)
    nav.reset_options()

    def cavity(mesh, steps, scheme):
        W = d.VectorFunctionSpace(mesh, "CG", 2)
        P = d.FunctionSpace(mesh, "CG", 1)
        bcs = [d.DirichletBC(W, (0.0, 0.0, 0.0), "on_boundary"),
               d.DirichletBC(W, d.Expression(("4*x[0]*(1-x[0])", "0.0", "0.0"), degree=2), lambda x, on: x[2] > 1 - 1e-12)]
        u, p = d.Function(W), d.Function(P)
        f = {0: d.Constant((0.0, 0.1, -1.0)), 1: d.Constant((0.0, 0.1, -1.0))}
        hist = []
        for _ in range(steps):
            u, p = scheme.step(d.Constant(0.02), {0: u}, p, bcs, [], d.Constant(1.0), d.Constant(0.02), f, verbose=False, tol=1e-11)
            hist.append(nav.last_stats())
        return W, P, u, p, hist

    g = d.UnitCubeMesh(n, n, n)
    results = {}
    for name, scheme in (("ipcs", nav.IPCS()), ("rotational_cn", nav.Rotational("crank-nicolson"))):
        _, _, ug, pg, hg = cavity(g, 2, scheme)
        results[name] = (ug._vec.copy(), pg._vec.copy(), hg)

    ctx = _lib.context()
    p2p = parallel.init_comm(ctx, rank, world, parallel.torch_broadcast(local))
    if rank == 0:
        print("dist_check transport: %s" % ("peer memory (NVLink windows)" if p2p else "NCCL"), flush=True)
    m = parallel.distributed_mesh(g, rank, world)
    ok = True
    for name, scheme in (("ipcs", nav.IPCS()), ("rotational_cn", nav.Rotational("crank-nicolson"))):
        W, P, u, p, h = cavity(m, 2, scheme)
        ug, pg, hg = results[name]
        plu, plp = W.nodes.plan, P.nodes.plan
        ul = u._vec.reshape(-1, 3)
        ref_u = ug.reshape(-1, 3)[plu.l2g]
        ref_p = pg[plp.l2g]
        # owned AND ghost copies are current on return
        num = torch.tensor([float(((ul[: plu.n_owned] - ref_u[: plu.n_owned]) ** 2).sum()), float((ref_u[: plu.n_owned] ** 2).sum()),
                            float(p._vec[: plp.n_owned].sum()), float(plp.n_owned)], dtype=torch.float64, device="cuda")
        dist.all_reduce(num)
        eu = (num[0] / num[1]).sqrt().item()
        pmean = (num[2] / num[3]).item()
        dp = (p._vec[: plp.n_owned] - pmean) - (ref_p[: plp.n_owned] - pg.mean())
        nump = torch.tensor([float((dp ** 2).sum()), float(((ref_p[: plp.n_owned] - pg.mean()) ** 2).sum())], dtype=torch.float64, device="cuda")
        dist.all_reduce(nump)
        ep = (nump[0] / nump[1]).sqrt().item()
        ghost_err = float(np.abs(ul[plu.n_owned:] - ref_u[plu.n_owned:]).max()) if plu.n_owned < ul.shape[0] else 0.0
        its = [(a["newton_its"], a["momentum_its"], a["pressure_its"], a["correction_its"]) for a in h]
        its_g = [(a["newton_its"], a["momentum_its"], a["pressure_its"], a["correction_its"]) for a in hg]
        good = eu < 1e-8 and ep < 1e-7 and ghost_err < 1e-8
        ok = ok and good
        if rank == 0:
            print("dist_check %s world=%d n=%d: |u-u1gpu|/|u| = %.2e  |p-p1gpu|/|p| = %.2e  ghost max err = %.1e  its(dist)=%s its(1gpu)=%s %s"
                  % (name, world, n, eu, ep, ghost_err, its, its_g, "OK" if good else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
