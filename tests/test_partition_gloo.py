"""N>1 path on CPU (gloo, world_size 2): the partitioned problem must reproduce the single-rank operator.

Each rank builds its sub-problem (owned chunk of the Morton curve + one ghost-cell layer), assembles its local
matrices with the oracle (ghost-cell-redundant assembly, no compress), refreshes ghost entries of a source
vector through HaloPlan (grouped point-to-point, the same code path NCCL runs on GPUs) and applies its owned
rows.  Rank 0 compares the gathered result with the operator of the unpartitioned mesh.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, refine, out, family="classic"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dycore_b200  # noqa: F401
        from dycore_b200 import distributed, harness, params
        from oracle import oracle as orc
        feec = family == "feec"
        mp_ = params.NAMED["shell_3d_feec" if feec else "shell_3d_classic"]
        prm = orc.params_from(mp_)
        assemble = orc.feec_assemble_nse_system if feec else orc.assemble_nse_system
        P = harness.Problem(geometry="shell", refine=refine, n_ranks=world, rank=rank, family=family)
        n = P.scalar("nse.n_dofs")
        keys, owners = P["nse.dof_key"], P["nse.dof_owner"]
        # owned dofs come first inside every block (Trilinos-style local numbering)
        blocks = ("n_w", "n_u", "n_p") if feec else ("n_u", "n_p")
        start, owned = 0, []
        for b in blocks:
            owned.append(start + np.arange(P.scalar("nse." + b + "_owned")))
            start += P.scalar("nse." + b)
        owned = np.concatenate(owned)
        assert (owners[owned] == rank).all() and (np.delete(owners, owned) != rank).all()

        # fields defined through the global key so that every rank sees the same function
        def field(k, seed):
            return np.sin(0.37 * (k % 1000003) + seed) + 0.1 * np.cos(0.011 * (k % 7919))
        u = field(keys, 1.0) * 0.1
        tkeys = P["temp.dof_key"]
        T = 2.0 + 0.2 * field(tkeys, 2.0)
        vals, rhs = assemble(P, prm, np.ascontiguousarray(u), np.ascontiguousarray(T))
        rp, col, _, _ = P.csr("nse.full")

        halo = distributed.HaloPlan(keys, owners, rank, world)
        x = torch.from_numpy(field(keys, 3.0).copy())
        x_ref = x.clone()
        ghost = np.flatnonzero(owners != rank)
        x[torch.from_numpy(ghost)] = 0.0          # ghosts are stale before the exchange
        halo.exchange(x)
        assert torch.equal(x, x_ref), "halo exchange did not reproduce the owner's values"
        y = orc.spmv(rp, col, vals, x.numpy())
        gathered = [None] * world
        dist.all_gather_object(gathered, (keys[owned], y[owned], rhs[owned]))
        if rank == 0:
            G = harness.Problem(geometry="shell", refine=refine, family=family)
            gk = G["nse.dof_key"]
            ug = field(gk, 1.0) * 0.1
            Tg = 2.0 + 0.2 * field(G["temp.dof_key"], 2.0)
            gv, grhs = assemble(G, prm, np.ascontiguousarray(ug), np.ascontiguousarray(Tg))
            grp, gcol, _, _ = G.csr("nse.full")
            yg = orc.spmv(grp, gcol, gv, field(gk, 3.0))
            order = np.argsort(gk)
            kk = np.concatenate([g[0] for g in gathered])
            yy = np.concatenate([g[1] for g in gathered])
            rr = np.concatenate([g[2] for g in gathered])
            assert len(kk) == len(gk) and len(np.unique(kk)) == len(kk), "owned sets do not tile the global dofs"
            pos = np.searchsorted(gk[order], kk)
            assert (gk[order][pos] == kk).all()
            err_y = np.abs(yy - yg[order][pos]).max() / np.abs(yg).max()
            err_r = np.abs(rr - grhs[order][pos]).max() / np.abs(grhs).max()
            out.put((err_y, err_r))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("refine,family", [(1, "classic"), (2, "classic"), (2, "feec")])
def test_two_ranks_reproduce_single_rank_operator(refine, family):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + refine + (7 if family == "feec" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, refine, out, family)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    err_y, err_r = out.get(timeout=10)
    assert err_y <= 1e-12 and err_r <= 1e-12, (err_y, err_r)
