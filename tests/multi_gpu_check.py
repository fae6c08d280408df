"""Launched with torchrun (N ranks, one GPU each): partitioned device assembly + halo SpMV vs the oracle on the
unpartitioned mesh.  Used by tests/test_gpu_multi.py and by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def field(k, seed):
    return np.sin(0.37 * (k % 1000003) + seed) + 0.1 * np.cos(0.011 * (k % 7919))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import dycore_b200  # noqa: F401
    from dycore_b200 import device, distributed, harness, params
    from oracle import oracle as orc
    refine = int(os.environ.get("DCP_CHECK_REFINE", "2"))
    family = os.environ.get("DCP_CHECK_FAMILY", "classic")
    feec = family == "feec"
    mp_ = params.NAMED["shell_3d_feec" if feec else "shell_3d_classic"]
    assemble_ref = orc.feec_assemble_nse_system if feec else orc.assemble_nse_system
    assemble_tm = orc.feec_assemble_temperature_matrix if feec else orc.assemble_temperature_matrix
    assemble_tr = orc.feec_assemble_temperature_rhs if feec else orc.assemble_temperature_rhs
    P = harness.Problem(geometry="shell", refine=refine, n_ranks=world, rank=rank, family=family)
    keys, owners = P["nse.dof_key"], P["nse.dof_owner"]
    blocks = ("n_w", "n_u", "n_p") if feec else ("n_u", "n_p")
    start, owned, owned_counts = 0, [], []
    for b in blocks:
        owned_counts.append(P.scalar("nse." + b + "_owned"))
        owned.append(start + np.arange(owned_counts[-1]))
        start += P.scalar("nse." + b)
    owned = np.concatenate(owned)
    u = np.ascontiguousarray(field(keys, 1.0) * 0.1)
    T = np.ascontiguousarray(2.0 + 0.2 * field(P["temp.dof_key"], 2.0))
    ctx = device.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    torch.cuda.set_stream(stream)
    model = device.BoussinesqModel.from_problem(ctx, P, mp_)
    model.set_owned(owned_counts, P.scalar("temp.n_owned"))
    halo = distributed.HaloPlan(keys, owners, rank, world, device="cuda")
    d_u, d_T = torch.from_numpy(u).cuda(), torch.from_numpy(T).cuda()
    model.assemble_nse_system(d_u, d_T)
    x = field(keys, 3.0)
    xs = x.copy()
    xs[owners != rank] = 0.0
    d_x = torch.from_numpy(xs).cuda()
    d_y = torch.zeros_like(d_x)
    distributed.DistributedMatrix(model.nse_matrix, halo, ctx).vmult(d_y, d_x)
    ctx.synchronize()
    y = d_y.cpu().numpy()
    # the same product with the ghost exchange hidden behind the interior rows: identical result, owned rows
    d_x2 = torch.from_numpy(xs).cuda()
    d_y2 = torch.full_like(d_x2, float("nan"))
    over = distributed.OverlappedMatrix(model.nse_matrix, halo, local, stream)
    over.vmult(d_y2, d_x2)
    torch.cuda.synchronize()
    y2 = d_y2.cpu().numpy()
    same = bool(np.array_equal(y2[owned], y[owned]))
    over.close()
    # temperature: matrices, right-hand side (advected with the nse vector) and a halo product of temperature_matrix
    tkeys, towners = P["temp.dof_key"], P["temp.dof_owner"]
    n_to = P.scalar("temp.n_owned")
    halo_t = distributed.HaloPlan(tkeys, towners, rank, world, device="cuda")
    model.assemble_temperature_matrix()
    model.assemble_temperature_rhs(d_T, d_u)
    xt = field(tkeys, 5.0)
    xt[towners != rank] = 0.0
    d_xt = torch.from_numpy(xt).cuda()
    d_yt = torch.zeros_like(d_xt)
    distributed.DistributedMatrix(model.temperature_matrix, halo_t, ctx).vmult(d_yt, d_xt)
    ctx.synchronize()
    yt = d_yt.cpu().numpy()
    trhs = model.temperature_rhs
    tgathered = [None] * world
    dist.all_gather_object(tgathered, (tkeys[:n_to], yt[:n_to], trhs[:n_to]))
    rhs = model.nse_rhs
    gathered = [None] * world
    dist.all_gather_object(gathered, (keys[owned], y[owned], rhs[owned]))
    ok = True
    if rank == 0:
        G = harness.Problem(geometry="shell", refine=refine, family=family)
        gk = G["nse.dof_key"]
        prm = orc.params_from(mp_)
        gv, grhs = assemble_ref(G, prm, np.ascontiguousarray(field(gk, 1.0) * 0.1),
                                np.ascontiguousarray(2.0 + 0.2 * field(G["temp.dof_key"], 2.0)))
        grp, gcol, _, _ = G.csr("nse.full")
        yg = orc.spmv(grp, gcol, gv, field(gk, 3.0))
        order = np.argsort(gk)
        kk = np.concatenate([g[0] for g in gathered])
        yy = np.concatenate([g[1] for g in gathered])
        rr = np.concatenate([g[2] for g in gathered])
        pos = np.searchsorted(gk[order], kk)
        err_y = np.abs(yy - yg[order][pos]).max() / np.abs(yg).max()
        err_r = np.abs(rr - grhs[order][pos]).max() / np.abs(grhs).max()
        ok = len(kk) == len(gk) and err_y <= 1e-12 and err_r <= 1e-12
        # temperature reference on the unpartitioned mesh
        gtk = G["temp.dof_key"]
        uT = np.ascontiguousarray(field(gk, 1.0) * 0.1)
        TT = np.ascontiguousarray(2.0 + 0.2 * field(gtk, 2.0))
        rm, rk = assemble_tm(G, prm)
        tm = orc.temperature_matrix_combine(rm, rk, mp_.time_step / mp_.NSE_solver_interval)
        gtr = assemble_tr(G, prm, TT, uT)
        trp, tcol, _, _ = G.csr("temp.pat")
        ytg = orc.spmv(trp, tcol, tm, field(gtk, 5.0))
        torder = np.argsort(gtk)
        tk = np.concatenate([g[0] for g in tgathered])
        ty = np.concatenate([g[1] for g in tgathered])
        tr = np.concatenate([g[2] for g in tgathered])
        tpos = np.searchsorted(gtk[torder], tk)
        err_ty = np.abs(ty - ytg[torder][tpos]).max() / np.abs(ytg).max()
        err_tr = np.abs(tr - gtr[torder][tpos]).max() / np.abs(gtr).max()
        ok = ok and len(tk) == len(gtk) and err_ty <= 1e-12 and err_tr <= 1e-12
        print(f"multi_gpu_check temperature: spmv err {err_ty:.2e}, rhs err {err_tr:.2e}")
        print(f"multi_gpu_check {family} world={world} refine={refine}: spmv err {err_y:.2e}, rhs err {err_r:.2e} -> {'OK' if ok else 'FAIL'}")
    flags = [None] * world
    dist.all_gather_object(flags, same)
    if rank == 0:
        print(f"multi_gpu_check overlapped product identical on all ranks: {all(flags)}")
        ok = ok and all(flags)
    model.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
