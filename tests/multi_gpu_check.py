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


def time_step_check(ctx, comm, model, P, mp_, rank, world, refine):
    """One time step of the named shell config on the partitioned mesh (boussinesq_model.tpp:1867-1905): FGMRES(30) with
    the block Schur preconditioner and the temperature CG, every product through the library's halo exchange, every
    inner product all-reduced -- iteration counts must equal the single-rank CPU mirror's, fields agree to 1e-8."""
    import torch
    import torch.distributed as dist
    import krylov_util as K
    from dycore_b200 import device, distributed, harness
    from dycore_b200 import solvers as S
    n_u, n_p, n_t = P.scalar("nse.n_u"), P.scalar("nse.n_p"), P.scalar("temp.n_dofs")
    ou, op, ot = P.scalar("nse.n_u_owned"), P.scalar("nse.n_p_owned"), P.scalar("temp.n_owned")
    keys, owners = P["nse.dof_key"], P["nse.dof_owner"]
    tkeys, towners = P["temp.dof_key"], P["temp.dof_owner"]
    plan_full = distributed.HaloPlan(keys, owners, rank, world, device="cuda")
    plan_u = distributed.HaloPlan(keys[:n_u], owners[:n_u], rank, world, device="cuda")
    plan_p = distributed.HaloPlan(keys[n_u:], owners[n_u:], rank, world, device="cuda")
    plan_t = distributed.HaloPlan(tkeys, towners, rank, world, device="cuda")
    h_full, h_u, h_p, h_t = (distributed.DeviceHalo(pl, comm) for pl in (plan_full, plan_u, plan_p, plan_t))
    B = distributed.DistributedDeviceBackend(ctx, comm, {n_u + n_p: [(0, ou), (n_u, n_u + op)], n_u: [(0, ou)], n_p: [(0, op)],
                                                         n_t: [(0, ot)]})
    T0 = K.initial_temperature(P, mp_)
    d_u, d_T = torch.zeros(n_u + n_p, dtype=torch.float64, device="cuda"), torch.from_numpy(T0).cuda()
    model.assemble_nse_system(d_u, d_T)
    model.build_nse_preconditioner()
    model.assemble_temperature_matrix()
    model.assemble_temperature_rhs(d_T, d_u)
    rhs = torch.from_numpy(model.nse_rhs).cuda()
    trhs = torch.from_numpy(model.temperature_rhs).cuda()
    A = S.Wrap(distributed.HaloMatrix(model, device.MAT_NSE, h_full, overlap=True))
    halo_of = {0: h_u, 1: h_p}
    blocks = {(i, j): S.Wrap(distributed.HaloBlock(model.nse_matrix.block(i, j), halo_of[j])) for (i, j) in ((0, 0), (0, 1), (1, 0))}
    x, its, inner = S.solve_nse_block_preconditioned(B, A, blocks, S.Wrap(model.Mu_plus_A_preconditioner), rhs, d_u, n_u, n_p,
                                                     mp_.time_step)
    Tm = S.Wrap(distributed.HaloMatrix(model, device.MAT_TEMP, h_t, overlap=True))
    t, cg = S.solve_temperature(B, Tm, S.Wrap(model.T_preconditioner), trhs, d_T)
    xh, th = B.to_numpy(x), B.to_numpy(t)
    own = np.concatenate([np.arange(ou), n_u + np.arange(op)])
    gathered = [None] * world
    dist.all_gather_object(gathered, (keys[own], xh[own], tkeys[:ot], th[:ot], its, list(inner), cg))
    ok = True
    if rank == 0:
        G = harness.Problem(geometry="shell", refine=refine)
        ref = K.cpu_time_step(G, mp_, np.zeros(G.scalar("nse.n_dofs")), K.initial_temperature(G, mp_))
        counts = {(g[4], tuple(g[5]), g[6]) for g in gathered}
        ok = len(counts) == 1 and its == ref["fgmres"] and list(inner) == list(ref["inner"]) and cg == ref["cg"]
        # fields before constraints.distribute / pressure rescaling, on free dofs: compare through the solved system
        gk, gtk = G["nse.dof_key"], G["temp.dof_key"]
        gn_u = G.scalar("nse.n_u")
        kk = np.concatenate([g[0] for g in gathered])
        vv = np.concatenate([g[1] for g in gathered])
        full = np.zeros(len(gk))
        full[np.argsort(gk)[np.searchsorted(np.sort(gk), kk)]] = vv
        full = S.distribute(S.NumpyBackend(), K.cs_lines(G, "nse.cs"), full)
        full[gn_u:] /= mp_.time_step
        tk = np.concatenate([g[2] for g in gathered])
        tv = np.concatenate([g[3] for g in gathered])
        tfull = np.zeros(len(gtk))
        tfull[np.argsort(gtk)[np.searchsorted(np.sort(gtk), tk)]] = tv
        tfull = S.distribute(S.NumpyBackend(), K.cs_lines(G, "temp.cs"), tfull)
        errs = [np.abs(full[:gn_u] - ref["nse"][:gn_u]).max() / np.abs(ref["nse"][:gn_u]).max(),
                np.abs(full[gn_u:] - ref["nse"][gn_u:]).max() / np.abs(ref["nse"][gn_u:]).max(),
                np.abs(tfull - ref["temp"]).max() / np.abs(ref["temp"]).max()]
        ok = ok and max(errs) <= 1e-8
        print(f"multi_gpu_check time step world={world}: fgmres {its} (cpu {ref['fgmres']}), inner {list(inner)} (cpu {list(ref['inner'])}), "
              f"cg {cg} (cpu {ref['cg']}), field errors {errs[0]:.1e} {errs[1]:.1e} {errs[2]:.1e} -> {'OK' if ok else 'FAIL'}")
    for h in (h_full, h_u, h_p, h_t):
        h.close()
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    return bool(flag[0])


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
    # ---- the library's own data plane (dcp_comm_* / dcp_halo_*): same products, plain and overlapped, and an
    # all-reduced inner product
    comm = distributed.Communicator(ctx, rank, world)
    dh = distributed.DeviceHalo(halo, comm)
    lib_same = True
    for overlap in (False, True):
        d_x3 = torch.from_numpy(xs).cuda()
        d_y3 = torch.full_like(d_x3, float("nan"))
        dh.vmult(model, device.MAT_NSE, d_y3, d_x3, overlap=overlap)
        ctx.synchronize()
        torch.cuda.synchronize()
        lib_same = lib_same and bool(np.array_equal(d_y3.cpu().numpy()[owned], y[owned]))
    dh_t = distributed.DeviceHalo(halo_t, comm)
    d_xt3 = torch.from_numpy(xt).cuda()
    d_yt3 = torch.full_like(d_xt3, float("nan"))
    dh_t.vmult(model, device.MAT_TEMP, d_yt3, d_xt3, overlap=True)
    ctx.synchronize()
    torch.cuda.synchronize()
    lib_same = lib_same and bool(np.array_equal(d_yt3.cpu().numpy()[:n_to], yt[:n_to]))
    ranges, start2 = [], 0
    for b, cnt in zip(blocks, owned_counts):
        ranges.append((start2, start2 + cnt))
        start2 += P.scalar("nse." + b)
    dsum = comm.dot(d_y, d_x, ranges)
    parts = [None] * world
    dist.all_gather_object(parts, float(np.dot(y[owned], xs[owned])))
    lib_same = lib_same and abs(dsum - sum(parts)) <= 1e-12 * max(1.0, abs(sum(parts)))
    mx = comm.max([float(rank), -float(rank)])
    lib_same = lib_same and mx == [float(world - 1), 0.0]
    step_ok = True
    if not feec:
        step_ok = time_step_check(ctx, comm, model, P, mp_, rank, world, refine)
    dh.close()
    dh_t.close()
    flags = [None] * world
    dist.all_gather_object(flags, same and lib_same and step_ok)
    if rank == 0:
        print(f"multi_gpu_check overlapped / library-halo products, all-reduced dot and the partitioned time step agree on all ranks: {all(flags)}")
        ok = ok and all(flags)
    comm.close()
    model.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
