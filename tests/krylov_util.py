"""Helpers for the solver-level parity tests: one Boussinesq time step on either backend."""
import math

import numpy as np
import scipy.sparse as sp


def initial_temperature(P, mp):
    """TemperatureInitialValues<3> (boussinesq_model_data.tpp:92-147) interpolated at the temperature support points
    (the reference projects it, boussinesq_model.tpp:1802-1829; the input vector is the same for both backends)."""
    dim = P.dim
    x = P["temp.dof_xyz"].reshape(-1, dim)
    R0, R1 = mp.R0_scaled, mp.R1_scaled
    cov = 20.0 / ((R1 - R0) / 2.0)
    c1 = np.zeros(dim)
    c2 = np.zeros(dim)
    c1[0] = R0 + (R1 - R0) * 0.35
    c2[1] = R0 + (R1 - R0) * 0.65
    nrm = math.sqrt((2 * math.pi) ** dim)
    det = cov ** dim
    q1 = cov * ((x - c1) ** 2).sum(axis=1)
    q2 = cov * ((x - c2) ** 2).sum(axis=1)
    return np.ascontiguousarray(math.sqrt(det) * (np.exp(-0.5 * q1) + np.exp(-0.5 * q2)) / nrm)


def cs_lines(P, prefix):
    return (P[prefix + ".line_dof"], P[prefix + ".line_ptr"], P[prefix + ".entry_dof"], P[prefix + ".entry_w"],
            P[prefix + ".inhom"])


def cpu_time_step(P, mp, u0, T0):
    """Reference-order time step (boussinesq_model.tpp:1867-1905) on the CPU: oracle assembly + numpy solvers."""
    from dycore_b200 import solvers as S
    from oracle import oracle as orc
    prm = orc.params_from(mp)
    B = S.NumpyBackend()
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    n_u, n_p = P.scalar("nse.n_u"), P.scalar("nse.n_p")
    vals, rhs = orc.assemble_nse_system(P, prm, u0, T0)
    pv = orc.assemble_nse_preconditioner(P, prm)
    m, k = orc.assemble_temperature_matrix(P, prm)
    tm = orc.temperature_matrix_combine(m, k, mp.time_step / mp.NSE_solver_interval)
    trhs = orc.assemble_temperature_rhs(P, prm, T0, u0)
    rp, col, _, _ = P.csr("nse.full")
    A = sp.csr_matrix((vals, col, rp), shape=(n, n))
    rp, col, _, _ = P.csr("pre.full")
    Pm = sp.csr_matrix((pv, col, rp), shape=(n, n))
    rp, col, _, _ = P.csr("temp.pat")
    Tm = sp.csr_matrix((tm, col, rp), shape=(nT, nT))

    def mat(M):
        return S.Wrap(lambda dst, src, M=M: dst.__setitem__(slice(None), M @ src))

    def jac(M):
        d = M.diagonal()
        dinv = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 0.0)
        return S.Wrap(lambda dst, src, dinv=dinv: dst.__setitem__(slice(None), dinv * src))
    blocks = {(0, 0): mat(A[:n_u, :n_u]), (0, 1): mat(A[:n_u, n_u:]), (1, 0): mat(A[n_u:, :n_u])}
    x, its, inner = S.solve_nse_block_preconditioned(B, mat(A), blocks, jac(Pm[:n_u, :n_u]), rhs, u0, n_u, n_p, mp.time_step)
    x = S.distribute(B, cs_lines(P, "nse.cs"), x)
    x[n_u:] /= mp.time_step
    t, cg = S.solve_temperature(B, mat(Tm), jac(Tm), trhs, T0)
    t = S.distribute(B, cs_lines(P, "temp.cs"), t)
    return dict(nse=x, temp=t, fgmres=its, inner=inner, cg=cg)


def gpu_time_step(ctx, P, mp, u0, T0):
    """The same step with every operator on the device; vectors stay in HBM until the final download."""
    import torch
    from dycore_b200 import device
    from dycore_b200 import solvers as S
    B = S.DeviceBackend(ctx)
    n_u, n_p = P.scalar("nse.n_u"), P.scalar("nse.n_p")
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    d_u, d_T = torch.from_numpy(u0).cuda(), torch.from_numpy(T0).cuda()
    torch.cuda.synchronize()
    model.assemble_nse_system(d_u, d_T)
    model.build_nse_preconditioner()
    model.assemble_temperature_matrix()
    model.assemble_temperature_rhs(d_T, d_u)
    ptr, nn = __import__("ctypes").c_void_p(), __import__("ctypes").c_int64()

    def dev_vec(which, n):
        device.check(device.lib().dcp_vector_device(model._h, which, __import__("ctypes").byref(ptr), __import__("ctypes").byref(nn)))
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        device.check(device.lib().dcp_vec_copy(ctx._h, n, ptr, __import__("ctypes").c_void_p(out.data_ptr())))
        return out
    rhs = dev_vec(device.VEC_NSE_RHS, n_u + n_p)
    trhs = dev_vec(device.VEC_TEMP_RHS, P.scalar("temp.n_dofs"))
    blocks = {(i, j): S.Wrap(model.nse_matrix.block(i, j)) for (i, j) in ((0, 0), (0, 1), (1, 0))}
    x, its, inner = S.solve_nse_block_preconditioned(B, S.Wrap(model.nse_matrix), blocks, S.Wrap(model.Mu_plus_A_preconditioner),
                                                     rhs, d_u, n_u, n_p, mp.time_step)
    xh = S.distribute(B, cs_lines(P, "nse.cs"), B.to_numpy(x))
    xh[n_u:] /= mp.time_step
    t, cg = S.solve_temperature(B, S.Wrap(model.temperature_matrix), S.Wrap(model.T_preconditioner), trhs, d_T)
    th = S.distribute(B, cs_lines(P, "temp.cs"), B.to_numpy(t))
    model.close()
    return dict(nse=xh, temp=th, fgmres=its, inner=inner, cg=cg)


# ---- classic Schur-complement path (data/aqua_planet_test_2d.prm: use_schur_complement_solver = true) -------------
def cpu_schur_step(P, mp, u0, T0):
    """assemble_nse_system + solve_NSE_Schur_complement (boussinesq_model.tpp:1867-1870, 1248-1414) on the CPU:
    oracle assembly, oracle ILU(0), numpy Krylov vectors."""
    from dycore_b200 import solvers as S
    from oracle import oracle as orc
    prm = orc.params_from(mp)
    B = S.NumpyBackend()
    n = P.scalar("nse.n_dofs")
    n_u, n_p = P.scalar("nse.n_u"), P.scalar("nse.n_p")
    vals, rhs = orc.assemble_nse_system(P, prm, u0, T0)
    rp, col, _, _ = P.csr("nse.full")
    A = sp.csr_matrix((vals, col, rp), shape=(n, n))
    A00 = A[:n_u, :n_u].tocsr()
    A00.sort_indices()
    rp0, col0 = A00.indptr.astype(np.int64), A00.indices.astype(np.int32)
    assert np.array_equal(rp0, P["nse.b00.rowptr"]) and np.array_equal(col0, P["nse.b00.col"])
    lu = orc.ilu0_factor(rp0, col0, np.ascontiguousarray(A00.data))

    def mat(M):
        return S.Wrap(lambda dst, src, M=M: dst.__setitem__(slice(None), M @ src))
    ilu = S.Wrap(lambda dst, src: dst.__setitem__(slice(None), orc.ilu0_solve(rp0, col0, lu, np.ascontiguousarray(src))))
    blocks = {(0, 0): mat(A00), (0, 1): mat(A[:n_u, n_u:]), (1, 0): mat(A[n_u:, :n_u])}
    lines = cs_lines(P, "nse.cs")
    x, its, inner = S.solve_nse_schur_complement(B, blocks, ilu, rhs, u0, n_u, n_p, mp.time_step,
                                                 lambda v: S.distribute(B, lines, v))
    return dict(nse=x, gmres=its, inner=inner)


def gpu_schur_step(ctx, P, mp, u0, T0):
    """The same with assembly, ILU(0), SpMVs and Krylov vectors on the device."""
    import ctypes
    import torch
    from dycore_b200 import device
    from dycore_b200 import solvers as S
    B = S.DeviceBackend(ctx)
    n_u, n_p = P.scalar("nse.n_u"), P.scalar("nse.n_p")
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    d_u, d_T = torch.from_numpy(u0).cuda(), torch.from_numpy(T0).cuda()
    torch.cuda.synchronize()
    model.assemble_nse_system(d_u, d_T)
    ptr, nn = ctypes.c_void_p(), ctypes.c_int64()
    device.check(device.lib().dcp_vector_device(model._h, device.VEC_NSE_RHS, ctypes.byref(ptr), ctypes.byref(nn)))
    rhs = torch.empty(n_u + n_p, dtype=torch.float64, device="cuda")
    device.check(device.lib().dcp_vec_copy(ctx._h, n_u + n_p, ptr, ctypes.c_void_p(rhs.data_ptr())))
    ilu = device.PreconditionILU(model, device.MAT_NSE, 0)
    blocks = {(i, j): S.Wrap(model.nse_matrix.block(i, j)) for (i, j) in ((0, 0), (0, 1), (1, 0))}
    x, its, inner = S.solve_nse_schur_complement(B, blocks, S.Wrap(ilu), rhs, d_u, n_u, n_p, mp.time_step,
                                                 model.distribute_nse_constraints)
    out = dict(nse=B.to_numpy(x), gmres=its, inner=inner)
    ilu.close()
    model.close()
    return out


# ---- FEEC block-preconditioned path (data/aqua_planet_shell_test_3d-feec.prm) --------------------------------------
def feec_mean_weights(P, n_gauss):
    """|K| / sum |K| in pressure-dof order: the weights of VectorTools::compute_mean_value for the DG0 pressure."""
    from dycore_b200 import solvers as S
    nw, nu, n = P.scalar("nse.n_w"), P.scalar("nse.n_u"), P.scalar("nse.n_dofs")
    vol = S.feec_cell_volumes(P["cell_vertices"], n_gauss)
    l2g = P["nse.l2g"].reshape(P.scalar("n_cells"), -1)
    pdof = l2g[:, -1].astype(np.int64) - (nw + nu)          # FE_DGQ(0): the last cell dof
    assert pdof.min() >= 0 and pdof.max() < n - nw - nu
    w = np.zeros(n - nw - nu)
    np.add.at(w, pdof, vol)
    return w / vol.sum()


def cpu_feec_step(P, mp, u0, T0):
    """assemble_nse_system + solve_NSE_block_preconditioned of ExteriorCalculus::BoussinesqModel
    (boussineq_model_FEEC.tpp:2255-2299, 1268-1477) on the CPU: oracle assembly + numpy Krylov vectors."""
    from dycore_b200 import solvers as S
    from oracle import oracle as orc
    prm = orc.params_from(mp)
    B = S.NumpyBackend()
    n = P.scalar("nse.n_dofs")
    nw, nu = P.scalar("nse.n_w"), P.scalar("nse.n_u")
    sizes = (nw, nu, n - nw - nu)
    vals, rhs = orc.feec_assemble_nse_system(P, prm, u0, T0)
    rp, col, _, _ = P.csr("nse.full")
    A = sp.csr_matrix((vals, col, rp), shape=(n, n))
    off = [0, nw, nw + nu, n]

    class Mat:
        def __init__(self, M):
            self.M = M

        def vmult(self, dst, src, B=None):
            dst[...] = self.M @ src

        def vmult_add(self, dst, src, B=None):
            dst += self.M @ src

    def jac(M):
        d = M.diagonal()
        return S.Wrap(lambda dst, src, d=d: dst.__setitem__(slice(None), src / d))
    blocks = {}
    for i in range(3):
        for j in range(3):
            Mij = A[off[i]:off[i + 1], off[j]:off[j + 1]].tocsr()
            if P.scalar(f"nse.b{i}{j}.nnz") > 0:
                blocks[(i, j)] = Mat(Mij)
    x, its, inner = S.solve_nse_block_preconditioned_feec(
        B, Mat(A), blocks, jac(A[:nw, :nw]), jac(A[nw:nw + nu, nw:nw + nu]), rhs, u0, sizes, mp.time_step,
        S.MeanValue(feec_mean_weights(P, 1), B), S.MeanValue(feec_mean_weights(P, 2), B))
    x = S.distribute(B, cs_lines(P, "nse.cs"), x)
    x[nw + nu:] /= mp.time_step
    return dict(nse=x, gmres=its, inner=inner)


def gpu_feec_step(ctx, P, mp, u0, T0):
    """The same with assembly, SpMVs, Jacobi sweeps and Krylov vectors on the device."""
    import ctypes
    import torch
    from dycore_b200 import device
    from dycore_b200 import solvers as S
    B = S.DeviceBackend(ctx)
    n = P.scalar("nse.n_dofs")
    nw, nu = P.scalar("nse.n_w"), P.scalar("nse.n_u")
    sizes = (nw, nu, n - nw - nu)
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    d_u, d_T = torch.from_numpy(u0).cuda(), torch.from_numpy(T0).cuda()
    torch.cuda.synchronize()
    model.assemble_nse_system(d_u, d_T)
    ptr, nn = ctypes.c_void_p(), ctypes.c_int64()
    device.check(device.lib().dcp_vector_device(model._h, device.VEC_NSE_RHS, ctypes.byref(ptr), ctypes.byref(nn)))
    rhs = torch.empty(n, dtype=torch.float64, device="cuda")
    device.check(device.lib().dcp_vec_copy(ctx._h, n, ptr, ctypes.c_void_p(rhs.data_ptr())))
    blocks = {(i, j): S.Wrap(model.nse_matrix.block(i, j)) for i in range(3) for j in range(3)
              if P.scalar(f"nse.b{i}{j}.nnz") > 0}
    x, its, inner = S.solve_nse_block_preconditioned_feec(
        B, S.Wrap(model.nse_matrix), blocks, S.Wrap(device.PreconditionJacobi(model, device.MAT_NSE, 0)),
        S.Wrap(device.PreconditionJacobi(model, device.MAT_NSE, 1)), rhs, d_u, sizes, mp.time_step,
        S.MeanValue(feec_mean_weights(P, 1), B), S.MeanValue(feec_mean_weights(P, 2), B))
    xh = S.distribute(B, cs_lines(P, "nse.cs"), B.to_numpy(x))
    xh[nw + nu:] /= mp.time_step
    model.close()
    return dict(nse=xh, gmres=its, inner=inner)
