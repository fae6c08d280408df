"""Pins for the CPU oracle, part 2: an independent numpy / scipy re-derivation.

Different code path on purpose: the local matrices come from vectorised einsum contractions of the block
formulas (not the dense i,j,q loop over FEValues views), the global matrix is assembled UNCONSTRAINED as a COO
sum, and the boundary conditions are applied algebraically afterwards,
    A_c = C^T A C  (+ sum_K |L_K[ii]| on constrained diagonals),   b_c = C^T (f - A k),
with C the constraint matrix (unit rows for free dofs, master weights for constrained ones) and k the vector of
inhomogeneities -- whereas the oracle resolves constraints entry by entry during the scatter like
AffineConstraints::distribute_local_to_global.  Agreement to 1e-12 pins both.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc


def _constraint_matrix(P, prefix, n):
    ld, lp = P[prefix + ".line_dof"], P[prefix + ".line_ptr"]
    ed, ew, ih = P[prefix + ".entry_dof"], P[prefix + ".entry_w"], P[prefix + ".inhom"]
    constrained = np.zeros(n, dtype=bool)
    constrained[ld] = True
    free = np.flatnonzero(~constrained)
    rows = [free]
    cols = [free]
    vals = [np.ones(len(free))]
    for l, g in enumerate(ld):
        sl = slice(lp[l], lp[l + 1])
        rows.append(np.full(lp[l + 1] - lp[l], g))
        cols.append(ed[sl])
        vals.append(ew[sl])
    C = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    k = np.zeros(n)
    k[ld] = ih
    return C, k, constrained


def _local_th(P, mp):
    dim, nc, nq = P.dim, P.n_cells, P.scalar("q_nse.nq")
    geo = P["geom.qn"].reshape(nc, 1 + dim * dim + dim, nq)
    w = geo[:, 0, :]
    Kinv = geo[:, 1:1 + dim * dim, :].reshape(nc, dim, dim, nq)          # [c, e, d, q]
    xq = geo[:, 1 + dim * dim:, :]                                        # [c, d, q]
    phi = P["tab.u_qn.phi"].reshape(nq, -1)
    dphi = P["tab.u_qn.dphi"].reshape(nq, -1, dim)
    psi = P["tab.p_qn.phi"].reshape(nq, -1)
    G = np.einsum("cedq,qae->cqad", Kinv, dphi)
    m = np.einsum("cq,qa,qb->cab", w, phi, phi)
    g = np.einsum("cq,cqad,cqbe->cabde", w, G, G)                         # g[a,b,d,e] = int d_d phi_a d_e phi_b
    bp = -np.einsum("cq,cqad,qb->cadb", w, G, psi)
    return w, G, xq, phi, psi, m, g, bp


def _sys_maps(P):
    field, base = P["nse.local_field"], P["nse.local_base"]
    dim = P.dim
    nu = int(base[field == 0].max()) + 1
    sys_u = np.zeros((dim, nu), dtype=int)
    for i, (f, b) in enumerate(zip(field, base)):
        if f < dim:
            sys_u[f, b] = i
    sys_p = np.array([i for i, f in enumerate(field) if f == dim])
    return sys_u, sys_p


@pytest.mark.parametrize("spec", [dict(geometry="shell", refine=1), dict(geometry="cube", refine=1)],
                         ids=["shell", "cube"])
def test_numpy_rederivation_matches_oracle(problem_factory, spec):
    from dycore_b200 import params
    from util import synthetic_fields
    P = problem_factory(**spec)
    mp = params.NAMED["cube_3d" if spec["geometry"] == "cube" else "shell_3d_classic"]
    prm = orc.params_from(mp)
    dim, nc = P.dim, P.n_cells
    n, nT, nd = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs"), P.scalar("nse.n_local")
    l2g = P["nse.l2g"].reshape(nc, nd)
    u, T = synthetic_fields(P)
    w, G, xq, phi, psi, m, g, bp = _local_th(P, mp)
    sys_u, sys_p = _sys_maps(P)
    nu_ = mp.time_step * mp.inv_re
    L = np.zeros((nc, nd, nd))
    Lp = np.zeros((nc, nd, nd))
    tr = np.einsum("cabdd->cab", g)
    for c in range(dim):
        for d in range(dim):
            blk = nu_ * g[:, :, :, d, c] + (m + nu_ * tr if c == d else 0.0)
            L[:, sys_u[c][:, None], sys_u[d][None, :]] = blk
        Lp[:, sys_u[c][:, None], sys_u[c][None, :]] = m + nu_ * tr
        L[:, sys_u[c][:, None], sys_p[None, :]] = bp[:, :, c, :]
        L[:, sys_p[:, None], sys_u[c][None, :]] = np.transpose(bp[:, :, c, :], (0, 2, 1))
    Lp[:, sys_p[:, None], sys_p[None, :]] = np.einsum("cq,qa,qb->cab", w, psi, psi)

    # local rhs
    Uc = u[l2g]                                                   # [c, nd]
    uq = np.stack([Uc[:, sys_u[c]] @ phi.T for c in range(dim)], axis=1)            # [c, comp, q]
    guq = np.stack([np.einsum("ca,cqad->cqd", Uc[:, sys_u[c]], G) for c in range(dim)], axis=1)  # [c,comp,q,d]
    phit = P["tab.t_qn.phi"].reshape(w.shape[1], -1)
    Tq = T[P["temp.l2g"].reshape(nc, -1)] @ phit.T
    rho = 1.0 - mp.expansion_coefficient * (Tq - mp.ref_temperature)
    if mp.cuboid_geometry:
        grav = np.zeros_like(xq)
        grav[:, dim - 1, :] = -mp.gravity_constant
        om = np.array([0.0, 0.0, mp.cor_scale * mp.omega])
    else:
        r = np.sqrt((xq ** 2).sum(axis=1, keepdims=True))
        grav = -mp.gravity_constant * xq / np.where(r > 1, r, np.sqrt(r))
        om = np.zeros(3)
    adv = np.einsum("cdq,ckqd->ckq", uq, guq)
    cor = 2.0 * np.stack([om[1] * uq[:, 2] - om[2] * uq[:, 1], om[2] * uq[:, 0] - om[0] * uq[:, 2],
                          om[0] * uq[:, 1] - om[1] * uq[:, 0]], axis=1)
    F = uq + mp.time_step * rho[:, None, :] * mp.g_scale * grav - mp.time_step * adv - mp.time_step * cor
    l = np.zeros((nc, nd))
    for c in range(dim):
        l[:, sys_u[c]] = np.einsum("cq,cq,qa->ca", w, F[:, c, :], phi)

    # unconstrained global objects, then algebraic constraints
    ii = np.repeat(l2g[:, :, None], nd, axis=2).ravel()
    jj = np.repeat(l2g[:, None, :], nd, axis=1).ravel()
    A = sp.csr_matrix((L.ravel(), (ii, jj)), shape=(n, n))
    Ap = sp.csr_matrix((Lp.ravel(), (ii, jj)), shape=(n, n))
    f = np.bincount(l2g.ravel(), weights=l.ravel(), minlength=n)
    C, k, constrained = _constraint_matrix(P, "nse.cs", n)
    def constrained_diag(Lk, idx, nn):
        # deal.II keeps |L_ii| on a constrained diagonal, or the cell's mean |diag| when L_ii is exactly zero
        d = np.abs(np.einsum("cii->ci", Lk))
        d = np.where(d != 0.0, d, d.mean(axis=1, keepdims=True))
        return np.bincount(idx.ravel(), weights=d.ravel(), minlength=nn)
    dabs = constrained_diag(L, l2g, n)
    dabs_p = constrained_diag(Lp, l2g, n)
    A_c = (C.T @ A @ C + sp.diags(np.where(constrained, dabs, 0.0))).tocsr()
    Ap_c = (C.T @ Ap @ C + sp.diags(np.where(constrained, dabs_p, 0.0))).tocsr()
    b_c = C.T @ (f - A @ k)

    vals, rhs = orc.assemble_nse_system(P, prm, u, T)
    rp, col, _, _ = P.csr("nse.full")
    Ao = sp.csr_matrix((vals, col, rp), shape=(n, n))
    assert abs(Ao - A_c).max() <= 1e-12 * abs(Ao).max()
    assert np.abs(rhs - b_c).max() <= 1e-12 * np.abs(rhs).max()
    # every structurally possible entry of the independent matrix lies inside the harness pattern
    pat = sp.csr_matrix((np.ones(len(col)), col, rp), shape=(n, n))
    Ac_nz = A_c.copy()
    Ac_nz.data = (np.abs(Ac_nz.data) > 1e-300).astype(float)
    assert (Ac_nz - Ac_nz.multiply(pat)).nnz == 0
    pv = orc.assemble_nse_preconditioner(P, prm)
    rp, col, _, _ = P.csr("pre.full")
    Po = sp.csr_matrix((pv, col, rp), shape=(n, n))
    assert abs(Po - Ap_c).max() <= 1e-12 * abs(Po).max()

    # temperature: M, K, rhs with the matrix_for_bc elimination
    nqt = P.scalar("q_temp.nq")
    geo = P["geom.qt"].reshape(nc, 1 + dim * dim + dim, nqt)
    wt = geo[:, 0, :]
    Kt = geo[:, 1:1 + dim * dim, :].reshape(nc, dim, dim, nqt)
    ph = P["tab.t_qt.phi"].reshape(nqt, -1)
    dph = P["tab.t_qt.dphi"].reshape(nqt, -1, dim)
    ndt = ph.shape[1]
    Gt = np.einsum("cedq,qae->cqad", Kt, dph)
    LM = np.einsum("cq,qa,qb->cab", wt, ph, ph)
    LK = mp.inv_pe * np.einsum("cq,cqad,cqbd->cab", wt, Gt, Gt)
    tl2g = P["temp.l2g"].reshape(nc, ndt)
    ti = np.repeat(tl2g[:, :, None], ndt, axis=2).ravel()
    tj = np.repeat(tl2g[:, None, :], ndt, axis=1).ravel()
    M = sp.csr_matrix((LM.ravel(), (ti, tj)), shape=(nT, nT))
    K = sp.csr_matrix((LK.ravel(), (ti, tj)), shape=(nT, nT))
    Ct, kt, ct = _constraint_matrix(P, "temp.cs", nT)
    dM = constrained_diag(LM, tl2g, nT)
    dK = constrained_diag(LK, tl2g, nT)
    M_c = Ct.T @ M @ Ct + sp.diags(np.where(ct, dM, 0.0))
    K_c = Ct.T @ K @ Ct + sp.diags(np.where(ct, dK, 0.0))
    om_, ok_ = orc.assemble_temperature_matrix(P, prm)
    rp, col, _, _ = P.csr("temp.pat")
    assert abs(sp.csr_matrix((om_, col, rp), shape=(nT, nT)) - M_c).max() <= 1e-12 * np.abs(om_).max()
    assert abs(sp.csr_matrix((ok_, col, rp), shape=(nT, nT)) - K_c).max() <= 1e-12 * np.abs(ok_).max()
    tau = mp.time_step / mp.NSE_solver_interval
    phu = P["tab.u_qt.phi"].reshape(nqt, -1)
    uqt = np.stack([u[l2g][:, sys_u[c]] @ phu.T for c in range(dim)], axis=1)
    Tc = T[tl2g]
    Tq = Tc @ ph.T
    gT = np.einsum("ca,cqad->cqd", Tc, Gt)
    ugT = np.einsum("cdq,cqd->cq", uqt, gT)
    lt = np.einsum("cq,cq,qa->ca", wt, Tq - tau * ugT, ph)
    ft = np.bincount(tl2g.ravel(), weights=lt.ravel(), minlength=nT)
    bt = Ct.T @ (ft - (M + tau * K) @ kt)
    rt = orc.assemble_temperature_rhs(P, prm, T, u)
    assert np.abs(rt - bt).max() <= 1e-12 * np.abs(rt).max()
