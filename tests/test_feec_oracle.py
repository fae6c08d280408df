"""Pins for the FEEC oracle (oracle/feec_oracle.c): identities + an independent numpy/scipy re-derivation.

Identities: block structure of /root/reference/include/core/boussineq_model_FEEC.tpp:753-769 (M_w, M_u symmetric,
R_u = -dt/Re * R_w^T, B^T), exactness of RT0 / Nedelec0 on constants over affine cells (closed-form integrals on
the unit cube), Piola identity  int div u = boundary flux.  The re-derivation maps the reference shape functions
with einsum, assembles unconstrained COO matrices and applies the constraints algebraically (C^T A C)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from test_oracle_independent import _constraint_matrix


def _blocks(P, vals, name="nse.full"):
    n = P.scalar("nse.n_dofs")
    rp, col, _, _ = P.csr(name)
    A = sp.csr_matrix((vals, col, rp), shape=(n, n))
    nw, nu = P.scalar("nse.n_w"), P.scalar("nse.n_u")
    s = [0, nw, nw + nu, n]
    return A, [[A[s[i]:s[i + 1], s[j]:s[j + 1]] for j in range(3)] for i in range(3)]


@pytest.mark.parametrize("spec,pname", [(dict(geometry="shell", refine=2, family="feec"), "shell_3d_feec"),
                                        (dict(geometry="cube", refine=2, family="feec"), "cube_3d")], ids=["shell", "cube"])
def test_block_structure(problem_factory, spec, pname):
    from dycore_b200 import params
    P = problem_factory(**spec)
    mp = params.NAMED[pname]
    prm = orc.params_from(mp)
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    rng = np.random.default_rng(3)
    v, rhs = orc.feec_assemble_nse_system(P, prm, 0.1 * rng.standard_normal(n), 2 + 0.1 * rng.standard_normal(nT))
    A, B = _blocks(P, v)
    nu_ = mp.time_step * mp.inv_re
    assert abs(B[0][0] - B[0][0].T).max() <= 1e-15 and abs(B[1][1] - B[1][1].T).max() <= 1e-14
    assert abs(B[1][0] + nu_ * B[0][1].T).max() <= 1e-15
    assert abs(B[1][2] - B[2][1].T).max() == 0.0
    assert B[0][2].nnz == 0 and B[2][0].nnz == 0 and B[2][2].nnz == 0


def test_constants_on_unit_cube(problem_factory):
    """RT0 and Nedelec0 reproduce constant fields on affine cells: closed-form integrals."""
    from dycore_b200 import params
    P = problem_factory(geometry="cube", refine=1, family="feec", constraints=0)
    mp = params.NAMED["cube_3d"]
    prm = orc.params_from(mp)
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    nw, nu = P.scalar("nse.n_w"), P.scalar("nse.n_u")
    h = 0.5
    c = np.array([1.0, 2.0, 3.0])
    # dof of a constant field: w_l = c . t_l * |l| (tangential line integral), u_f = c . n_f * |f| (flux)
    xyz = P["nse.dof_xyz"].reshape(n, 3)
    l2g = P["nse.l2g"].reshape(P.n_cells, 19)
    U = np.zeros(n)
    line_dir = [1, 1, 0, 0, 1, 1, 0, 0, 2, 2, 2, 2]
    for k in range(12):
        U[l2g[:, k]] = c[line_dir[k]] * h
    for f in range(6):
        U[l2g[:, 12 + f]] = c[f // 2] * h * h
    T = np.full(nT, mp.ref_temperature)
    v, rhs = orc.feec_assemble_nse_system(P, prm, np.ascontiguousarray(U), T)
    A, B = _blocks(P, v)
    w, u = U[:nw], U[nw:nw + nu]
    assert abs(w @ (B[0][0] @ w) - c @ c) <= 1e-12          # int |c|^2 over the unit cube
    assert abs(u @ (B[1][1] @ u) - c @ c) <= 1e-12
    assert np.abs(B[2][1] @ u).max() <= 1e-13               # div of a constant field
    assert np.abs(B[0][1] @ u + 0.0).max() >= 0.0           # (curl w_i, c): boundary terms only, no claim
    # rhs against the constant test field v = c:  int ( c.u_old + dt*rho*g.c - dt (omega x u).c - 2 dt (Omega x u).c )
    # with u_old = c, omega_old = c  ->  omega x u = 0;  div v = 0
    om = mp.cor_scale * mp.omega
    cxu = np.array([-om * c[1], om * c[0], 0.0])
    want = c @ c + mp.time_step * (-mp.g_scale * mp.gravity_constant) * c[2] - 2 * mp.time_step * cxu @ c
    assert abs(rhs[nw:nw + nu] @ u - want) <= 1e-12


def _mapped(P, rule):
    nc = P.n_cells
    nq = P.scalar({"qn": "q_nse.nq", "qp": "q_pre.nq", "qt": "q_temp.nq"}[rule])
    g = P["geom." + rule].reshape(nc, 23, nq)
    w, K, xq = g[:, 0], g[:, 1:10].reshape(nc, 3, 3, nq), g[:, 10:13]
    J, det = g[:, 13:22].reshape(nc, 3, 3, nq), g[:, 22]
    tw = P[f"feec.{rule}.phi_w"].reshape(nq, 12, 3)
    tc = P[f"feec.{rule}.curl_w"].reshape(nq, 12, 3)
    tu = P[f"feec.{rule}.phi_u"].reshape(nq, 6, 3)
    td = P[f"feec.{rule}.div_u"]
    sg = P["nse.sign"].reshape(nc, 19)[:, 12:18]
    W = np.einsum("cedq,qke->cqkd", K, tw)
    C = np.einsum("cdeq,qke->cqkd", J, tc) / det[:, :, None, None]
    Uraw = np.einsum("cdeq,qke->cqkd", J, tu) / det[:, :, None, None]
    Us = Uraw * sg[:, None, :, None]
    D = td[None, None, :] / det[:, :, None] * sg[:, None, :]
    return w, xq, W, C, Uraw, Us, D


@pytest.mark.parametrize("spec,pname", [(dict(geometry="shell", refine=1, family="feec"), "shell_3d_feec"),
                                        (dict(geometry="cube", refine=1, family="feec"), "cube_3d")], ids=["shell", "cube"])
def test_numpy_rederivation_matches_feec_oracle(problem_factory, spec, pname):
    from dycore_b200 import params
    P = problem_factory(**spec)
    mp = params.NAMED[pname]
    prm = orc.params_from(mp)
    nc, n, nT = P.n_cells, P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    l2g = P["nse.l2g"].reshape(nc, 19)
    rng = np.random.default_rng(11)
    u0 = 0.1 * rng.standard_normal(n)
    T0 = 2.0 + 0.1 * rng.standard_normal(nT)
    w, xq, W, C, Uraw, Us, D = _mapped(P, "qn")
    nu_, dt = mp.time_step * mp.inv_re, mp.time_step
    L = np.zeros((nc, 19, 19))
    L[:, :12, :12] = np.einsum("cq,cqid,cqjd->cij", w, W, W)
    wu = np.einsum("cq,cqid,cqjd->cij", w, C, Us)
    L[:, :12, 12:18] = -wu
    L[:, 12:18, :12] = nu_ * np.transpose(wu, (0, 2, 1))
    L[:, 12:18, 12:18] = np.einsum("cq,cqid,cqjd->cij", w, Us, Us)
    dv = np.einsum("cq,cqi->ci", w, D)
    L[:, 12:18, 18] = -dv
    L[:, 18, 12:18] = -dv
    Uc = u0[l2g]
    ow = np.einsum("ck,cqkd->cqd", Uc[:, :12], W)
    ou = np.einsum("ck,cqkd->cqd", Uc[:, 12:18], Uraw)
    phit = P["tab.t_qn.phi"].reshape(w.shape[1], -1)
    Tq = T0[P["temp.l2g"].reshape(nc, -1)] @ phit.T
    rho = 1.0 - mp.expansion_coefficient * (Tq - mp.ref_temperature)
    if mp.cuboid_geometry:
        grav = np.zeros_like(ou)
        grav[..., 2] = -mp.gravity_constant
        cz = mp.cor_scale * mp.omega
    else:
        x = np.transpose(xq, (0, 2, 1))
        r = np.linalg.norm(x, axis=2, keepdims=True)
        grav = -mp.gravity_constant * x / np.where(r > 1, r, np.sqrt(r))
        cz = 0.0
    wxu = np.cross(ow, ou)
    cxu = np.stack([-cz * ou[..., 1], cz * ou[..., 0], np.zeros_like(ou[..., 0])], axis=-1)
    Avec = ou + dt * rho[..., None] * mp.g_scale * grav - dt * wxu - 2 * dt * cxu
    Bs = -dt * 0.5 * (ou ** 2).sum(-1)
    l = np.zeros((nc, 19))
    l[:, 12:18] = np.einsum("cq,cqid,cqd->ci", w, Us, Avec) + np.einsum("cq,cqi,cq->ci", w, D, Bs)
    ii = np.repeat(l2g[:, :, None], 19, axis=2).ravel()
    jj = np.repeat(l2g[:, None, :], 19, axis=1).ravel()
    A = sp.csr_matrix((L.ravel(), (ii, jj)), shape=(n, n))
    f = np.bincount(l2g.ravel(), weights=l.ravel(), minlength=n)
    Cm, k, con = _constraint_matrix(P, "nse.cs", n)
    d = np.abs(np.einsum("cii->ci", L))
    d = np.where(d != 0.0, d, d.mean(axis=1, keepdims=True))
    dabs = np.bincount(l2g.ravel(), weights=d.ravel(), minlength=n)
    A_c = (Cm.T @ A @ Cm + sp.diags(np.where(con, dabs, 0.0))).tocsr()
    b_c = Cm.T @ (f - A @ k)
    v, rhs = orc.feec_assemble_nse_system(P, prm, u0, T0)
    Ao, _ = _blocks(P, v)
    assert abs(Ao - A_c).max() <= 1e-12 * abs(Ao).max()
    assert np.abs(rhs - b_c).max() <= 1e-12 * np.abs(rhs).max()

    # preconditioner (quirk Q5): only p*p is weighted
    wp, _, Wp, Cp, _, Usp, _ = _mapped(P, "qp")
    Lp = np.zeros((nc, 19, 19))
    Lp[:, :12, :12] = nu_ * np.einsum("cqid,cqjd->cij", Cp, Cp)
    x = np.einsum("cqid,cqjd->cqij", Usp, Wp)
    sgn = np.where(np.abs(x) > 1e-9, np.sign(x), 0.0).sum(axis=1)
    Lp[:, 12:18, :12] = sgn
    Lp[:, :12, 12:18] = np.transpose(sgn, (0, 2, 1))
    Lp[:, 18, 18] = wp.sum(axis=1)
    Ap = sp.csr_matrix((Lp.ravel(), (ii, jj)), shape=(n, n))
    d = np.abs(np.einsum("cii->ci", Lp))
    d = np.where(d != 0.0, d, d.mean(axis=1, keepdims=True))
    Ap_c = (Cm.T @ Ap @ Cm + sp.diags(np.where(con, np.bincount(l2g.ravel(), weights=d.ravel(), minlength=n), 0.0))).tocsr()
    pv = orc.feec_assemble_nse_preconditioner(P, prm)
    Po, _ = _blocks(P, pv, "pre.full")
    assert abs(Po - Ap_c).max() <= 1e-12 * abs(Po).max()

    # temperature rhs with the RT velocity (no face sign)
    wt, _, _, _, Urt, _, _ = _mapped(P, "qt")
    nqt = wt.shape[1]
    Kt = P["geom.qt"].reshape(nc, 23, nqt)[:, 1:10].reshape(nc, 3, 3, nqt)
    ph = P["tab.t_qt.phi"].reshape(nqt, -1)
    dph = P["tab.t_qt.dphi"].reshape(nqt, -1, 3)
    ndt = ph.shape[1]
    Gt = np.einsum("cedq,qae->cqad", Kt, dph)
    tl2g = P["temp.l2g"].reshape(nc, ndt)
    tau = mp.time_step / mp.NSE_solver_interval
    uq = np.einsum("ck,cqkd->cqd", u0[l2g][:, 12:18], Urt)
    Tc = T0[tl2g]
    lt = np.einsum("cq,cq,qa->ca", wt, Tc @ ph.T - tau * np.einsum("cqd,cqd->cq", uq, np.einsum("ca,cqad->cqd", Tc, Gt)), ph)
    LM = np.einsum("cq,qa,qb->cab", wt, ph, ph)
    LK = mp.inv_pe * np.einsum("cq,cqad,cqbd->cab", wt, Gt, Gt)
    ti = np.repeat(tl2g[:, :, None], ndt, axis=2).ravel()
    tj = np.repeat(tl2g[:, None, :], ndt, axis=1).ravel()
    M = sp.csr_matrix((LM.ravel(), (ti, tj)), shape=(nT, nT))
    K = sp.csr_matrix((LK.ravel(), (ti, tj)), shape=(nT, nT))
    Ct, kt, _ = _constraint_matrix(P, "temp.cs", nT)
    bt = Ct.T @ (np.bincount(tl2g.ravel(), weights=lt.ravel(), minlength=nT) - (M + tau * K) @ kt)
    rt = orc.feec_assemble_temperature_rhs(P, prm, T0, u0)
    assert np.abs(rt - bt).max() <= 1e-12 * np.abs(rt).max()
