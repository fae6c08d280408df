"""ctypes binding of the CPU oracle (oracle/boussinesq_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product path.  Parity unpinned by the reference (see the header of
boussinesq_oracle.c); pinned by identities, an independent numpy derivation and golden fixtures.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_lp = ctypes.POINTER(ctypes.c_int64)


class Params(ctypes.Structure):
    """Mirror of `orc_params`; values derived as in SURVEY.md Appendix B."""
    _fields_ = [("dim", ctypes.c_int32), ("cuboid", ctypes.c_int32), ("nse_interval", ctypes.c_int32),
                ("pad", ctypes.c_int32), ("dt", ctypes.c_double), ("inv_re", ctypes.c_double),
                ("inv_pe", ctypes.c_double), ("beta", ctypes.c_double), ("T_ref", ctypes.c_double),
                ("g_scale", ctypes.c_double), ("g_const", ctypes.c_double), ("cor_scale", ctypes.c_double),
                ("omega", ctypes.c_double)]


class _Cs(ctypes.Structure):
    _fields_ = [("n_dofs", ctypes.c_int64), ("line_of_dof", c_ip), ("line_ptr", c_ip), ("entry_dof", c_ip),
                ("entry_w", c_dp), ("inhom", c_dp)]


class _Csr(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_int64), ("rowptr", c_lp), ("col", c_ip), ("val", c_dp)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        _LIB = ctypes.CDLL(path)
        _LIB.orc_missing_entries.restype = ctypes.c_int
        _LIB.orc_max_threads.restype = ctypes.c_int
    return _LIB


def _dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dp)


def _ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_ip)


def _lp(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_lp)


def _pad1(a, dtype):
    """ctypes needs a valid pointer even for empty arrays."""
    return a if a.size else np.zeros(1, dtype=dtype)


def make_cs(P, prefix):
    keep = [P[prefix + ".line_of_dof"], _pad1(P[prefix + ".line_ptr"], np.int32),
            _pad1(P[prefix + ".entry_dof"], np.int32), _pad1(P[prefix + ".entry_w"], np.float64),
            _pad1(P[prefix + ".inhom"], np.float64)]
    cs = _Cs(len(keep[0]), _ip(keep[0]), _ip(keep[1]), _ip(keep[2]), _dp(keep[3]), _dp(keep[4]))
    cs._keep = keep
    return cs


def make_csr(rowptr, col, val):
    A = _Csr(len(rowptr) - 1, _lp(rowptr), _ip(_pad1(col, np.int32)), _dp(_pad1(val, np.float64)))
    A._keep = (rowptr, col, val)
    return A


def _check_missing(what):
    n = lib().orc_missing_entries()
    lib().orc_reset_missing()
    if n:
        raise AssertionError(f"oracle {what}: {n} scatter targets are not in the sparsity pattern")


def assemble_nse_system(P, prm, old_nse, old_temp, use_omp=False):
    """-> (values of the block-concatenated nse_matrix on pattern `nse.full`, rhs[n_u+n_p])"""
    rowptr, col, n, _ = P.csr("nse.full")
    val = np.zeros(len(col))
    rhs = np.zeros(n)
    A = make_csr(rowptr, col, val)
    cs = make_cs(P, "nse.cs")
    tu, tp, tt = "tab.u_qn", "tab.p_qn", "tab.t_qn"
    lib().orc_assemble_nse_system(
        ctypes.byref(prm), ctypes.c_int64(P.n_cells), P.scalar("nse.n_local"), P.scalar("q_nse.nq"),
        P.scalar(tu + ".nd"), P.scalar(tp + ".nd"), P.scalar(tt + ".nd"), _ip(P["nse.local_field"]),
        _ip(P["nse.local_base"]), _dp(P[tu + ".phi"]), _dp(P[tu + ".dphi"]), _dp(P[tp + ".phi"]),
        _dp(P[tt + ".phi"]), _dp(P["geom.qn"]), _ip(P["nse.l2g"]), _ip(P["temp.l2g"]), _dp(old_nse), _dp(old_temp),
        ctypes.byref(cs), ctypes.byref(A), _dp(rhs), ctypes.c_int64(n), int(use_omp))
    _check_missing("nse_system")
    return val, rhs


def assemble_nse_preconditioner(P, prm, use_omp=False):
    rowptr, col, n, _ = P.csr("pre.full")
    val = np.zeros(len(col))
    A = make_csr(rowptr, col, val)
    cs = make_cs(P, "nse.cs")
    tu, tp = "tab.u_qn", "tab.p_qn"
    lib().orc_assemble_nse_preconditioner(
        ctypes.byref(prm), ctypes.c_int64(P.n_cells), P.scalar("nse.n_local"), P.scalar("q_nse.nq"),
        P.scalar(tu + ".nd"), P.scalar(tp + ".nd"), _ip(P["nse.local_field"]), _ip(P["nse.local_base"]),
        _dp(P[tu + ".phi"]), _dp(P[tu + ".dphi"]), _dp(P[tp + ".phi"]), _dp(P["geom.qn"]), _ip(P["nse.l2g"]),
        ctypes.byref(cs), ctypes.byref(A), int(use_omp))
    _check_missing("nse_preconditioner")
    return val


def assemble_temperature_matrix(P, prm, use_omp=False):
    rowptr, col, n, _ = P.csr("temp.pat")
    m = np.zeros(len(col))
    k = np.zeros(len(col))
    M, K = make_csr(rowptr, col, m), make_csr(rowptr, col, k)
    cs = make_cs(P, "temp.cs")
    tt = "tab.t_qt"
    lib().orc_assemble_temperature_matrix(
        ctypes.byref(prm), ctypes.c_int64(P.n_cells), P.scalar("temp.n_local"), P.scalar("q_temp.nq"),
        _dp(P[tt + ".phi"]), _dp(P[tt + ".dphi"]), _dp(P["geom.qt"]), _ip(P["temp.l2g"]), ctypes.byref(cs),
        ctypes.byref(M), ctypes.byref(K), int(use_omp))
    _check_missing("temperature_matrix")
    return m, k


def temperature_matrix_combine(mass, stiff, factor):
    out = np.empty_like(mass)
    lib().orc_temperature_matrix_combine(ctypes.c_int64(len(mass)), _dp(mass), _dp(stiff), ctypes.c_double(factor),
                                         _dp(out))
    return out


def assemble_temperature_rhs(P, prm, old_temp, nse_solution, use_omp=False):
    n = P.scalar("temp.n_dofs")
    rhs = np.zeros(n)
    cs = make_cs(P, "temp.cs")
    tt, tu = "tab.t_qt", "tab.u_qt"
    lib().orc_assemble_temperature_rhs(
        ctypes.byref(prm), ctypes.c_int64(P.n_cells), P.scalar("temp.n_local"), P.scalar("q_temp.nq"),
        P.scalar("nse.n_local"), P.scalar(tu + ".nd"), _dp(P[tt + ".phi"]), _dp(P[tt + ".dphi"]),
        _ip(P["nse.local_field"]), _ip(P["nse.local_base"]), _dp(P[tu + ".phi"]), _dp(P["geom.qt"]),
        _ip(P["temp.l2g"]), _ip(P["nse.l2g"]), _dp(old_temp), _dp(nse_solution), ctypes.byref(cs), _dp(rhs),
        ctypes.c_int64(n), int(use_omp))
    return rhs


def spmv(rowptr, col, val, x, y=None, add=False, use_omp=False):
    n = len(rowptr) - 1
    if y is None:
        y = np.zeros(n)
    lib().orc_spmv(ctypes.c_int64(n), _lp(rowptr), _ip(_pad1(col, np.int32)), _dp(_pad1(val, np.float64)), _dp(x),
                   _dp(y), int(add), int(use_omp))
    return y


def max_threads():
    return lib().orc_max_threads()


def params_from(mp):
    """orc_params from a dycore_b200.params.ModelParameters (derived numbers as in SURVEY.md Appendix B)."""
    return Params(dim=mp.space_dimension, cuboid=int(mp.cuboid_geometry), nse_interval=mp.NSE_solver_interval, pad=0,
                  dt=mp.time_step, inv_re=mp.inv_re, inv_pe=mp.inv_pe, beta=mp.expansion_coefficient,
                  T_ref=mp.ref_temperature, g_scale=mp.g_scale, g_const=mp.gravity_constant, cor_scale=mp.cor_scale,
                  omega=mp.omega)


# ---- FEEC family (oracle/feec_oracle.c) ------------------------------------------------------------------
def _feec_tabs(P, rule):
    return [_dp(P[f"feec.{rule}.{n}"]) for n in ("phi_w", "curl_w", "phi_u", "div_u")]


def feec_assemble_nse_system(P, prm, old_nse, old_temp, use_omp=False):
    rowptr, col, n, _ = P.csr("nse.full")
    val = np.zeros(len(col))
    rhs = np.zeros(n)
    A = make_csr(rowptr, col, val)
    cs = make_cs(P, "nse.cs")
    tw, tc, tu, td = _feec_tabs(P, "qn")
    lib().orc_feec_assemble_nse_system(
        ctypes.byref(prm), ctypes.c_int64(P.n_cells), P.scalar("q_nse.nq"), P.scalar("tab.t_qn.nd"), tw, tc, tu, td,
        _dp(P["tab.t_qn.phi"]), _dp(P["geom.qn"]), _dp(P["nse.sign"]), _ip(P["nse.l2g"]), _ip(P["temp.l2g"]),
        _dp(old_nse), _dp(old_temp), ctypes.byref(cs), ctypes.byref(A), _dp(rhs), ctypes.c_int64(n), int(use_omp))
    _check_missing("feec nse_system")
    return val, rhs


def feec_assemble_nse_preconditioner(P, prm, use_omp=False):
    rowptr, col, n, _ = P.csr("pre.full")
    val = np.zeros(len(col))
    A = make_csr(rowptr, col, val)
    cs = make_cs(P, "nse.cs")
    tw, tc, tu, td = _feec_tabs(P, "qp")
    lib().orc_feec_assemble_nse_preconditioner(
        ctypes.byref(prm), ctypes.c_int64(P.n_cells), P.scalar("q_pre.nq"), tw, tc, tu, td, _dp(P["geom.qp"]),
        _dp(P["nse.sign"]), _ip(P["nse.l2g"]), ctypes.byref(cs), ctypes.byref(A), int(use_omp))
    _check_missing("feec nse_preconditioner")
    return val


def feec_compact_geometry(P, name):
    """First 13 rows (JxW, Kinv, xq) of the extended FEEC record, for the classic temperature-matrix oracle."""
    nq = P.scalar("q_temp.nq" if name == "geom.qt" else "q_nse.nq")
    return np.ascontiguousarray(P[name].reshape(P.n_cells, 23, nq)[:, :13, :])


def feec_assemble_temperature_matrix(P, prm, use_omp=False):
    rowptr, col, n, _ = P.csr("temp.pat")
    m = np.zeros(len(col))
    k = np.zeros(len(col))
    M, K = make_csr(rowptr, col, m), make_csr(rowptr, col, k)
    cs = make_cs(P, "temp.cs")
    g13 = feec_compact_geometry(P, "geom.qt")
    lib().orc_assemble_temperature_matrix(
        ctypes.byref(prm), ctypes.c_int64(P.n_cells), P.scalar("temp.n_local"), P.scalar("q_temp.nq"),
        _dp(P["tab.t_qt.phi"]), _dp(P["tab.t_qt.dphi"]), _dp(g13), _ip(P["temp.l2g"]), ctypes.byref(cs),
        ctypes.byref(M), ctypes.byref(K), int(use_omp))
    _check_missing("feec temperature_matrix")
    return m, k


def feec_assemble_temperature_rhs(P, prm, old_temp, nse_solution, use_omp=False):
    n = P.scalar("temp.n_dofs")
    rhs = np.zeros(n)
    cs = make_cs(P, "temp.cs")
    lib().orc_feec_assemble_temperature_rhs(
        ctypes.byref(prm), ctypes.c_int64(P.n_cells), P.scalar("temp.n_local"), P.scalar("q_temp.nq"),
        _dp(P["tab.t_qt.phi"]), _dp(P["tab.t_qt.dphi"]), _dp(P["feec.qt.phi_u"]), _dp(P["geom.qt"]),
        _ip(P["temp.l2g"]), _ip(P["nse.l2g"]), _dp(old_temp), _dp(nse_solution), ctypes.byref(cs), _dp(rhs),
        ctypes.c_int64(n), int(use_omp))
    return rhs


# ---- passes next to the solves, numpy restatements -----------------------------------------
def cell_diameters(P):
    """cell->diameter(): the longest diagonal between opposite vertices (deal.II TriaAccessor::diameter)."""
    X = P["cell_vertices"].reshape(P.n_cells, 1 << P.dim, P.dim)
    nv = 1 << P.dim
    d = [np.linalg.norm(X[:, nv - 1 - v] - X[:, v], axis=1) for v in range(nv // 2)]
    return np.max(np.stack(d, axis=1), axis=1)


def velocity_extrema(P, nse_solution):
    """(get_maximal_velocity, get_cfl_number) over the locally owned cells.

    Classic: boussinesq_model.tpp:1023-1061, 1064-1098 -- the QIterated(QTrapez, degree) points are the Lagrange
    nodes of the velocity element, the values there are the nodal values.  FEEC: boussineq_model_FEEC.tpp:1158-1240
    -- the 8 vertices, default Q1 mapping, Raviart-Thomas component mapped by J u_hat / det J, no face signs."""
    dim, nc = P.dim, P.scalar("n_owned_cells") or P.n_cells
    l2g = P["nse.l2g"].reshape(P.n_cells, -1)[:nc]
    feec = "feec" in P.spec.get("family", "classic")
    if not feec:
        field, base = P["nse.local_field"], P["nse.local_base"]
        ndu = 3 ** dim
        U = np.zeros((nc, ndu, dim))
        for k in range(l2g.shape[1]):
            if field[k] < dim:
                U[:, base[k], field[k]] = nse_solution[l2g[:, k]]
        speed = np.linalg.norm(U, axis=2).max(axis=1)
    else:
        X = P["cell_vertices"].reshape(P.n_cells, 8, 3)[:nc]
        Uf = nse_solution[l2g[:, 12:18]]
        speed = np.zeros(nc)
        for v in range(8):
            b = [(v >> k) & 1 for k in range(3)]
            J = np.zeros((nc, 3, 3))
            for s in range(8):
                t = [(s >> k) & 1 for k in range(3)]
                for j in range(3):
                    g = 1.0 if t[j] else -1.0
                    for k in range(3):
                        if k != j and t[k] != b[k]:
                            g = 0.0
                    if g != 0.0:
                        J[:, :, j] += g * X[:, s, :]
            uh = np.stack([Uf[:, 0 + b[0]], Uf[:, 2 + b[1]], Uf[:, 4 + b[2]]], axis=1)
            u = np.einsum("cij,cj->ci", J, uh) / np.linalg.det(J)[:, None]
            speed = np.maximum(speed, np.linalg.norm(u, axis=1))
    cfl = (np.maximum(speed, 1e-10) / cell_diameters(P)[:nc]).max()
    return float(speed.max()), float(cfl)


def constraints_distribute(P, prefix, x):
    """AffineConstraints::distribute: x[line] = sum_k w_k x[master_k] + inhomogeneity (boussinesq_model.tpp:1233, 1442)."""
    x = x.copy()
    line_dof, line_ptr = P[prefix + ".line_dof"], P[prefix + ".line_ptr"]
    entry_dof, entry_w, inhom = P[prefix + ".entry_dof"], P[prefix + ".entry_w"], P[prefix + ".inhom"]
    src = x.copy()
    for l, g in enumerate(line_dof):
        sl = slice(line_ptr[l], line_ptr[l + 1])
        x[g] = float(np.dot(entry_w[sl], src[entry_dof[sl]])) + inhom[l]
    return x


# ---- ILU(0) (oracle/ilu_oracle.c) -----------------------------------------------------------------------------------
def ilu0_factor(rowptr, col, val, n=None):
    """ILU(0) factors on the pattern of A (strict lower part L with unit diagonal, rest U)."""
    n = len(rowptr) - 1 if n is None else n
    lu = np.zeros(len(col))
    L = lib()
    L.orc_ilu0_factor.restype = ctypes.c_int
    rc = L.orc_ilu0_factor(ctypes.c_int64(n), _lp(rowptr), _ip(col), _dp(val), _dp(lu))
    if rc:
        raise RuntimeError(f"ilu0: row {rc - 1} has no diagonal entry")
    return lu


def ilu0_solve(rowptr, col, lu, x, n=None):
    n = len(rowptr) - 1 if n is None else n
    y = np.zeros(n)
    lib().orc_ilu0_solve(ctypes.c_int64(n), _lp(rowptr), _ip(col), _dp(lu), _dp(np.ascontiguousarray(x[:n])), _dp(y))
    return y
