"""Shared helpers for the parity tests (comparators, field generators, block splitting)."""
import numpy as np


def split_blocks(P, prefix, vals, nb=2):
    """Values of the block-concatenated matrix `<prefix>.full` -> {(bi,bj): values in block-pattern order}."""
    rp, col = P[prefix + ".full.rowptr"], P[prefix + ".full.col"]
    n_u = P.scalar("nse.n_u")
    n = P.scalar("nse.n_dofs")
    start = [0, n_u, n]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    out = {}
    for bi in range(nb):
        for bj in range(nb):
            m = (rows >= start[bi]) & (rows < start[bi + 1]) & (col >= start[bj]) & (col < start[bj + 1])
            out[(bi, bj)] = vals[m]
            assert out[(bi, bj)].size == P.scalar(f"{prefix}.b{bi}{bj}.nnz")
    return out


def rel_err_max(a, b):
    """max|a-b| / max|b| : the norm-relative comparator of SURVEY.md 8c(5)."""
    if a.size == 0 and b.size == 0:
        return 0.0
    assert a.shape == b.shape
    denom = np.abs(b).max()
    return float(np.abs(a - b).max() / (denom if denom > 0 else 1.0))


def rel_err_rows(a, b, rowptr):
    """max over rows of max|a_row - b_row| / max|b_row|: every row is judged against its own scale, so a row whose
    entries are 1e-6 of the block maximum cannot hide a 1e-6 relative error."""
    if a.size == 0:
        return 0.0
    rowptr = np.asarray(rowptr, dtype=np.int64)
    nz = np.flatnonzero(np.diff(rowptr) > 0)
    starts = rowptr[nz]
    scale = np.maximum.reduceat(np.abs(b), starts)
    err = np.maximum.reduceat(np.abs(a - b), starts)
    ok = scale > 0
    worst = float((err[ok] / scale[ok]).max()) if ok.any() else 0.0
    assert not (err[~ok] > 0).any(), "non-zero entries in a row the oracle leaves empty"
    return worst


def synthetic_fields(P, seed=20261018, amplitude=0.1):
    """Seeded smooth state: low-order polynomials of the support point + uniform perturbation,
    so that advection / buoyancy terms are exercised.  Returns (nse_vector, temperature_vector)."""
    rng = np.random.default_rng(seed)
    dim = P.dim
    n = P.scalar("nse.n_dofs")
    x = P["nse.dof_xyz"].reshape(n, dim)
    comp = P["nse.dof_comp"]
    u = np.zeros(n)
    poly = [lambda p: 0.3 * p[:, 1] - 0.2 * p[:, 0] * p[:, -1], lambda p: -0.25 * p[:, 0] + 0.1 * p[:, 1] ** 2,
            lambda p: 0.15 * p[:, 0] * p[:, 1] - 0.05 * p[:, -1], lambda p: 0.5 + 0.1 * p[:, 0]]
    for c in range(dim + 1):
        m = comp == c
        f = poly[c if c < dim else 3]
        u[m] = f(x[m])
    u += amplitude * rng.uniform(-1, 1, n)
    nT = P.scalar("temp.n_dofs")
    xt = P["temp.dof_xyz"].reshape(nT, dim)
    T = 2.0 + 0.3 * xt[:, 0] - 0.1 * xt[:, 1] * xt[:, -1] + amplitude * rng.uniform(-1, 1, nT)
    return np.ascontiguousarray(u), np.ascontiguousarray(T)
