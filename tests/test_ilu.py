"""ILU(0): the oracle restatement against its defining identities, the device factorisation and
triangular solves against the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp


def _block00(P, vals):
    from util import split_blocks
    rp, col = P["nse.b00.rowptr"], P["nse.b00.col"]
    return rp, col, split_blocks(P, "nse", vals)[(0, 0)]


def test_oracle_ilu0_identities(problem_factory):
    """(L U)_ij = A_ij on the pattern of A (the defining property of ILU(0)); on a pattern without fill (tridiagonal)
    the factorisation is exact and the solve inverts A."""
    from dycore_b200 import params
    from oracle import oracle as orc
    from util import synthetic_fields
    P = problem_factory(geometry="annulus", dim=2, R0=10.0, R1=30.0, temperature_degree=2, refine=2)
    prm = orc.params_from(params.NAMED["annulus_2d"])
    u, T = synthetic_fields(P)
    vals, _ = orc.assemble_nse_system(P, prm, u, T)
    rp, col, v00 = _block00(P, vals)
    n = len(rp) - 1
    lu = orc.ilu0_factor(rp, col, v00)
    A = sp.csr_matrix((v00, col, rp), shape=(n, n))
    F = sp.csr_matrix((lu, col, rp), shape=(n, n))
    L = sp.tril(F, -1) + sp.identity(n)
    U = sp.triu(F, 0)
    R = (L @ U - A).tocsr()
    pattern = sp.csr_matrix((np.ones_like(v00), col, rp), shape=(n, n))
    on_pattern = R.multiply(pattern)
    assert abs(on_pattern).max() <= 1e-13 * abs(A).max()
    x = np.random.default_rng(0).standard_normal(n)
    y = orc.ilu0_solve(rp, col, lu, x)
    assert np.abs(L @ (U @ y) - x).max() <= 1e-10 * np.abs(x).max()
    # tridiagonal: no fill, ILU(0) == LU
    m = 50
    T3 = sp.diags([-1.0 * np.ones(m - 1), 2.5 * np.ones(m), -1.2 * np.ones(m - 1)], [-1, 0, 1]).tocsr()
    T3.sort_indices()
    lu3 = orc.ilu0_factor(T3.indptr.astype(np.int64), T3.indices.astype(np.int32), T3.data)
    b = np.arange(1.0, m + 1)
    z = orc.ilu0_solve(T3.indptr.astype(np.int64), T3.indices.astype(np.int32), lu3, b)
    assert np.abs(T3 @ z - b).max() <= 1e-12 * np.abs(b).max()


@pytest.mark.gpu
@pytest.mark.parametrize("spec,pname", [(dict(geometry="annulus", dim=2, R0=10.0, R1=30.0, temperature_degree=2, refine=3), "annulus_2d"),
                                        (dict(geometry="shell", refine=2), "shell_3d_classic")], ids=["annulus-r3", "shell-r2"])
def test_device_ilu_matches_oracle(problem_factory, spec, pname):
    import torch
    import dycore_b200  # noqa: F401
    from dycore_b200 import device, params
    from oracle import oracle as orc
    from util import synthetic_fields
    P = problem_factory(**spec)
    mp = params.NAMED[pname]
    u, T = synthetic_fields(P)
    ctx = device.Context(0)
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    model.assemble_nse_system(u, T)
    ilu = device.PreconditionILU(model, device.MAT_NSE, 0)
    rp, col = P["nse.b00.rowptr"], P["nse.b00.col"]
    v00 = model.nse_matrix.block(0, 0).values()      # factorise the SAME values on the CPU
    lu = orc.ilu0_factor(rp, col, v00)
    n = len(rp) - 1
    x = np.random.default_rng(5).standard_normal(n)
    ref = orc.ilu0_solve(rp, col, lu, x)
    y = np.zeros(n)
    ilu.vmult(y, x)
    assert np.abs(y - ref).max() <= 1e-11 * np.abs(ref).max()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)   # order torch's fills / copies with the library's kernels
    dx, dy = torch.from_numpy(x).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda")
    ilu.vmult(dy, dx)
    assert np.array_equal(dy.cpu().numpy(), y)
    nl, nu = ilu.levels()
    assert 1 <= nl <= n and 1 <= nu <= n
    # refactor after a re-assembly with another time step: the factors follow the matrix
    model.prm.dt *= 0.5
    model.assemble_nse_system(u, T)
    ilu.refactor()
    lu2 = orc.ilu0_factor(rp, col, model.nse_matrix.block(0, 0).values())
    ilu.vmult(y, x)
    assert np.abs(y - orc.ilu0_solve(rp, col, lu2, x)).max() <= 1e-11 * np.abs(ref).max()
    ilu.close()
    model.close()
    ctx.close()
