"""Host logic of the Krylov mirror (3d-dycoreplanet_b200/solvers.py) on the numpy backend: the restated deal.II
solvers against direct solves, and the two Stokes solve chains of the reference on the CPU (oracle assembly).
The same code runs on device vectors in tests/test_gpu_krylov.py."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import krylov_util as K


def _ops(S, M):
    mat = S.Wrap(lambda dst, src: dst.__setitem__(slice(None), M @ src))
    d = M.diagonal()
    jac = S.Wrap(lambda dst, src: dst.__setitem__(slice(None), src / d))
    return mat, jac


def test_cg_and_gmres_match_direct_solves():
    import dycore_b200  # noqa: F401
    from dycore_b200 import solvers as S
    B = S.NumpyBackend()
    rng = np.random.default_rng(0)
    n = 200
    A = sp.diags([-1.0 * np.ones(n - 1), 2.2 * np.ones(n), -1.0 * np.ones(n - 1)], [-1, 0, 1]).tocsr()
    b = rng.standard_normal(n)
    ref = spla.spsolve(A.tocsc(), b)
    mat, jac = _ops(S, A)
    x = np.zeros(n)
    its = S.solver_cg(B, mat, x, b, jac, 1e-12 * np.linalg.norm(b), 10 * n)
    assert 0 < its <= n and np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max()
    # nonsymmetric: GMRES (left preconditioned) and FGMRES (right preconditioned, as boussinesq_model.tpp:1191-1199)
    N = (A + sp.diags([0.4 * np.ones(n - 1)], [1])).tocsr()
    refn = spla.spsolve(N.tocsc(), b)
    matn, jacn = _ops(S, N)
    for flexible in (False, True):
        x = np.zeros(n)
        its = S.solver_gmres(B, matn, x, b, jacn, 1e-11 * np.linalg.norm(b), 2000, restart=28, flexible=flexible)
        assert its > 0 and np.abs(x - refn).max() <= 1e-7 * np.abs(refn).max(), flexible
    # SolverControl semantics: a tolerance that cannot be met raises like SolverControl::NoConvergence
    with pytest.raises(S.NoConvergence):
        S.solver_cg(B, mat, np.zeros(n), b, jac, 1e-300, 3)


def test_inverse_matrix_and_schur_operators():
    import dycore_b200  # noqa: F401
    from dycore_b200 import solvers as S
    B = S.NumpyBackend()
    rng = np.random.default_rng(1)
    n_u, n_p = 60, 20
    A = sp.diags([-1.0 * np.ones(n_u - 1), 3.0 * np.ones(n_u), -1.0 * np.ones(n_u - 1)], [-1, 0, 1]).tocsr()
    Bt = sp.random(n_u, n_p, density=0.2, random_state=3, format="csr")
    matA, jacA = _ops(S, A)
    inv = S.InverseMatrix(matA, jacA)                      # inverse_matrix.hpp:90-121: CG to 1e-6 |src|
    r = rng.standard_normal(n_u)
    y = np.zeros(n_u)
    inv.vmult(y, r, B)
    assert np.linalg.norm(A @ y - r) <= 1e-6 * np.linalg.norm(r) * 1.0001 and inv.iterations[-1] > 0
    wrap = lambda M: S.Wrap(lambda dst, src, M=M: dst.__setitem__(slice(None), M @ src))  # noqa: E731
    schur = S.SchurComplement(wrap(Bt), wrap(Bt.T.tocsr()), inv, n_u, B)   # schur_complement.hpp:143-150
    p = rng.standard_normal(n_p)
    out = np.zeros(n_p)
    schur.vmult(out, p, B)
    exact = Bt.T @ spla.spsolve(A.tocsc(), Bt @ p)
    assert np.abs(out - exact).max() <= 1e-4 * max(np.abs(exact).max(), 1e-30)


@pytest.mark.parametrize("which", ["block_preconditioned", "schur_complement"])
def test_stokes_solve_chains_reduce_the_residual(problem_factory, which):
    """One Stokes solve of the reference's two solver chains on the CPU mirror: the solution satisfies the assembled
    system to the solver tolerance (block-preconditioned FGMRES: 1e-8 |rhs|, boussinesq_model.tpp:1165)."""
    import dycore_b200  # noqa: F401
    from dycore_b200 import params
    from oracle import oracle as orc
    if which == "block_preconditioned":
        mp = params.NAMED["shell_3d_classic"]
        P = problem_factory(geometry="shell", refine=1)
        step = K.cpu_time_step
    else:
        mp = params.NAMED["annulus_2d"]
        P = problem_factory(geometry="annulus", dim=2, R0=10.0, R1=30.0, temperature_degree=2, refine=2,
                            renumber="cuthill_mckee")
        step = K.cpu_schur_step
    n, n_u = P.scalar("nse.n_dofs"), P.scalar("nse.n_u")
    u0 = np.zeros(n)
    T0 = K.initial_temperature(P, mp)
    out = step(P, mp, u0, T0)
    vals, rhs = orc.assemble_nse_system(P, orc.params_from(mp), u0, T0)
    rp, col, _, _ = P.csr("nse.full")
    A = sp.csr_matrix((vals, col, rp), shape=(n, n))
    x = out["nse"].copy()
    x[n_u:] *= mp.time_step                     # the system is solved for the dt-scaled pressure (:1151, 1283)
    free = np.ones(n, bool)                     # constrained rows are placeholders (diagonal only); distribute() overwrote
    free[P["nse.cs.line_dof"]] = False          # their entries with the constraint values
    res = np.linalg.norm((A @ x - rhs)[free]) / np.linalg.norm(rhs)
    assert res <= (1e-7 if which == "block_preconditioned" else 1e-4), res
    assert out["fgmres" if which == "block_preconditioned" else "gmres"] >= 1
    assert math.isfinite(np.abs(out["nse"]).max())
