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


@pytest.mark.parametrize("refine", [1, 2])
def test_feec_block_preconditioned_chain_on_cpu(problem_factory, refine):
    """ExteriorCalculus::BoussinesqModel::solve_NSE_block_preconditioned (boussineq_model_FEEC.tpp:1268-1477) on the
    numpy backend: BlockSchurPreconditionerFEEC with the shifted / nested Schur-complement inverses converges within
    the reference's 500 steps, the result solves the assembled system, and the preconditioner's third block carries
    the -2 src_p quirk (block_schur_preconditioner.hpp:137-143)."""
    import dycore_b200  # noqa: F401
    from dycore_b200 import params
    from dycore_b200 import solvers as S
    from oracle import oracle as orc
    mp = params.NAMED["shell_3d_feec"]
    P = problem_factory(geometry="shell", refine=refine, family="feec")
    n = P.scalar("nse.n_dofs")
    nw, nu = P.scalar("nse.n_w"), P.scalar("nse.n_u")
    u0, T0 = np.zeros(n), K.initial_temperature(P, mp)
    out = K.cpu_feec_step(P, mp, u0, T0)
    assert 0 < out["gmres"] <= 500
    assert len(out["inner"]["shifted"]) == len(out["inner"]["nested"]) >= out["gmres"]
    assert max(out["inner"]["shifted"]) <= 30 and max(out["inner"]["nested"]) <= 100
    # the returned field (pressure scaled back, constraints distributed) solves the assembled system on the free dofs
    vals, rhs = orc.feec_assemble_nse_system(P, orc.params_from(mp), u0, T0)
    rp, col, _, _ = P.csr("nse.full")
    A = sp.csr_matrix((vals, col, rp), shape=(n, n))
    x = out["nse"].copy()
    x[nw + nu:] *= mp.time_step
    free = np.ones(n, bool)
    free[P["nse.cs.line_dof"]] = False
    res = (A @ x - rhs)[free]
    assert np.linalg.norm(res) <= 1e-6 * np.linalg.norm(rhs)
    # zero-mean weights: a constant pressure has mean 1 under both quadratures
    for ng in (1, 2):
        w = K.feec_mean_weights(P, ng)
        assert abs(w.sum() - 1.0) <= 1e-14 and (w > 0).all()
    B = S.NumpyBackend()
    p = np.full(n - nw - nu, 3.5)
    assert abs(S.MeanValue(K.feec_mean_weights(P, 1), B).subtract(p, B) - 3.5) <= 1e-13 and np.abs(p).max() <= 1e-13


def test_feec_preconditioner_third_block_quirk():
    """BlockSchurPreconditionerFEEC::vmult with identity inverses: dst_p = -2 src_p + B21 dst_u (Q8)."""
    import dycore_b200  # noqa: F401
    from dycore_b200 import solvers as S
    B = S.NumpyBackend()
    rng = np.random.default_rng(5)
    nw, nu, npr = 7, 5, 3
    M10, M21 = rng.standard_normal((nu, nw)), rng.standard_normal((npr, nu))

    class Mat:
        def __init__(self, M):
            self.M = M

        def vmult(self, dst, src, B=None):
            dst[...] = self.M @ src

        def vmult_add(self, dst, src, B=None):
            dst += self.M @ src
    ident = S.Identity()
    Pf = S.BlockSchurPreconditionerFEEC({(1, 0): Mat(M10), (2, 1): Mat(M21)}, ident, ident, ident, (nw, nu, npr), B)
    src = rng.standard_normal(nw + nu + npr)
    dst = np.full(nw + nu + npr, np.nan)                 # an uninitialised destination must not leak into the result
    Pf.vmult(dst, src, B)
    dw = src[:nw]
    du = src[nw:nw + nu] - M10 @ dw
    dp = -2.0 * src[nw + nu:] + M21 @ du
    assert np.allclose(dst, np.concatenate([dw, du, dp]), rtol=0, atol=1e-13)


def test_block_solve_falls_back_to_do_solve_A():
    """boussinesq_model.tpp:1203-1232: NoConvergence of the 40-step FGMRES(30) is caught, the solve continues with
    do_solve_A = true and FGMRES(50), and the reported count is the sum."""
    import dycore_b200  # noqa: F401
    from dycore_b200 import solvers as S
    B = S.NumpyBackend()
    rng = np.random.default_rng(2)
    n_u, n_p = 80, 20
    A00 = sp.diags([-1.0 * np.ones(n_u - 1), 2.05 * np.ones(n_u), -1.0 * np.ones(n_u - 1)], [-1, 0, 1]).tocsr()
    Bt = sp.random(n_u, n_p, density=0.15, random_state=4, format="csr")
    Afull = sp.bmat([[A00, Bt], [Bt.T, None]]).tocsr()
    mat = lambda M: S.Wrap(lambda dst, src, M=M: dst.__setitem__(slice(None), M @ src))
    d = A00.diagonal()
    jac = S.Wrap(lambda dst, src: dst.__setitem__(slice(None), src / d))
    blocks = {(0, 0): mat(A00), (0, 1): mat(Bt), (1, 0): mat(Bt.T.tocsr())}
    rhs = rng.standard_normal(n_u + n_p)
    x, its, inner = S.solve_nse_block_preconditioned(B, mat(Afull), blocks, jac, rhs, np.zeros(n_u + n_p), n_u, n_p, 1.0,
                                                     max_steps=2)
    assert its > 2, "the fall-back's steps are added to the first solve's"
    assert np.linalg.norm(Afull @ x - rhs) <= 1e-8 * np.linalg.norm(rhs) * 1.01
    # constrained pressure dofs are zeroed in the initial guess (:1160-1162)
    x0 = np.ones(n_u + n_p)
    x2, _, _ = S.solve_nse_block_preconditioned(B, mat(Afull), blocks, jac, rhs, x0, n_u, n_p, 1.0, constrained_pressure=[0, 3])
    assert np.linalg.norm(Afull @ x2 - rhs) <= 1e-8 * np.linalg.norm(rhs) * 1.01
