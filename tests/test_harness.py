"""Host logic of the stand-in harness: counts from SURVEY.md 8, pattern / constraint invariants, partition tiling."""
import numpy as np
import pytest


def test_shell_counts_match_survey(problem_factory):
    P = problem_factory(geometry="shell", refine=2)
    assert (P.n_cells, P.scalar("nse.n_u"), P.scalar("nse.n_p"), P.scalar("temp.n_dofs")) == (384, 10422, 490, 490)
    Q = problem_factory(geometry="shell", refine=2, constraints=0)
    assert Q.scalar("nse.full.nnz") == 2128014          # NSE nnz (unconstrained) of the named config
    assert problem_factory(geometry="shell", refine=1, temperature_degree=2).scalar("temp.n_dofs") == 490


@pytest.mark.parametrize("spec", [dict(geometry="shell", refine=2), dict(geometry="cube", refine=2)])
def test_patterns_are_sorted_and_constrained_rows_are_diagonal(problem_factory, spec):
    P = problem_factory(**spec)
    for name in ("nse.full", "pre.full", "temp.pat"):
        rp, col, n, _ = P.csr(name)
        rows = np.repeat(np.arange(n), np.diff(rp))
        d = np.diff(col.astype(np.int64))
        same_row = rows[1:] == rows[:-1]
        assert (d[same_row] > 0).all(), "columns must ascend within a row"
    rp, col, n, _ = P.csr("nse.full")
    lod = P["nse.cs.line_of_dof"]
    con = np.flatnonzero(lod >= 0)
    assert (np.diff(rp)[con] == 1).all() and (col[rp[con]] == con).all()
    assert not np.isin(col, con)[np.repeat(lod < 0, np.diff(rp))].any(), "constrained columns are eliminated"
    # velocity triplets are adjacent dofs (what the position-table path relies on)
    l2g = P["nse.l2g"].reshape(P.n_cells, -1)
    f, b = P["nse.local_field"], P["nse.local_base"]
    for a in range(27):
        i0, i1, i2 = [np.flatnonzero((f == c) & (b == a))[0] for c in range(3)]
        assert (l2g[:, i1] == l2g[:, i0] + 1).all() and (l2g[:, i2] == l2g[:, i0] + 2).all()


def test_partition_tiles_the_global_dofs(problem_factory):
    G = problem_factory(geometry="shell", refine=2)
    keys = []
    for p in range(4):
        P = problem_factory(geometry="shell", refine=2, n_ranks=4, rank=p)
        own = P["nse.dof_owner"] == p
        keys.append(P["nse.dof_key"][own])
        n_u, n_uo, n_po = P.scalar("nse.n_u"), P.scalar("nse.n_u_owned"), P.scalar("nse.n_p_owned")
        idx = np.flatnonzero(own)
        assert (idx == np.concatenate([np.arange(n_uo), n_u + np.arange(n_po)])).all(), "owned dofs come first per block"
    allk = np.concatenate(keys)
    assert len(allk) == G.scalar("nse.n_dofs") and len(np.unique(allk)) == len(allk)
    assert (np.sort(allk) == np.sort(G["nse.dof_key"])).all()


def test_mapping_data_positive_and_volume_converges(problem_factory):
    exact = 4.0 / 3.0 * np.pi * (27.0 - 1.0)
    errs = []
    for r in (1, 2, 3):
        P = problem_factory(geometry="shell", refine=r)
        g = P["geom.qn"].reshape(P.n_cells, 13, 27)
        assert (g[:, 0, :] > 0).all()
        errs.append(abs(g[:, 0, :].sum() - exact) / exact)
    assert errs[2] < errs[1] < 0.05


@pytest.mark.parametrize("mode", ["cuthill_mckee", "random"])
def test_renumbering_permutes_the_system(problem_factory, mode):
    """`renumber` (the reference applies DoFRenumbering::Cuthill_McKee before component_wise on the Schur path,
    boussinesq_model.tpp:198-202) only permutes the dofs inside their blocks: the oracle's right-hand side and
    matrix-vector product agree entry by entry when matched through the numbering-independent dof keys."""
    import numpy as np
    from dycore_b200 import params
    from oracle import oracle as orc
    spec = dict(geometry="annulus", dim=2, R0=10.0, R1=30.0, temperature_degree=2, refine=2)
    P0, P1 = problem_factory(**spec), problem_factory(renumber=mode, **spec)
    k0, k1 = P0["nse.dof_key"], P1["nse.dof_key"]
    assert sorted(k0) == sorted(k1) and not np.array_equal(k0, k1)
    assert P0.scalar("nse.n_u") == P1.scalar("nse.n_u")
    prm = orc.params_from(params.NAMED["annulus_2d"])

    def field(keys, s):
        return np.ascontiguousarray(np.sin(0.37 * s * (keys % 1013)) + 0.2 * np.cos(0.11 * (keys % 7919)))
    T0 = np.ascontiguousarray(2.0 + 0.1 * np.sin(0.5 * (P0["temp.dof_key"] % 101)))
    out = []
    for P, k in ((P0, k0), (P1, k1)):
        vals, rhs = orc.assemble_nse_system(P, prm, field(k, 1.0) * 0.1, T0)
        rp, col, _, _ = P.csr("nse.full")
        y = orc.spmv(rp, col, vals, field(k, 2.0))
        order = np.argsort(k)
        out.append((rhs[order], y[order]))
    assert np.abs(out[0][0] - out[1][0]).max() <= 1e-13 * np.abs(out[0][0]).max()
    assert np.abs(out[0][1] - out[1][1]).max() <= 1e-12 * np.abs(out[0][1]).max()
