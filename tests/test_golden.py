"""Golden fixtures (tests/golden/*.npz, generator committed): CPU test pins the oracle + harness against them;
the GPU test compares the CUDA path with the same fixtures without running the oracle at all."""
import hashlib
import os

import numpy as np
import pytest

from util import split_blocks, synthetic_fields

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "shell_r2_classic": (dict(geometry="shell", refine=2), "shell_3d_classic"),
    "cube_r2_classic": (dict(geometry="cube", refine=2), "cube_3d"),
    "shell_r1_classic_Tq2": (dict(geometry="shell", refine=1, temperature_degree=2), "shell_3d_classic"),
    "annulus_r3_classic_2d": (dict(geometry="annulus", dim=2, refine=3, R0=10.0, R1=30.0, temperature_degree=2), "annulus_2d"),
}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check(name, vals, G, tol):
    scale = max(float(G[name + ".max"]), 1e-300)
    idx = G[name + ".sample_idx"]
    if idx.size:
        assert np.abs(vals[idx] - G[name + ".sample_val"]).max() <= tol * scale, name
    assert abs(np.sqrt((vals ** 2).sum()) - float(G[name + ".l2"])) <= tol * max(float(G[name + ".l2"]), 1e-300), name
    assert abs(vals.sum() - float(G[name + ".sum"])) <= tol * scale * max(np.sqrt(vals.size), 1.0), name


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_and_harness_match_golden(problem_factory, case):
    from dycore_b200 import params
    from oracle import oracle as orc
    spec, pname = CASES[case]
    G = np.load(os.path.join(HERE, "golden", case + ".npz"))
    P = problem_factory(**spec)
    mp = params.NAMED[pname]
    prm = orc.params_from(mp)
    # index maps and patterns: bit-exact
    assert (P.n_cells, P.scalar("nse.n_u"), P.scalar("nse.n_p"), P.scalar("temp.n_dofs")) == \
        (int(G["n_cells"]), int(G["n_u"]), int(G["n_p"]), int(G["n_T"]))
    for pat in ("nse.full", "pre.full", "temp.pat"):
        assert P.scalar(pat + ".nnz") == int(G[pat + ".nnz"])
        assert _sha(P[pat + ".rowptr"]) + _sha(P[pat + ".col"]) == str(G[pat + ".sha"])
    assert _sha(P["nse.l2g"]) == str(G["nse.l2g.sha"])
    assert _sha(P["nse.cs.line_dof"]) + _sha(P["nse.cs.entry_dof"]) == str(G["nse.cs.sha"])
    u, T = synthetic_fields(P)
    vals, rhs = orc.assemble_nse_system(P, prm, u, T)
    for (bi, bj), v in split_blocks(P, "nse", vals).items():
        _check(f"nse.b{bi}{bj}", v, G, 1e-13)
    _check("nse.rhs", rhs, G, 1e-13)
    m, k = orc.assemble_temperature_matrix(P, prm)
    _check("temp.mass", m, G, 1e-13)
    _check("temp.stiff", k, G, 1e-13)
    _check("temp.rhs", orc.assemble_temperature_rhs(P, prm, T, u), G, 1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_cuda_path_matches_golden(problem_factory, case):
    from dycore_b200 import device, params
    spec, pname = CASES[case]
    G = np.load(os.path.join(HERE, "golden", case + ".npz"))
    P = problem_factory(**spec)
    mp = params.NAMED[pname]
    u, T = synthetic_fields(P)
    ctx = device.Context(0)
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    tol = 1e-12
    model.assemble_nse_system(u, T)
    for bi in range(2):
        for bj in range(2):
            _check(f"nse.b{bi}{bj}", model.nse_matrix.block(bi, bj).values(), G, tol)
    _check("nse.rhs", model.nse_rhs, G, tol)
    model.assemble_nse_preconditioner()
    for bi in range(2):
        for bj in range(2):
            _check(f"pre.b{bi}{bj}", model.nse_preconditioner_matrix.block(bi, bj).values(), G, tol)
    model.assemble_temperature_matrix()
    _check("temp.mass", model.temperature_mass_matrix.values(), G, tol)
    _check("temp.stiff", model.temperature_stiffness_matrix.values(), G, tol)
    model.assemble_temperature_rhs(T, u)
    _check("temp.matrix", model.temperature_matrix.values(), G, tol)
    _check("temp.rhs", model.temperature_rhs, G, tol)
    n = P.scalar("nse.n_dofs")
    y = np.zeros(n)
    model.nse_matrix.vmult(y, np.random.default_rng(1).standard_normal(n))
    _check("spmv.nse", y, G, 1e-11)
    nT = P.scalar("temp.n_dofs")
    yt = np.zeros(nT)
    model.temperature_matrix.vmult(yt, np.random.default_rng(2).standard_normal(nT))
    _check("spmv.temp", yt, G, 1e-11)
    model.close()
    ctx.close()


FEEC_CASES = {"shell_r2_feec": (dict(geometry="shell", refine=2, family="feec"), "shell_3d_feec"),
              "cube_r2_feec": (dict(geometry="cube", refine=2, family="feec"), "cube_3d")}


def _feec_fields(P):
    rng = np.random.default_rng(20261018)
    return (np.ascontiguousarray(0.1 * rng.uniform(-1, 1, P.scalar("nse.n_dofs")) + 0.05),
            np.ascontiguousarray(2.0 + 0.3 * rng.uniform(-1, 1, P.scalar("temp.n_dofs"))))


def _full_from_blocks(P, model_matrix, prefix):
    """Device blocks -> values in `<prefix>.full` order (three blocks)."""
    rp, col = P[prefix + ".full.rowptr"], P[prefix + ".full.col"]
    n, nw, nu = P.scalar("nse.n_dofs"), P.scalar("nse.n_w"), P.scalar("nse.n_u")
    start = [0, nw, nw + nu, n]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    out = np.zeros(len(col))
    for bi in range(3):
        for bj in range(3):
            m = (rows >= start[bi]) & (rows < start[bi + 1]) & (col >= start[bj]) & (col < start[bj + 1])
            out[m] = model_matrix.block(bi, bj).values()
    return out


@pytest.mark.parametrize("case", sorted(FEEC_CASES))
def test_feec_oracle_and_harness_match_golden(problem_factory, case):
    from dycore_b200 import params
    from oracle import oracle as orc
    spec, pname = FEEC_CASES[case]
    G = np.load(os.path.join(HERE, "golden", case + ".npz"))
    P = problem_factory(**spec)
    prm = orc.params_from(params.NAMED[pname])
    assert (P.n_cells, P.scalar("nse.n_w"), P.scalar("nse.n_u"), P.scalar("nse.n_p")) == \
        (int(G["n_cells"]), int(G["n_w"]), int(G["n_u"]), int(G["n_p"]))
    for pat in ("nse.full", "pre.full", "temp.pat"):
        assert _sha(P[pat + ".rowptr"]) + _sha(P[pat + ".col"]) == str(G[pat + ".sha"])
    assert _sha(P["nse.l2g"]) == str(G["nse.l2g.sha"]) and _sha(P["nse.sign"]) == str(G["nse.sign.sha"])
    u, T = _feec_fields(P)
    vals, rhs = orc.feec_assemble_nse_system(P, prm, u, T)
    _check("nse.full", vals, G, 1e-13)
    _check("nse.rhs", rhs, G, 1e-13)
    _check("pre.full", orc.feec_assemble_nse_preconditioner(P, prm), G, 1e-13)
    _check("temp.rhs", orc.feec_assemble_temperature_rhs(P, prm, T, u), G, 1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(FEEC_CASES))
def test_feec_cuda_path_matches_golden(problem_factory, case):
    from dycore_b200 import device, params
    spec, pname = FEEC_CASES[case]
    G = np.load(os.path.join(HERE, "golden", case + ".npz"))
    P = problem_factory(**spec)
    u, T = _feec_fields(P)
    ctx = device.Context(0)
    model = device.BoussinesqModel.from_problem(ctx, P, params.NAMED[pname])
    model.assemble_nse_system(u, T)
    _check("nse.full", _full_from_blocks(P, model.nse_matrix, "nse"), G, 1e-12)
    _check("nse.rhs", model.nse_rhs, G, 1e-12)
    model.assemble_nse_preconditioner()
    _check("pre.full", _full_from_blocks(P, model.nse_preconditioner_matrix, "pre"), G, 1e-12)
    model.assemble_temperature_matrix()
    _check("temp.mass", model.temperature_mass_matrix.values(), G, 1e-12)
    _check("temp.stiff", model.temperature_stiffness_matrix.values(), G, 1e-12)
    model.assemble_temperature_rhs(T, u)
    _check("temp.rhs", model.temperature_rhs, G, 1e-12)
    n = P.scalar("nse.n_dofs")
    y = np.zeros(n)
    model.nse_matrix.vmult(y, np.random.default_rng(1).standard_normal(n))
    _check("spmv.nse", y, G, 1e-11)
    model.close()
    ctx.close()
