"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): sparsity pattern / index maps are inputs shared bit-exactly; assembled
matrix, RHS and SpMV within 1e-12 relative in FP64 (norm-relative comparator)."""
import numpy as np
import pytest

from util import rel_err_max, rel_err_rows, split_blocks, synthetic_fields

TOL = 1e-12
ROW_TOL = 1e-11   # per-row-scaled comparator: every row against its own largest entry (rows of small cells count like the others)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import dycore_b200  # noqa: F401
    from dycore_b200 import device
    c = device.Context(0)
    yield c
    c.close()


ANNULUS = dict(geometry="annulus", dim=2, R0=10.0, R1=30.0, temperature_degree=2)   # data/aqua_planet_test_2d.prm
CASES = [dict(geometry="shell", refine=1), dict(geometry="shell", refine=2), dict(geometry="cube", refine=2),
         dict(geometry="shell", refine=1, temperature_degree=2), dict(refine=2, **ANNULUS), dict(refine=4, **ANNULUS),
         dict(geometry="shell", refine=3), dict(geometry="cube", refine=3)]


def _params(spec):
    from dycore_b200 import params
    return params.NAMED[{"cube": "cube_3d", "annulus": "annulus_2d"}.get(spec["geometry"], "shell_3d_classic")]


@pytest.mark.parametrize("spec", CASES, ids=lambda s: "-".join(f"{k}{v}" for k, v in s.items()))
@pytest.mark.parametrize("strategy", [0, 1, 2, 3, 4, 5, 6],
                         ids=["search", "positions", "owner", "staged", "staged-chunk37", "staged-unfused", "staged-chunk1000"])
def test_classic_assembly_matches_oracle(ctx, problem_factory, spec, strategy, monkeypatch):
    from dycore_b200 import device
    from oracle import oracle as orc
    if spec["refine"] == 3 and spec["geometry"] != "annulus" and strategy in (0, 2):
        pytest.skip("refine 3: the default and the staged strategies only (run time)")
    P = problem_factory(**spec)
    mp = _params(spec)
    if strategy == 2 and P.dim == 2:
        pytest.skip("row-owner tiles are built for the 3-D family only")
    if strategy >= 3 and spec["geometry"] != "shell":
        pytest.skip("write-once staging: classic 3-D family with every cell in the position plan (the shell)")
    u, T = synthetic_fields(P)
    if strategy == 4:   # many small chunks: nodes are visited by several chunks (store first, accumulate later)
        monkeypatch.setenv("DCP_GATHER_CHUNK", "37")
    if strategy == 5:   # the preconditioner in its own pass: the gather writes nse_matrix only
        monkeypatch.setenv("DCP_NO_FUSED_PRECONDITIONER", "1")
    if strategy == 6:   # a few chunks with a ragged last one
        monkeypatch.setenv("DCP_GATHER_CHUNK", "1000")
    model = device.BoussinesqModel.from_problem(ctx, P, mp, owner_plan=(strategy == 2))
    if spec["geometry"] == "shell":
        assert model.strategy == device.STRATEGY_STAGED, "the shell qualifies for the write-once path by default"
    model.set_strategy(min(strategy, 3))
    oprm = orc.params_from(mp)

    # NSE system
    model.assemble_nse_system(u, T)
    ref_vals, ref_rhs = orc.assemble_nse_system(P, oprm, u, T)
    ref_blocks = split_blocks(P, "nse", ref_vals)
    for (bi, bj), rv in ref_blocks.items():
        gv = model.nse_matrix.block(bi, bj).values()
        assert rel_err_max(gv, rv) <= TOL, f"nse block {bi}{bj}"
        assert rel_err_rows(gv, rv, P[f"nse.b{bi}{bj}.rowptr"]) <= ROW_TOL, f"nse block {bi}{bj} (row-scaled)"
    assert rel_err_max(model.nse_rhs, ref_rhs) <= TOL

    # NSE preconditioner
    model.assemble_nse_preconditioner()
    ref_blocks = split_blocks(P, "pre", orc.assemble_nse_preconditioner(P, oprm))
    for (bi, bj), rv in ref_blocks.items():
        gv = model.nse_preconditioner_matrix.block(bi, bj).values()
        assert rel_err_max(gv, rv) <= TOL, f"pre block {bi}{bj}"
        assert rel_err_rows(gv, rv, P[f"pre.b{bi}{bj}.rowptr"]) <= ROW_TOL, f"pre block {bi}{bj} (row-scaled)"

    # every time step re-assembles (quirk Q16): a second pass must not accumulate onto the first
    model.assemble_nse_system(u, T)
    model.assemble_nse_preconditioner()
    for (bi, bj), rv in split_blocks(P, "nse", ref_vals).items():
        assert rel_err_max(model.nse_matrix.block(bi, bj).values(), rv) <= TOL, f"second pass, nse block {bi}{bj}"
    assert rel_err_max(model.nse_rhs, ref_rhs) <= TOL
    for (bi, bj), rv in ref_blocks.items():
        assert rel_err_max(model.nse_preconditioner_matrix.block(bi, bj).values(), rv) <= TOL, f"second pass, pre block {bi}{bj}"

    # temperature matrices, combined matrix and rhs
    model.assemble_temperature_matrix()
    rm, rk = orc.assemble_temperature_matrix(P, oprm)
    assert rel_err_max(model.temperature_mass_matrix.values(), rm) <= TOL
    assert rel_err_max(model.temperature_stiffness_matrix.values(), rk) <= TOL
    model.assemble_temperature_rhs(T, u)
    rt = orc.temperature_matrix_combine(rm, rk, mp.time_step / mp.NSE_solver_interval)
    assert rel_err_max(model.temperature_matrix.values(), rt) <= TOL
    assert rel_err_max(model.temperature_rhs, orc.assemble_temperature_rhs(P, oprm, T, u)) <= TOL
    model.close()


@pytest.mark.parametrize("spec", CASES[:3], ids=lambda s: "-".join(f"{k}{v}" for k, v in s.items()))
def test_spmv_matches_oracle(ctx, problem_factory, spec):
    from dycore_b200 import device
    from oracle import oracle as orc
    P = problem_factory(**spec)
    mp = _params(spec)
    u, T = synthetic_fields(P)
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    model.assemble_nse_system(u, T)
    model.assemble_temperature_matrix()
    model.assemble_temperature_rhs(T, u)
    rng = np.random.default_rng(1)
    n_u, n_p, n_t = P.scalar("nse.n_u"), P.scalar("nse.n_p"), P.scalar("temp.n_dofs")
    sizes = [n_u, n_p]
    for bi in range(2):
        for bj in range(2):
            A = model.nse_matrix.block(bi, bj)
            if A.nnz == 0:
                continue
            rp, col, _, _ = P.csr(f"nse.b{bi}{bj}")
            vals = A.values()
            for x in (rng.standard_normal(sizes[bj]), np.ones(sizes[bj])):
                y = np.zeros(sizes[bi])
                A.vmult(y, x)
                yr = orc.spmv(rp, col, vals, x)
                # scale = max(|A||x|): rows of B sum to ~0 on constants, so ||y|| itself can be round-off
                scale = orc.spmv(rp, col, np.abs(vals), np.abs(x)).max()
                assert np.abs(y - yr).max() <= TOL * scale
                y2 = rng.standard_normal(sizes[bi])
                y2r = orc.spmv(rp, col, vals, x, y=y2.copy(), add=True)
                A.vmult_add(y2, x)
                assert np.abs(y2 - y2r).max() <= TOL * max(scale, np.abs(y2r).max())
    # full block vmult and temperature matrix
    x = rng.standard_normal(n_u + n_p)
    y = np.zeros(n_u + n_p)
    model.nse_matrix.vmult(y, x)
    rp, col, _, _ = P.csr("nse.full")
    full_vals = np.zeros(len(col))
    # rebuild full values from the blocks through the oracle ordering
    ref_vals, _ = orc.assemble_nse_system(P, orc.params_from(mp), u, T)
    assert rel_err_max(y, orc.spmv(rp, col, ref_vals, x)) <= 1e-11
    xt = rng.standard_normal(n_t)
    yt = np.zeros(n_t)
    model.temperature_matrix.vmult(yt, xt)
    rp, col, _, _ = P.csr("temp.pat")
    assert rel_err_max(yt, orc.spmv(rp, col, model.temperature_matrix.values(), xt)) <= TOL
    # Jacobi
    d = np.zeros(n_t)
    model.T_preconditioner.vmult(d, xt)
    import scipy.sparse as sp
    Tm = sp.csr_matrix((model.temperature_matrix.values(), col, rp), shape=(n_t, n_t))
    assert rel_err_max(d, xt / Tm.diagonal()) <= TOL
    model.close()


def split_blocks3(P, prefix, vals):
    """FEEC: values on `<prefix>.full` -> {(bi,bj)} for the three blocks (w,u,p)."""
    rp, col = P[prefix + ".full.rowptr"], P[prefix + ".full.col"]
    n = P.scalar("nse.n_dofs")
    nw, nu = P.scalar("nse.n_w"), P.scalar("nse.n_u")
    start = [0, nw, nw + nu, n]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    out = {}
    for bi in range(3):
        for bj in range(3):
            m = (rows >= start[bi]) & (rows < start[bi + 1]) & (col >= start[bj]) & (col < start[bj + 1])
            out[(bi, bj)] = vals[m]
    return out


FEEC_CASES = [(dict(geometry="shell", refine=1, family="feec"), "shell_3d_feec"),
              (dict(geometry="shell", refine=3, family="feec"), "shell_3d_feec"),
              (dict(geometry="cube", refine=2, family="feec"), "cube_3d")]


@pytest.mark.parametrize("spec,pname", FEEC_CASES, ids=["shell-r1", "shell-r3-named", "cube-r2"])
def test_feec_assembly_and_spmv_match_oracle(ctx, problem_factory, spec, pname):
    from dycore_b200 import device, params
    from oracle import oracle as orc
    P = problem_factory(**spec)
    mp = params.NAMED[pname]
    oprm = orc.params_from(mp)
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    rng = np.random.default_rng(20261018)
    u = np.ascontiguousarray(0.1 * rng.uniform(-1, 1, n) + 0.05)
    T = np.ascontiguousarray(2.0 + 0.3 * rng.uniform(-1, 1, nT))
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    model.assemble_nse_system(u, T)
    ref_vals, ref_rhs = orc.feec_assemble_nse_system(P, oprm, u, T)
    for (bi, bj), rv in split_blocks3(P, "nse", ref_vals).items():
        assert rel_err_max(model.nse_matrix.block(bi, bj).values(), rv) <= TOL, f"feec nse block {bi}{bj}"
    assert rel_err_max(model.nse_rhs, ref_rhs) <= TOL
    model.assemble_nse_preconditioner()
    for (bi, bj), rv in split_blocks3(P, "pre", orc.feec_assemble_nse_preconditioner(P, oprm)).items():
        assert rel_err_max(model.nse_preconditioner_matrix.block(bi, bj).values(), rv) <= TOL, f"feec pre block {bi}{bj}"
    model.assemble_temperature_matrix()
    rm, rk = orc.feec_assemble_temperature_matrix(P, oprm)
    assert rel_err_max(model.temperature_mass_matrix.values(), rm) <= TOL
    assert rel_err_max(model.temperature_stiffness_matrix.values(), rk) <= TOL
    model.assemble_temperature_rhs(T, u)
    assert rel_err_max(model.temperature_rhs, orc.feec_assemble_temperature_rhs(P, oprm, T, u)) <= TOL
    # 3x3-block vmult as called by SolverGMRES (boussineq_model_FEEC.tpp:1372-1401) and the leaves of the FEEC
    # preconditioner (block_schur_preconditioner.hpp:131,144)
    x = rng.standard_normal(n)
    y = np.zeros(n)
    model.nse_matrix.vmult(y, x)
    rp, col, _, _ = P.csr("nse.full")
    scale = orc.spmv(rp, col, np.abs(ref_vals), np.abs(x)).max()
    assert np.abs(y - orc.spmv(rp, col, ref_vals, x)).max() <= 1e-12 * scale
    nw, nu = P.scalar("nse.n_w"), P.scalar("nse.n_u")
    A10 = model.nse_matrix.block(1, 0)
    y10 = np.zeros(nu)
    A10.vmult(y10, np.ascontiguousarray(x[:nw]))
    rp10, col10, _, _ = P.csr("nse.b10")
    assert np.abs(y10 - orc.spmv(rp10, col10, A10.values(), np.ascontiguousarray(x[:nw]))).max() <= 1e-12 * scale
    model.close()


DIAG_CASES = [(dict(geometry="shell", refine=2), "shell_3d_classic"), (dict(geometry="cube", refine=2), "cube_3d"),
              (dict(refine=3, **ANNULUS), "annulus_2d"), (dict(geometry="shell", refine=2, family="feec"), "shell_3d_feec"),
              (dict(geometry="cube", refine=2, family="feec"), "cube_3d")]


@pytest.mark.parametrize("spec,pname", DIAG_CASES, ids=["shell", "cube", "annulus", "shell-feec", "cube-feec"])
def test_velocity_extrema_and_distribute_match_oracle(ctx, problem_factory, spec, pname):
    """get_maximal_velocity / get_cfl_number (boussinesq_model.tpp:1023-1098) and AffineConstraints::distribute
    (:1233, 1442) on the device vs the numpy restatement; host and device vector arguments."""
    import torch
    from dycore_b200 import device, params
    from oracle import oracle as orc
    P = problem_factory(**spec)
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    rng = np.random.default_rng(7)
    x = np.ascontiguousarray(rng.uniform(-1, 1, n))
    model = device.BoussinesqModel.from_problem(ctx, P, params.NAMED[pname])
    vmax, cfl = model.velocity_extrema(x)
    rv, rc = orc.velocity_extrema(P, x)
    assert abs(vmax - rv) <= 1e-13 * rv and abs(cfl - rc) <= 1e-13 * rc
    dx = torch.from_numpy(x).cuda()
    assert model.velocity_extrema(dx) == (vmax, cfl)
    assert model.get_maximal_velocity(np.zeros(n)) == 0.0
    # distribute: masters keep their value, constrained entries are overwritten
    y = x.copy()
    model.distribute_nse_constraints(y)
    assert np.abs(y - orc.constraints_distribute(P, "nse.cs", x)).max() <= 1e-15
    model.distribute_nse_constraints(dx)
    assert np.array_equal(dx.cpu().numpy(), y)
    t = np.ascontiguousarray(rng.uniform(-1, 1, nT))
    t2 = t.copy()
    model.distribute_temperature_constraints(t2)
    assert np.abs(t2 - orc.constraints_distribute(P, "temp.cs", t)).max() <= 1e-15
    model.close()


GEOM_CASES = [(dict(geometry="shell", refine=2), "shell_3d_classic"), (dict(geometry="shell", refine=1, temperature_degree=2), "shell_3d_classic"),
              (dict(refine=3, **ANNULUS), "annulus_2d"), (dict(geometry="shell", refine=2, family="feec"), "shell_3d_feec")]


@pytest.mark.parametrize("spec,pname", GEOM_CASES, ids=["shell", "shell-Tq2", "annulus", "shell-feec"])
def test_device_geometry_matches_host_mapping(ctx, problem_factory, spec, pname):
    """dcp_geometry_create (MappingQ(3) on boundary cells, Q1 elsewhere; JxW, inverse Jacobian, x_q, J, det J) vs the
    harness' host evaluation of the same mapping, and an assembly pass running on the device-made records."""
    from dycore_b200 import device, params
    from oracle import oracle as orc
    P = problem_factory(**spec)
    feec = spec.get("family") == "feec"
    for rule, name in (("qn", "geom.qn"), ("qt", "geom.qt")) + ((("qp", "geom.qp"),) if feec else ()):
        ptr = device.geometry_create(ctx, P, rule)
        host = P[name]
        dev = ctx.download_f64(ptr, host.size)
        ctx.free(ptr)
        nq = P[f"map.{rule}.w"].size
        h, d = host.reshape(P.n_cells, -1, nq), dev.reshape(P.n_cells, -1, nq)
        for f in range(h.shape[1]):   # field by field: JxW, K entries, x, ...
            scale = np.abs(h[:, f]).max()
            assert np.abs(h[:, f] - d[:, f]).max() <= 1e-13 * max(scale, 1e-300), f"{rule} field {f}"
    mp = params.NAMED[pname]
    model = device.BoussinesqModel.from_problem(ctx, P, mp, device_geometry=True)
    oprm = orc.params_from(mp)
    if feec:
        n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
        rng = np.random.default_rng(3)
        u, T = np.ascontiguousarray(0.1 * rng.uniform(-1, 1, n)), np.ascontiguousarray(2 + 0.3 * rng.uniform(-1, 1, nT))
        ref_vals, ref_rhs = orc.feec_assemble_nse_system(P, oprm, u, T)
        blocks = split_blocks3(P, "nse", ref_vals)
    else:
        u, T = synthetic_fields(P)
        ref_vals, ref_rhs = orc.assemble_nse_system(P, oprm, u, T)
        blocks = split_blocks(P, "nse", ref_vals)
    model.assemble_nse_system(u, T)
    for (bi, bj), rv in blocks.items():
        assert rel_err_max(model.nse_matrix.block(bi, bj).values(), rv) <= TOL
    assert rel_err_max(model.nse_rhs, ref_rhs) <= TOL
    model.close()


def test_error_behaviour_of_the_c_abi(ctx, problem_factory):
    """Every failure is a status code + dcp_last_error(), which the mirrors turn into an exception (the reference's
    convention: std::runtime_error reaching main.cxx:128-156); no call corrupts the model."""
    import ctypes
    from dycore_b200 import device, params
    from oracle import oracle as orc
    P = problem_factory(geometry="shell", refine=1)
    mp = params.NAMED["shell_3d_classic"]
    u, T = synthetic_fields(P)
    # inconsistent description: dofs per cell do not match the element
    d = device.model_desc_from_problem(P)
    d.nse_n_local = 88
    with pytest.raises(device.DcpError, match="classic family expects"):
        device.BoussinesqModel(ctx, d, mp)
    d = device.model_desc_from_problem(P)
    d.geom_qn = None
    with pytest.raises(device.DcpError, match="geom_qn"):
        device.BoussinesqModel(ctx, d, mp)
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    # call order: the rhs needs the temperature matrices (T = M + dt/n K is built from them, :975-978)
    with pytest.raises(device.DcpError):
        model.assemble_temperature_rhs(T, u)
    # bad selectors
    rc = device.lib().dcp_vmult(model._h, 99, 0, 0, ctypes.c_void_p(u.ctypes.data), ctypes.c_void_p(u.ctypes.data), device.HOST)
    assert rc != 0 and b"invalid" in device.lib().dcp_last_error()
    with pytest.raises(device.DcpError):
        device.PreconditionILU(model, device.MAT_NSE, 5)
    assert device.lib().dcp_velocity_extrema(model._h, None, device.HOST, None) != 0
    # the model still works after the failed calls
    model.assemble_nse_system(u, T)
    ref_vals, ref_rhs = orc.assemble_nse_system(P, orc.params_from(mp), u, T)
    assert rel_err_max(model.nse_rhs, ref_rhs) <= TOL
    model.close()


def test_row_classes_partition_the_owned_rows(ctx, problem_factory):
    """dcp_block_vmult_rows / dcp_vmult_rows: INTERIOR + GHOSTED rows together give exactly the plain product on the
    owned rows of a rank's subdomain (one rank of a two-rank partition, no exchange needed for this identity)."""
    import torch
    from dycore_b200 import device, params
    P = problem_factory(geometry="shell", refine=2, n_ranks=2, rank=0)
    mp = params.NAMED["shell_3d_classic"]
    model = device.BoussinesqModel.from_problem(ctx, P, mp)
    n_u, n_uo, n_po = P.scalar("nse.n_u"), P.scalar("nse.n_u_owned"), P.scalar("nse.n_p_owned")
    model.set_owned([n_uo, n_po], P.scalar("temp.n_owned"))
    u, T = synthetic_fields(P)
    model.assemble_nse_system(u, T)
    model.assemble_temperature_matrix()
    model.assemble_temperature_rhs(T, u)
    n = model.n_nse
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    x = torch.from_numpy(np.random.default_rng(4).standard_normal(n)).cuda()
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    model.nse_matrix.vmult(y, x)
    z = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    model.nse_matrix.vmult_rows(z, x, device.ROWS_INTERIOR)
    n_interior = int((~torch.isnan(z)).sum())
    model.nse_matrix.vmult_rows(z, x, device.ROWS_GHOSTED)
    torch.cuda.synchronize()
    owned = np.concatenate([np.arange(n_uo), n_u + np.arange(n_po)])
    yh, zh = y.cpu().numpy(), z.cpu().numpy()
    assert np.array_equal(yh[owned], zh[owned])
    assert 0 < n_interior < len(owned)            # both classes are populated on a partitioned mesh
    # single block and the temperature matrix
    nT = model.n_temp
    xt = torch.from_numpy(np.random.default_rng(5).standard_normal(nT)).cuda()
    yt = torch.zeros(nT, dtype=torch.float64, device="cuda")
    zt = torch.full((nT,), float("nan"), dtype=torch.float64, device="cuda")
    model.temperature_matrix.vmult(yt, xt)
    model.temperature_matrix.vmult_rows(zt, xt, device.ROWS_INTERIOR)
    model.temperature_matrix.vmult_rows(zt, xt, device.ROWS_GHOSTED)
    torch.cuda.synchronize()
    no = P.scalar("temp.n_owned")
    assert np.array_equal(yt.cpu().numpy()[:no], zt.cpu().numpy()[:no])
    ctx.set_stream(None)
    model.close()


RENUMBERED = [dict(renumber="cuthill_mckee", refine=2, **ANNULUS), dict(renumber="random", refine=2, **ANNULUS),
              dict(geometry="shell", refine=1, renumber="random"), dict(geometry="shell", refine=1, renumber="cuthill_mckee")]


@pytest.mark.parametrize("spec", RENUMBERED, ids=["annulus-cmk", "annulus-random", "shell-random", "shell-cmk"])
def test_any_dof_numbering_is_assembled_correctly(ctx, problem_factory, spec):
    """The device path takes the index maps as they come: with Cuthill-McKee (the reference's numbering on its
    Schur-complement path, boussinesq_model.tpp:198-202) or a random permutation the position-table plans reject the
    cells whose layout they cannot verify and the general AffineConstraints scatter takes over -- same matrices."""
    from dycore_b200 import device
    from oracle import oracle as orc
    P = problem_factory(**spec)
    mp = _params(spec)
    u, T = synthetic_fields(P)
    model = device.BoussinesqModel.from_problem(ctx, P, mp)      # default strategy: POSITIONS
    oprm = orc.params_from(mp)
    model.assemble_nse_system(u, T)
    ref_vals, ref_rhs = orc.assemble_nse_system(P, oprm, u, T)
    for (bi, bj), rv in split_blocks(P, "nse", ref_vals).items():
        assert rel_err_max(model.nse_matrix.block(bi, bj).values(), rv) <= TOL, f"nse block {bi}{bj}"
    assert rel_err_max(model.nse_rhs, ref_rhs) <= TOL
    model.assemble_nse_preconditioner()
    for (bi, bj), rv in split_blocks(P, "pre", orc.assemble_nse_preconditioner(P, oprm)).items():
        assert rel_err_max(model.nse_preconditioner_matrix.block(bi, bj).values(), rv) <= TOL, f"pre block {bi}{bj}"
    model.assemble_temperature_matrix()
    model.assemble_temperature_rhs(T, u)
    assert rel_err_max(model.temperature_rhs, orc.assemble_temperature_rhs(P, oprm, T, u)) <= TOL
    model.close()
