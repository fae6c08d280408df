"""Size-independent properties of the CUDA path at sizes the CPU oracle cannot reach in seconds.

Default refine 4 (0.67 M DoFs); DCP_PROPERTY_REFINE=5 or 6 runs the bench sizes (5.2 M / 41 M DoFs).  All
products stay on the device; only scalars and O(n) vectors come back."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REFINE = int(os.environ.get("DCP_PROPERTY_REFINE", "4"))
TOL = 1e-12


@pytest.fixture(scope="module")
def setup():
    import torch
    import dycore_b200  # noqa: F401
    from dycore_b200 import device, harness, params
    from util import synthetic_fields
    P = harness.Problem(geometry="shell", refine=REFINE)
    ctx = device.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)   # torch's generators / fills and the library share one stream
    mp = params.NAMED["shell_3d_classic"]
    model = device.BoussinesqModel.from_problem(ctx, P, mp, device_geometry=True)
    u, T = synthetic_fields(P)
    yield dict(P=P, ctx=ctx, model=model, u=torch.from_numpy(u).cuda(), T=torch.from_numpy(T).cuda(), torch=torch, device=device)
    model.close()
    ctx.close()


def _rand(torch, n, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.randn(n, dtype=torch.float64, device="cuda", generator=g)


def test_strategies_agree_through_checksums(setup):
    """SEARCH (general AffineConstraints scatter, parity-checked against the oracle at small sizes), POSITIONS (DMMA
    kernel with position tables, reductions) and STAGED (write-once: node-major staging + TMA gather, the default on the
    shell; the preconditioner comes out of the system pass) must assemble the same matrices: compare A x for random x,
    and the rhs."""
    s, torch, device = setup, setup["torch"], setup["device"]
    m = s["model"]
    n = m.n_nse
    x = _rand(torch, n, 1)
    out = {}
    assert m.strategy == device.STRATEGY_STAGED, "the shell qualifies for the write-once path"
    for name, strat in (("search", device.STRATEGY_SEARCH), ("positions", device.STRATEGY_POSITIONS), ("staged", device.STRATEGY_STAGED)):
        m.set_strategy(strat)
        m.assemble_nse_system(s["u"], s["T"])
        m.assemble_nse_preconditioner()
        y, z = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
        m.nse_matrix.vmult(y, x)
        m.nse_preconditioner_matrix.vmult(z, x)
        torch.cuda.synchronize()
        out[name] = (y.clone(), z.clone(), torch.from_numpy(m.nse_rhs))
    for other in ("positions", "staged"):
        for a, b, what in zip(out["search"], out[other], ("nse_matrix x", "preconditioner x", "nse_rhs")):
            assert float((a - b).abs().max() / b.abs().max()) <= TOL, (other, what)


def test_nse_matrix_is_symmetric(setup):
    """<A x, y> = <x, A y>: the dt-scaled pressure coupling makes the Stokes matrix symmetric
    (boussinesq_model.tpp:626-637), constraints are resolved symmetrically."""
    s, torch = setup, setup["torch"]
    m = s["model"]
    m.assemble_nse_system(s["u"], s["T"])
    n = m.n_nse
    x, y = _rand(torch, n, 2), _rand(torch, n, 3)
    ax, ay = torch.zeros_like(x), torch.zeros_like(x)
    m.nse_matrix.vmult(ax, x)
    m.nse_matrix.vmult(ay, y)
    torch.cuda.synchronize()
    lhs, rhs = float(torch.dot(ax, y)), float(torch.dot(x, ay))
    scale = float(torch.dot(ax.abs(), y.abs()))
    assert abs(lhs - rhs) <= TOL * scale


def test_block_products_add_up(setup):
    """vmult over the whole block matrix equals the sum of its block products (block_schur_preconditioner.hpp:55)."""
    s, torch = setup, setup["torch"]
    m, P = s["model"], s["P"]
    n, n_u = m.n_nse, P.scalar("nse.n_u")
    x = _rand(torch, n, 4)
    y = torch.zeros_like(x)
    m.nse_matrix.vmult(y, x)
    z = torch.zeros_like(x)
    m.nse_matrix.block(0, 0).vmult(z[:n_u], x[:n_u])
    m.nse_matrix.block(0, 1).vmult_add(z[:n_u], x[n_u:])
    m.nse_matrix.block(1, 0).vmult(z[n_u:], x[:n_u])
    torch.cuda.synchronize()
    assert float((y - z).abs().max() / y.abs().max()) <= TOL


def test_temperature_matrices_properties(setup):
    """Mass matrix: x'Mx > 0 and symmetric; stiffness annihilates constants on rows away from constrained dofs;
    temperature_matrix = M + dt/n K (boussinesq_model.tpp:975-978); sum of JxW = shell volume to mapping accuracy."""
    s, torch = setup, setup["torch"]
    m, P = s["model"], s["P"]
    from dycore_b200 import params
    mp = params.NAMED["shell_3d_classic"]
    m.assemble_temperature_matrix()
    m.assemble_temperature_rhs(s["T"], s["u"])
    nT = m.n_temp
    x, y = _rand(torch, nT, 5), _rand(torch, nT, 6)
    mx, my, kx, tx = (torch.zeros_like(x) for _ in range(4))
    m.temperature_mass_matrix.vmult(mx, x)
    m.temperature_mass_matrix.vmult(my, y)
    m.temperature_stiffness_matrix.vmult(kx, x)
    m.temperature_matrix.vmult(tx, x)
    torch.cuda.synchronize()
    assert float(torch.dot(mx, x)) > 0
    assert abs(float(torch.dot(mx, y) - torch.dot(x, my))) <= TOL * float(torch.dot(mx.abs(), y.abs()))
    factor = mp.time_step / mp.NSE_solver_interval
    assert float((tx - (mx + factor * kx)).abs().max() / tx.abs().max()) <= TOL
    ones = torch.ones(nT, dtype=torch.float64, device="cuda")
    k1 = torch.zeros_like(ones)
    m.temperature_stiffness_matrix.vmult(k1, ones)
    # rows that couple to a constrained dof lose entries; all others must sum to zero
    lod = torch.from_numpy(P["temp.cs.line_of_dof"]).cuda()
    touched = torch.zeros(nT, dtype=torch.float64, device="cuda")
    touched[lod >= 0] = 1.0
    l2g = torch.from_numpy(P["temp.l2g"].reshape(P.n_cells, -1)).cuda().long()
    cell_flag = touched[l2g].amax(dim=1)
    near = torch.zeros(nT, dtype=torch.float64, device="cuda")
    near.index_reduce_(0, l2g.reshape(-1), cell_flag.repeat_interleave(l2g.shape[1]), "amax", include_self=True)
    free = near == 0
    diag_scale = float(kx.abs().max())
    assert float(k1[free].abs().max()) <= 1e-10 * diag_scale
    # volume: 1' M 1 over unconstrained-only rows is not the volume; use the mapping weights instead
    vol = 4.0 / 3.0 * np.pi * (3.0 ** 3 - 1.0 ** 3)
    from dycore_b200 import device
    g = device.geometry_create(s["ctx"], P, "qn")
    nq = 27
    jxw = torch.from_numpy(s["ctx"].download_f64(g, P.n_cells * 13 * nq).reshape(P.n_cells, 13, nq)[:, 0]).sum()
    s["ctx"].free(g)
    h = 2.0 ** -REFINE
    assert abs(float(jxw) - vol) <= 2.0 * h * h * vol   # trilinear interior cells: O(h^2) volume defect
