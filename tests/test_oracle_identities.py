"""Pins for the CPU oracle, part 1: analytic identities (SURVEY.md 8c "pins the new repo must create", item 1-2).

The reference ships no golden vectors for this path, so the oracle is checked against facts that hold for the
continuous problem / the FE spaces, independent of any implementation:
  * sum_ij M_ij = |Omega_h| (partition of unity), K 1 = 0, M and K symmetric;
  * the NSE matrix is symmetric (the dt*p scaling of the reference makes it so,
    /root/reference/include/core/boussinesq_model.tpp:633-635) and the preconditioner matrix is symmetric;
  * B u = -int psi div u for a velocity field the space represents exactly (affine cells);
  * right-hand sides of manufactured polynomial fields equal closed-form integrals on the unit cube.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc


def _mats(P, prm, u, T):
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    v, rhs = orc.assemble_nse_system(P, prm, u, T)
    rp, col, _, _ = P.csr("nse.full")
    A = sp.csr_matrix((v, col, rp), shape=(n, n))
    pv = orc.assemble_nse_preconditioner(P, prm)
    rp, col, _, _ = P.csr("pre.full")
    Pm = sp.csr_matrix((pv, col, rp), shape=(n, n))
    m, k = orc.assemble_temperature_matrix(P, prm)
    rp, col, _, _ = P.csr("temp.pat")
    M = sp.csr_matrix((m, col, rp), shape=(nT, nT))
    K = sp.csr_matrix((k, col, rp), shape=(nT, nT))
    return A, rhs, Pm, M, K


@pytest.mark.parametrize("spec", [dict(geometry="shell", refine=1, constraints=0),
                                  dict(geometry="cube", refine=1, constraints=0),
                                  dict(geometry="shell", refine=1, constraints=0, temperature_degree=2)],
                         ids=["shell", "cube", "shell-Tq2"])
def test_partition_of_unity_and_symmetry(problem_factory, spec):
    from dycore_b200 import params
    P = problem_factory(**spec)
    mp = params.NAMED["cube_3d" if spec["geometry"] == "cube" else "shell_3d_classic"]
    prm = orc.params_from(mp)
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    A, rhs, Pm, M, K = _mats(P, prm, np.zeros(n), np.zeros(nT))
    nq = P.scalar("q_temp.nq")
    vol = P["geom.qt"].reshape(P.n_cells, -1)[:, :nq].sum()
    assert abs(M.sum() - vol) <= 1e-12 * vol
    assert np.abs(K @ np.ones(nT)).max() <= 1e-12 * np.abs(K).max()
    for X in (M, K, A, Pm):
        assert abs(X - X.T).max() <= 1e-14 * abs(X).max()
    if spec["geometry"] == "cube":
        assert abs(vol - 1.0) <= 1e-13
    else:
        exact = 4.0 / 3.0 * np.pi * (3.0 ** 3 - 1.0)
        assert abs(vol - exact) / exact < 0.15  # r=1 polyhedral / cubic-boundary approximation of the shell


def test_divergence_block_and_manufactured_rhs_on_cube(problem_factory):
    """Affine cells: Q2 reproduces polynomials of degree <= 2 exactly, so integrals are known in closed form."""
    from dycore_b200 import params
    P = problem_factory(geometry="cube", refine=1, constraints=0)
    mp = params.NAMED["cube_3d"]
    prm = orc.params_from(mp)
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    n_u = P.scalar("nse.n_u")
    x = P["nse.dof_xyz"].reshape(n, 3)
    comp = P["nse.dof_comp"]
    # u = (x, 2y, -z^2/2): div u = 1 + 2 - z = 3 - z
    u = np.zeros(n)
    u[comp == 0] = x[comp == 0, 0]
    u[comp == 1] = 2.0 * x[comp == 1, 1]
    u[comp == 2] = -0.5 * x[comp == 2, 2] ** 2
    xt = P["temp.dof_xyz"].reshape(nT, 3)
    T = 3.0 + 0.0 * xt[:, 0]          # T == T_ref  ->  density_scaling == 1
    A, rhs, Pm, M, K = _mats(P, prm, np.ascontiguousarray(u), np.ascontiguousarray(T))
    # pressure rows: (A u)_b = - int psi_b div u ;  sum_b psi_b = 1  ->  sum = - int (3 - z) = -(3 - 1/2)
    Bu = (A @ u)[n_u:]
    assert abs(Bu.sum() + 2.5) <= 1e-12
    # u-rows of A acting on a constant pressure p=1:  -int div phi_i  summed against u gives -int div u
    p1 = np.zeros(n)
    p1[n_u:] = 1.0
    assert abs(u @ (A @ p1) + 2.5) <= 1e-12
    # velocity mass + viscous block: u^T A_uu u = int |u|^2 + dt/Re * 2 int eps(u):eps(u)
    uu = u.copy()
    uu[n_u:] = 0.0
    mass = 1.0 / 3.0 + 4.0 / 3.0 + 0.25 / 5.0
    eps2 = 1.0 + 4.0 + 1.0 / 3.0      # eps = diag(1, 2, -z)
    want = mass + mp.time_step * mp.inv_re * 2.0 * eps2
    assert abs(uu @ (A @ uu) - want) <= 1e-12 * want
    # preconditioner: u^T P u = int |u|^2 + dt/Re int grad u : grad u  (same here: grad u is diagonal)
    wantp = mass + mp.time_step * mp.inv_re * eps2
    assert abs(uu @ (Pm @ uu) - wantp) <= 1e-12 * wantp
    # rhs tested against v = (1,0,0), (0,1,0), (0,0,1) (sum of the nodal basis of one component):
    #   int ( u_c + dt*rho*g_c - dt (u.grad)u_c - dt*2 (Omega x u)_c ),  Omega = (0,0,L*omega/U), g = -(L/U^2) g e_z
    dt, om = mp.time_step, mp.cor_scale * mp.omega
    # (u.grad)u = (x, 4y, z^3/2);  Omega x u = (-om*u_y, om*u_x, 0) = (-2 om y, om x, 0)
    want_r = [0.5 - dt * 0.5 - dt * 2.0 * (-2.0 * om * 0.5),
              1.0 - dt * 2.0 - dt * 2.0 * (om * 0.5),
              -1.0 / 6.0 + dt * (-mp.g_scale * mp.gravity_constant) - dt * (0.5 / 4.0)]
    for c in range(3):
        got = rhs[:n_u][comp[:n_u] == c].sum()
        assert abs(got - want_r[c]) <= 1e-12, (c, got, want_r[c])


def test_temperature_rhs_manufactured_on_cube(problem_factory):
    from dycore_b200 import params
    P = problem_factory(geometry="cube", refine=1, constraints=0)
    mp = params.NAMED["cube_3d"]
    prm = orc.params_from(mp)
    n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    x = P["nse.dof_xyz"].reshape(n, 3)
    comp = P["nse.dof_comp"]
    u = np.zeros(n)
    u[comp == 0] = x[comp == 0, 1]            # u = (y, 0, x z)
    u[comp == 2] = x[comp == 2, 0] * x[comp == 2, 2]
    xt = P["temp.dof_xyz"].reshape(nT, 3)
    T = 1.0 + 2.0 * xt[:, 0] - xt[:, 2]       # grad T = (2, 0, -1)  (Q1 exact)
    r = orc.assemble_temperature_rhs(P, prm, np.ascontiguousarray(T), np.ascontiguousarray(u))
    tau = mp.time_step / mp.NSE_solver_interval
    # sum_i r_i = int ( T - tau u.grad T ) = int (1 + 2x - z) - tau int (2y - x z) = 1.5 - tau (1 - 1/4)
    assert abs(r.sum() - (1.5 - tau * 0.75)) <= 1e-13


def test_velocity_extrema_restatement_on_known_fields(problem_factory):
    """A constant velocity (1,2,2): max |u| = 3 and CFL = 3 / (smallest cell diameter); the cube's cells are
    identical boxes whose diameter is the box diagonal."""
    from oracle import oracle as orc
    P = problem_factory(geometry="cube", refine=2)
    n = P.scalar("nse.n_dofs")
    comp = P["nse.dof_comp"]
    x = np.zeros(n)
    for c, v in enumerate((1.0, 2.0, 2.0)):
        x[comp == c] = v
    vmax, cfl = orc.velocity_extrema(P, x)
    d = orc.cell_diameters(P)
    X = P["cell_vertices"].reshape(P.n_cells, 8, 3)
    assert np.allclose(d, np.linalg.norm(X[:, 7] - X[:, 0], axis=1))
    assert abs(vmax - 3.0) < 1e-14 and abs(cfl - 3.0 / d.min()) < 1e-13
    # distribute leaves unconstrained entries alone and is idempotent
    y = orc.constraints_distribute(P, "nse.cs", x)
    free = np.ones(n, bool)
    free[P["nse.cs.line_dof"]] = False
    assert np.array_equal(y[free], x[free])
    assert np.array_equal(orc.constraints_distribute(P, "nse.cs", y), y)
