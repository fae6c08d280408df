"""Krylov solvers and operator compositions that CALL the hot path -- host-side mirror of

  * deal.II SolverCG / SolverGMRES / SolverFGMRES as the reference uses them
    (/root/reference/include/core/boussinesq_model.tpp:1165-1199 FGMRES(30) on nse_matrix, :1426-1440 CG on
    temperature_matrix; include/linear_algebra/block_schur_preconditioner.hpp:46-51 inner GMRES),
  * LinearAlgebra::SchurComplement (schur_complement.hpp:143-150) and
    LinearAlgebra::BlockSchurPreconditioner::vmult (block_schur_preconditioner.hpp:41-70),
  * Standard::BoussinesqModel::solve_NSE_block_preconditioned (:1131-1246, incl. the double pressure scaling,
    quirk Q7) and solve_temperature (:1417-1447).

The algorithms are written once against a tiny vector backend, so the same code drives
  - `DeviceBackend`: vectors are torch CUDA tensors (device memory only), every operation is a libdcp kernel
    (dcp_vec_dot / axpy / sadd / scale / copy, dcp_vmult, dcp_jacobi_vmult) -- vectors never leave HBM, only the
    scalars of the dot products travel, like the reference's MPI_Allreduce;
  - `NumpyBackend` (tests only): numpy vectors and scipy matrices built from the CPU oracle.
deal.II's solver internals are un-vendored; what is restated (from memory, medium confidence):
CG stops on ||r||_2 <= tol (absolute); GMRES = restarted, modified Gram-Schmidt + Givens, default 30 temporary
vectors (restart length 28), left preconditioning, stops on the Givens residual estimate; FGMRES = flexible
right-preconditioned GMRES with the given basis size, same stopping rule.  `SolverControl` semantics: the check
happens before the first step (step 0) and after every step; `last_step` is returned.
"""
import ctypes
import math

import numpy as np


class NoConvergence(RuntimeError):
    def __init__(self, last_step, last_residual):
        super().__init__(f"no convergence after {last_step} steps, residual {last_residual:.3e}")
        self.last_step, self.last_residual = last_step, last_residual


# ---- backends -------------------------------------------------------------------------------------------
class NumpyBackend:
    """CPU vectors (tests only)."""

    def zeros(self, n):
        return np.zeros(n)

    def copy(self, x):
        return x.copy()

    def from_numpy(self, a):
        return np.array(a, dtype=np.float64)

    def assign(self, dst, src):
        dst[...] = src

    def dot(self, x, y):
        return float(np.dot(x, y))

    def axpy(self, a, x, y):
        y += a * x

    def sadd(self, s, a, x, y):
        y *= s
        y += a * x

    def scale(self, a, y):
        y *= a

    def zero(self, y):
        y[...] = 0.0                               # assignment (deal.II `dst = 0`), not a product: NaN/Inf do not survive

    def add_scalar(self, a, y):
        y += a

    def to_numpy(self, x):
        return np.array(x)


class DeviceBackend:
    """Vectors in HBM (torch CUDA float64 tensors as plain device memory); all math through libdcp."""

    def __init__(self, ctx):
        import torch
        from . import device
        self.torch, self.dv, self.ctx = torch, device, ctx
        self._res = ctypes.c_double()
        self.resident_cg = True      # SolverCG inside the library where the operands allow (dcp_cg_solve)
        self.supports_mgs = True     # Arnoldi orthogonalisation with one synchronisation (dcp_vec_mgs)
        # the library's own stream is non-blocking: run it on torch's current stream so that tensor creation
        # (torch.zeros, .cuda()) and the library's kernels on the same memory are ordered
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    def _p(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def zeros(self, n):
        return self.torch.zeros(n, dtype=self.torch.float64, device="cuda")

    def copy(self, x):
        y = self.torch.empty_like(x)
        self.assign(y, x)
        return y

    def from_numpy(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()

    def assign(self, dst, src):
        self.dv.check(self.dv.lib().dcp_vec_copy(self.ctx._h, dst.numel(), self._p(src), self._p(dst)), "dcp_vec_copy")

    def dot(self, x, y):
        self.dv.check(self.dv.lib().dcp_vec_dot(self.ctx._h, x.numel(), self._p(x), self._p(y), ctypes.byref(self._res)),
                      "dcp_vec_dot")
        return self._res.value

    def mgs(self, w, basis):
        """Modified Gram-Schmidt of w against `basis` on the device (dcp_vec_mgs): returns [w.v_0, ..., w.v_{k-1}, w.w]
        with one host synchronisation; w is updated in place."""
        k = len(basis)
        ptrs = (ctypes.c_void_p * max(k, 1))(*[b.data_ptr() for b in basis])
        h = (ctypes.c_double * (k + 1))()
        self.dv.check(self.dv.lib().dcp_vec_mgs(self.ctx._h, w.numel(), k, ptrs, self._p(w), h), "dcp_vec_mgs")
        return list(h)

    def axpy(self, a, x, y):
        self.dv.check(self.dv.lib().dcp_vec_axpy(self.ctx._h, x.numel(), float(a), self._p(x), self._p(y)), "dcp_vec_axpy")

    def sadd(self, s, a, x, y):
        self.dv.check(self.dv.lib().dcp_vec_sadd(self.ctx._h, x.numel(), float(s), float(a), self._p(x), self._p(y)),
                      "dcp_vec_sadd")

    def scale(self, a, y):
        self.dv.check(self.dv.lib().dcp_vec_scale(self.ctx._h, y.numel(), float(a), self._p(y)), "dcp_vec_scale")

    def zero(self, y):
        self.dv.check(self.dv.lib().dcp_vec_fill(self.ctx._h, y.numel(), 0.0, self._p(y)), "dcp_vec_fill")

    def add_scalar(self, a, y):
        self.dv.check(self.dv.lib().dcp_vec_shift(self.ctx._h, y.numel(), float(a), self._p(y)), "dcp_vec_shift")

    def to_numpy(self, x):
        self.ctx.synchronize()
        return x.cpu().numpy()


# ---- operators --------------------------------------------------------------------------------------------
class Identity:
    def vmult(self, dst, src, B):
        B.assign(dst, src)


class Wrap:
    """Adapts any object with vmult(dst, src) (device SparseMatrix / PreconditionJacobi, or a callable)."""

    def __init__(self, op):
        self.op = op
        self._tmp = None

    def vmult(self, dst, src, B):
        if callable(self.op):
            self.op(dst, src)
        else:
            self.op.vmult(dst, src)

    def vmult_add(self, dst, src, B):
        if callable(self.op):
            if self._tmp is None or len(self._tmp) != len(dst):
                self._tmp = B.zeros(len(dst))
            self.op(self._tmp, src)
            B.axpy(1.0, self._tmp, dst)
        else:
            self.op.vmult_add(dst, src)


# ---- deal.II solvers -----------------------------------------------------------------------------------------
def _device_cg_operands(B, A, P):
    """(device SparseMatrix, device preconditioner or None) when the whole solve can run inside libdcp (dcp_cg_solve:
    a square block of a device matrix, preconditioned by the identity, a Jacobi sweep or ILU(0)); else None."""
    if not isinstance(B, DeviceBackend) or not B.resident_cg:
        return None
    dv = B.dv
    mat = A.op if isinstance(A, Wrap) else A
    if not isinstance(mat, dv.SparseMatrix) or mat.n_rows != mat.n_cols:
        return None
    pre = P.op if isinstance(P, Wrap) else P
    if isinstance(pre, Identity):
        pre = None
    elif not isinstance(pre, (dv.PreconditionJacobi, dv.PreconditionILU)):
        return None
    return mat, pre


def solver_cg(B, A, x, b, P, tol, max_steps):
    """SolverCG<Vector>::solve(A, x, b, P) with SolverControl(max_steps, tol).  Returns last_step.  On the device
    backend, with a plain matrix block and an identity / Jacobi / ILU(0) preconditioner, the whole loop runs inside the
    library (no host synchronisation per iteration); every other combination uses the loop below."""
    ops = _device_cg_operands(B, A, P)
    if ops is not None:
        ok, step, res = B.dv.cg_solve(ops[0], x, b, tol, max_steps, ops[1])
        if not ok:
            raise NoConvergence(step, res)
        return step
    n = len(b) if hasattr(b, "__len__") else b.numel()
    r, z, p, Ap = B.zeros(n), B.zeros(n), B.zeros(n), B.zeros(n)
    A.vmult(r, x, B)
    B.sadd(-1.0, 1.0, b, r)                      # r = b - A x
    res = math.sqrt(B.dot(r, r))
    if res <= tol:
        return 0
    P.vmult(z, r, B)
    B.assign(p, z)
    rz = B.dot(r, z)
    for it in range(1, max_steps + 1):
        A.vmult(Ap, p, B)
        alpha = rz / B.dot(p, Ap)
        B.axpy(alpha, p, x)
        B.axpy(-alpha, Ap, r)
        res = math.sqrt(B.dot(r, r))
        if res <= tol:
            return it
        P.vmult(z, r, B)
        rz_new = B.dot(r, z)
        B.sadd(rz_new / rz, 1.0, z, p)            # p = z + beta p
        rz = rz_new
    raise NoConvergence(max_steps, res)


def _givens(h, cs, sn, gamma, col):
    """Apply the stored rotations to column `col` of the Hessenberg matrix and create the new one."""
    for i in range(col):
        t = cs[i] * h[i] + sn[i] * h[i + 1]
        h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1]
        h[i] = t
    s = math.hypot(h[col], h[col + 1])
    cs[col], sn[col] = h[col] / s, h[col + 1] / s
    h[col] = s
    h[col + 1] = 0.0
    gamma[col + 1] = -sn[col] * gamma[col]
    gamma[col] = cs[col] * gamma[col]


def _solve_upper(H, gamma, k):
    y = np.zeros(k)
    for i in range(k - 1, -1, -1):
        y[i] = (gamma[i] - sum(H[i][j] * y[j] for j in range(i + 1, k))) / H[i][i]
    return y


def solver_gmres(B, A, x, b, P, tol, max_steps, restart=28, flexible=False):
    """SolverGMRES (left preconditioned, restart = max_n_tmp_vectors - 2) or, with flexible=True,
    SolverFGMRES(max_basis_size = restart) (right preconditioned, stores the preconditioned basis)."""
    n = b.numel() if hasattr(b, "numel") else len(b)
    V = [B.zeros(n) for _ in range(restart + 1)]
    Z = [B.zeros(n) for _ in range(restart)] if flexible else None
    w, t = B.zeros(n), B.zeros(n)
    steps = 0
    while True:
        A.vmult(t, x, B)
        B.sadd(-1.0, 1.0, b, t)                  # t = b - A x
        if flexible:
            B.assign(V[0], t)
        else:
            P.vmult(V[0], t, B)                  # left preconditioning: residual of P^-1 A x = P^-1 b
        beta = math.sqrt(B.dot(V[0], V[0]))
        if beta <= tol:
            return steps
        if steps >= max_steps:
            raise NoConvergence(steps, beta)
        B.scale(1.0 / beta, V[0])
        H = [[0.0] * restart for _ in range(restart + 1)]
        cs, sn = [0.0] * restart, [0.0] * restart
        gamma = [0.0] * (restart + 1)
        gamma[0] = beta
        k_done = 0
        converged = False
        for k in range(restart):
            if flexible:
                P.vmult(Z[k], V[k], B)
                A.vmult(w, Z[k], B)
            else:
                A.vmult(t, V[k], B)
                P.vmult(w, t, B)
            if getattr(B, "supports_mgs", False) and k + 1 <= 256:   # the whole orthogonalisation with one host synchronisation
                h = B.mgs(w, V[:k + 1])
                h[k + 1] = math.sqrt(h[k + 1])
            else:
                h = [0.0] * (k + 2)
                for i in range(k + 1):               # modified Gram-Schmidt
                    h[i] = B.dot(w, V[i])
                    B.axpy(-h[i], V[i], w)
                h[k + 1] = math.sqrt(B.dot(w, w))
            if h[k + 1] != 0.0:
                B.assign(V[k + 1], w)
                B.scale(1.0 / h[k + 1], V[k + 1])
            _givens(h, cs, sn, gamma, k)
            for i in range(k + 2):
                H[i][k] = h[i]
            steps += 1
            k_done = k + 1
            res = abs(gamma[k + 1])
            if res <= tol:
                converged = True
                break
            if steps >= max_steps:
                break
        y = _solve_upper(H, gamma, k_done)
        basis = Z if flexible else V
        for i in range(k_done):
            B.axpy(float(y[i]), basis[i], x)
        if converged:
            return steps
        if steps >= max_steps:
            raise NoConvergence(steps, abs(gamma[k_done]))


# ---- LinearAlgebra::* compositions of the classic block solve ---------------------------------------------------
class SchurComplement:
    """B * inverse * B^T on the pressure space (schur_complement.hpp:143-150)."""

    def __init__(self, block_01, block_10, inverse, n_u, B):
        self.b01, self.b10, self.inv = block_01, block_10, inverse
        self.tmp1, self.tmp2 = B.zeros(n_u), B.zeros(n_u)

    def vmult(self, dst, src, B):
        self.b01.vmult(self.tmp1, src, B)
        self.inv.vmult(self.tmp2, self.tmp1, B)
        self.b10.vmult(dst, self.tmp2, B)


class InverseMatrix:
    """LinearAlgebra::InverseMatrix::vmult (inverse_matrix.hpp:90-121): CG to 1e-6 |src| with the given
    preconditioner, at most max(n, 1000) steps, dst starts from zero.  A solver exception is caught there and only
    `Assert`ed (:116-119), i.e. a release build continues with the last iterate: `strict=False` mirrors that,
    `strict=True` (default) re-raises like a debug build would abort."""

    def __init__(self, matrix, preconditioner, strict=True):
        self.matrix, self.preconditioner, self.strict = matrix, preconditioner, strict
        self.iterations = []

    def _solve(self, dst, src, B, max_steps):
        tol = 1e-6 * math.sqrt(B.dot(src, src))
        B.zero(dst)                                                   # dst = 0 (:101)
        try:
            self.iterations.append(solver_cg(B, self.matrix, dst, src, self.preconditioner, tol, max_steps))
        except NoConvergence as e:
            if self.strict:
                raise
            self.iterations.append(e.last_step)

    def vmult(self, dst, src, B):
        self._solve(dst, src, B, max(len(src), 1000))


class ApproximateInverseMatrix(InverseMatrix):
    """approximate_inverse.hpp:97-128: the same CG with SolverControl(n_iter, 1e-6 |src|); the reference passes
    numbers::invalid_unsigned_int for n_iter (boussinesq_model.tpp:1359-1372), i.e. no effective step limit."""

    def __init__(self, matrix, preconditioner, n_iter=4294967295, strict=True):
        super().__init__(matrix, preconditioner, strict)
        self.n_iter = n_iter

    def vmult(self, dst, src, B):
        self._solve(dst, src, B, self.n_iter)


class ApproximateSchurComplement:
    """approximate_schur_complement.hpp:129-142: block(1,0) * P(block(0,0)) * block(0,1) with P = PreconditionILU."""

    def __init__(self, block_01, block_10, preconditioner, n_u, B):
        self.b01, self.b10, self.prec = block_01, block_10, preconditioner
        self.tmp1, self.tmp2 = B.zeros(n_u), B.zeros(n_u)

    def vmult(self, dst, src, B):
        self.b01.vmult(self.tmp1, src, B)
        self.prec.vmult(self.tmp2, self.tmp1, B)
        self.b10.vmult(dst, self.tmp2, B)


class BlockSchurPreconditioner:
    """block_schur_preconditioner.hpp:17-86: note that the Schur complement is built with the A-preconditioner as
    its "inverse" (:32-35) and mp_preconditioner is never applied.  do_solve_A = true (:59-67, the fall-back of
    boussinesq_model.tpp:1203-1232) replaces the single A-preconditioner sweep by LA::SolverGMRES on block(0,0) to
    1e-2 |utmp| -- Trilinos AztecOO GMRES as deal.II's wrapper configures it [from memory: restart 30, right
    preconditioning, absolute residual test], i.e. the flexible solver below with a fixed preconditioner."""

    def __init__(self, blocks, a_preconditioner, n_u, n_p, B, do_solve_A=False):
        self.blocks, self.a_prec, self.n_u, self.n_p = blocks, a_preconditioner, n_u, n_p
        self.do_solve_A = do_solve_A
        self.schur = SchurComplement(blocks[(0, 1)], blocks[(1, 0)], a_preconditioner, n_u, B)
        self.utmp = B.zeros(n_u)
        self.inner_iterations = []
        self.a_iterations = []

    def vmult(self, dst, src, B):
        n_u = self.n_u
        du, dp = dst[:n_u], dst[n_u:]
        su, sp = src[:n_u], src[n_u:]
        B.zero(dst)                            # deal.II hands a fresh (zero) dst to the preconditioner
        tol = 1e-6 * math.sqrt(B.dot(sp, sp))
        its = solver_gmres(B, self.schur, dp, sp, Identity(), tol, 5000)     # :46-51
        self.inner_iterations.append(its)
        B.scale(-1.0, dp)
        self.blocks[(0, 1)].vmult(self.utmp, dp, B)                          # :55-57
        B.sadd(-1.0, 1.0, su, self.utmp)
        if self.do_solve_A:                                                  # :59-67
            tol_a = 1e-2 * math.sqrt(B.dot(self.utmp, self.utmp))
            self.a_iterations.append(solver_gmres(B, self.blocks[(0, 0)], du, self.utmp, self.a_prec, tol_a, 5000,
                                                  restart=30, flexible=True))
        else:
            self.a_prec.vmult(du, self.utmp, B)                              # :69


class BlockOperator:
    """LA::BlockSparseMatrix::vmult from its blocks (used with the numpy backend; the device has dcp_block_vmult)."""

    def __init__(self, blocks, n_u, B):
        self.blocks, self.n_u = blocks, n_u
        self.tmp = None

    def vmult(self, dst, src, B):
        n_u = self.n_u
        self.blocks[(0, 0)].vmult(dst[:n_u], src[:n_u], B)
        if self.tmp is None:
            self.tmp = B.zeros(n_u)
        self.blocks[(0, 1)].vmult(self.tmp, src[n_u:], B)
        B.axpy(1.0, self.tmp, dst[:n_u])
        self.blocks[(1, 0)].vmult(dst[n_u:], src[:n_u], B)


def distribute(B, cs_lines, x_np):
    """AffineConstraints::distribute on a host copy: x[line] = sum w x[master] + inhom."""
    line_dof, line_ptr, entry_dof, entry_w, inhom = cs_lines
    for l, g in enumerate(line_dof):
        sl = slice(line_ptr[l], line_ptr[l + 1])
        x_np[g] = float(np.dot(entry_w[sl], x_np[entry_dof[sl]])) + inhom[l]
    return x_np


def solve_nse_block_preconditioned(B, nse_matrix, blocks, a_preconditioner, nse_rhs, nse_solution, n_u, n_p, dt,
                                   constrained_pressure=None, max_steps=40):
    """boussinesq_model.tpp:1131-1246 up to (not including) constraints.distribute.  `constrained_pressure`: indices
    (inside the pressure block) of constrained pressure dofs, zeroed like :1160-1162 (none in the named configs).
    `max_steps` = 40 in the reference (:1166); when FGMRES(30) does not converge in that many steps the reference
    re-solves from the last iterate with do_solve_A = true, FGMRES(50) and nse_matrix.m() steps and reports the summed
    count (:1203-1232).  Returns (solution vector with the SCALED pressure, outer iterations, inner iteration list)."""
    x = B.copy(nse_solution)
    B.scale(dt, x[n_u:])                                  # :1151
    if constrained_pressure is not None and len(constrained_pressure):   # :1160-1162
        x[n_u:][constrained_pressure] = 0.0
    tol = 1e-8 * math.sqrt(B.dot(nse_rhs, nse_rhs))       # :1165
    B.scale(dt, x[n_u:])                                  # :1177 (quirk Q7: scaled twice)
    P = BlockSchurPreconditioner(blocks, a_preconditioner, n_u, n_p, B)
    try:
        its = solver_gmres(B, nse_matrix, x, nse_rhs, P, tol, max_steps, restart=30, flexible=True)   # :1191-1199
        inner = P.inner_iterations
    except NoConvergence as e:                            # :1203-1232
        P2 = BlockSchurPreconditioner(blocks, a_preconditioner, n_u, n_p, B, do_solve_A=True)
        its2 = solver_gmres(B, nse_matrix, x, nse_rhs, P2, tol, n_u + n_p, restart=50, flexible=True)
        its = e.last_step + its2
        inner = P.inner_iterations + P2.inner_iterations
    return x, its, inner


def solve_nse_schur_complement(B, blocks, ilu_00, nse_rhs, nse_solution, n_u, n_p, dt, distribute_fn,
                               constrained_pressure=None):
    """solve_NSE_Schur_complement (boussinesq_model.tpp:1248-1414), the path of data/aqua_planet_test_2d.prm.
    blocks: {(i,j): operator}; ilu_00: PreconditionILU of block(0,0) (inner_schur_preconditioner and the one inside
    ApproximateSchurComplement are both ILU(0) of the same block); distribute_fn(x) applies nse_constraints.distribute
    in place.  Returns (solution with the pressure scaled back, GMRES iterations, [inner CG iteration lists])."""
    x = B.copy(nse_solution)
    xu, xp = x[:n_u], x[n_u:]
    B.scale(dt, xp)                                                     # :1283
    if constrained_pressure is not None and len(constrained_pressure):  # :1291-1293
        xp[constrained_pressure] = 0.0
    block_inverse = InverseMatrix(blocks[(0, 0)], ilu_00)               # :1271-1275
    schur = SchurComplement(blocks[(0, 1)], blocks[(1, 0)], block_inverse, n_u, B)   # :1299-1305
    tmp = B.zeros(n_u)
    schur_rhs = B.zeros(n_p)
    block_inverse.vmult(tmp, nse_rhs[:n_u], B)                          # :1315
    blocks[(1, 0)].vmult(schur_rhs, tmp, B)                             # :1316
    B.axpy(-1.0, nse_rhs[n_u:], schur_rhs)                              # :1317
    approx_schur = ApproximateSchurComplement(blocks[(0, 1)], blocks[(1, 0)], ilu_00, n_u, B)   # :1342-1347
    prec = ApproximateInverseMatrix(approx_schur, Identity())           # :1359-1372
    tol = 1e-6 * math.sqrt(B.dot(schur_rhs, schur_rhs))                 # :1332-1333
    its = solver_gmres(B, schur, xp, schur_rhs, prec, tol, n_u + n_p)   # :1374-1377
    distribute_fn(x)                                                    # :1384
    blocks[(0, 1)].vmult(tmp, xp, B)                                    # :1396-1398
    B.sadd(-1.0, 1.0, nse_rhs[:n_u], tmp)
    block_inverse.vmult(xu, tmp, B)                                     # :1401
    distribute_fn(x)                                                    # :1406
    B.scale(1.0 / dt, xp)                                               # :1412
    return x, its, (block_inverse.iterations, prec.iterations)


# ---- FEEC block solve (boussineq_model_FEEC.tpp:1268-1477) -------------------------------------------------------
class MeanValue:
    """VectorTools::compute_mean_value of the DG0 pressure: sum_K p_K |K| / sum_K |K| with |K| from the quadrature the
    caller names (QGauss(1) in nested_schur_complement.hpp:180-182, 317-319; QGauss(2) in
    preconditioner_block_identity.hpp:38-40 and boussineq_model_FEEC.tpp:1381-1383).  `weights` = |K| / sum |K| in
    pressure-dof order (host array); one dot product and one shift on the backend."""

    def __init__(self, weights, B):
        self.w = B.from_numpy(weights)

    def value(self, p, B):
        return B.dot(self.w, p)

    def subtract(self, p, B):
        mean = self.value(p, B)
        B.add_scalar(-mean, p)
        return mean


class ShiftedSchurComplement:
    """shifted_schur_complement.hpp:155-171 (the `false` branch is dead): dst = B11 src - B10 inv(B01 src)."""

    def __init__(self, blocks, relevant_inverse, n_w, B):
        self.blocks, self.inv = blocks, relevant_inverse
        self.tmp1, self.tmp2 = B.zeros(n_w), B.zeros(n_w)

    def vmult(self, dst, src, B):
        self.blocks[(1, 1)].vmult(dst, src, B)
        self.blocks[(0, 1)].vmult(self.tmp1, src, B)
        self.inv.vmult(self.tmp2, self.tmp1, B)
        B.scale(-1.0, self.tmp2)
        self.blocks[(1, 0)].vmult_add(dst, self.tmp2, B)


class ApproxShiftedSchurComplementInverse:
    """shifted_schur_complement.hpp:271-298: GMRES(<= 30 steps, 1e-6 |src|) on the shifted Schur complement,
    preconditioned by the u-mass inverse (a Jacobi sweep in the reference's driver); NoConvergence is swallowed."""

    def __init__(self, blocks, mass_w_inverse, mass_u_inverse, n_w, B):
        self.op = ShiftedSchurComplement(blocks, mass_w_inverse, n_w, B)
        self.mu_inv = mass_u_inverse
        self.iterations = []

    def vmult(self, dst, src, B):
        tol = 1e-6 * math.sqrt(B.dot(src, src))
        try:
            self.iterations.append(solver_gmres(B, self.op, dst, src, self.mu_inv, tol, 30))
        except NoConvergence as e:
            self.iterations.append(e.last_step)


class SchurComplementLowerBlock:
    """schur_complement.hpp:255-276: block(2,1) * inverse * block(1,2); do_full_solve selects the strong inverse."""

    def __init__(self, blocks, relevant_inverse, relevant_approx_inverse, n_u, B, do_full_solve=False):
        self.blocks = blocks
        self.inv = relevant_inverse if do_full_solve else relevant_approx_inverse
        self.tmp1, self.tmp2 = B.zeros(n_u), B.zeros(n_u)

    def vmult(self, dst, src, B):
        self.blocks[(1, 2)].vmult(self.tmp1, src, B)
        self.inv.vmult(self.tmp2, self.tmp1, B)
        self.blocks[(2, 1)].vmult(dst, self.tmp2, B)


class NestedSchurComplement(SchurComplementLowerBlock):
    """nested_schur_complement.hpp:163-183: the same product followed by the zero-mean correction (QGauss(1))."""

    def __init__(self, blocks, relevant_inverse, n_u, B, mean):
        super().__init__(blocks, relevant_inverse, relevant_inverse, n_u, B, True)
        self.mean = mean

    def vmult(self, dst, src, B):
        super().vmult(dst, src, B)
        self.mean.subtract(dst, B)


class ApproxNestedSchurComplementInverse:
    """nested_schur_complement.hpp:292-320: GMRES(<= 100 steps, 1e-6 |src|, identity preconditioner) on the pressure
    Schur complement, every exception swallowed; then the zero-mean correction when `correct_to_zero_mean`.  (The
    Jacobi object the constructor builds from the preconditioner matrix's block(2,2) is never applied, :270.)"""

    def __init__(self, approx_pressure_schur_complement, mean, correct_to_zero_mean):
        self.op, self.mean, self.correct = approx_pressure_schur_complement, mean, correct_to_zero_mean
        self.iterations = []

    def vmult(self, dst, src, B):
        tol = 1e-6 * math.sqrt(B.dot(src, src))
        try:
            self.iterations.append(solver_gmres(B, self.op, dst, src, Identity(), tol, 100))
        except NoConvergence as e:
            self.iterations.append(e.last_step)
        if self.correct:
            self.mean.subtract(dst, B)


class BlockSchurPreconditionerFEEC:
    """block_schur_preconditioner.hpp:114-147.  Third block: ptmp = -2 src_p (quirk Q8, :137-143) + block(2,1) dst_u."""

    def __init__(self, blocks, mw_inverse, approx_mu_minus_sw_inverse, approx_nested_schur_complement_inverse, sizes, B):
        self.blocks, self.mw_inv = blocks, mw_inverse
        self.shifted_inv, self.nested_inv = approx_mu_minus_sw_inverse, approx_nested_schur_complement_inverse
        self.n_w, self.n_u, self.n_p = sizes
        self.utmp, self.ptmp = B.zeros(self.n_u), B.zeros(self.n_p)

    def vmult(self, dst, src, B):
        n_w, n_u = self.n_w, self.n_u
        dw, du, dp = dst[:n_w], dst[n_w:n_w + n_u], dst[n_w + n_u:]
        sw, su, sp = src[:n_w], src[n_w:n_w + n_u], src[n_w + n_u:]
        B.zero(dst)                                        # fresh destination (the inner solvers start from it)
        self.mw_inv.vmult(dw, sw, B)                                         # :122
        self.blocks[(1, 0)].vmult(self.utmp, dw, B)                          # :131-134
        B.sadd(-1.0, 1.0, su, self.utmp)
        self.shifted_inv.vmult(du, self.utmp, B)
        B.assign(self.ptmp, sp)                                              # :137-145
        B.scale(-2.0, self.ptmp)
        self.blocks[(2, 1)].vmult_add(self.ptmp, du, B)
        self.nested_inv.vmult(dp, self.ptmp, B)


class PreconditionerBlockIdentity:
    """preconditioner_block_identity.hpp:33-58: dst = src, then the pressure block loses its mean (QGauss(2))."""

    def __init__(self, sizes, mean, correct_pressure_mean_value):
        self.sizes, self.mean, self.correct = sizes, mean, correct_pressure_mean_value

    def vmult(self, dst, src, B):
        B.assign(dst, src)
        if self.correct:
            n_w, n_u, _ = self.sizes
            self.mean.subtract(dst[n_w + n_u:], B)


class BlockOperatorN:
    """LA::BlockSparseMatrix::vmult from its (non-empty) blocks, any number of block rows."""

    def __init__(self, blocks, sizes):
        self.blocks = blocks
        self.off = [0]
        for n in sizes:
            self.off.append(self.off[-1] + n)

    def vmult(self, dst, src, B):
        nb = len(self.off) - 1
        for i in range(nb):
            di = dst[self.off[i]:self.off[i + 1]]
            first = True
            for j in range(nb):
                if (i, j) not in self.blocks:
                    continue
                sj = src[self.off[j]:self.off[j + 1]]
                if first:
                    self.blocks[(i, j)].vmult(di, sj, B)
                    first = False
                else:
                    self.blocks[(i, j)].vmult_add(di, sj, B)
            if first:
                B.zero(di)


def solve_nse_block_preconditioned_feec(B, nse_matrix, blocks, mw_jacobi, mu_jacobi, nse_rhs, nse_solution, sizes, dt,
                                        mean_q1, mean_q2, use_block_preconditioner_feec=True,
                                        correct_pressure_to_zero_mean=True, constrained_pressure=None):
    """ExteriorCalculus::BoussinesqModel::solve_NSE_block_preconditioned (boussineq_model_FEEC.tpp:1268-1477) up to
    (not including) nse_constraints.distribute.  blocks: {(i,j): operator with vmult/vmult_add} of the 3x3 nse_matrix
    (w, u, p); mw_jacobi / mu_jacobi: the Jacobi preconditioners of block(0,0) / block(1,1) -- the driver hands these,
    not the CG inverses it also builds, to every composition (:1311-1341, :1408-1413); mean_q1 / mean_q2: MeanValue
    with the QGauss(1) / QGauss(2) cell volumes.  GMRES: 100 temporary vectors (:1397-1401), 500 steps with the block
    preconditioner, 15 000 without (:1372-1378).  The pre-correction of :1380-1394 shifts nse_solution, not the vector
    that is solved for, so it does not enter the result (reproduced by leaving it out of the solve).
    Returns (solution with the SCALED pressure, outer iterations, dict of inner iteration lists)."""
    n_w, n_u, n_p = sizes
    x = B.copy(nse_solution)
    xp = x[n_w + n_u:]
    B.scale(dt, xp)                                                         # :1351
    if constrained_pressure is not None and len(constrained_pressure):      # :1357-1370
        xp[constrained_pressure] = 0.0
    tol = 1e-8 * math.sqrt(B.dot(nse_rhs, nse_rhs))                         # :1374
    if use_block_preconditioner_feec:
        shifted_inv = ApproxShiftedSchurComplementInverse(blocks, mw_jacobi, mu_jacobi, n_w, B)          # :1311-1320
        lower = SchurComplementLowerBlock(blocks, shifted_inv, mu_jacobi, n_u, B, do_full_solve=False)   # :1322-1332
        nested_inv = ApproxNestedSchurComplementInverse(lower, mean_q1, correct_pressure_to_zero_mean)   # :1334-1341
        P = BlockSchurPreconditionerFEEC(blocks, mw_jacobi, shifted_inv, nested_inv, sizes, B)           # :1408-1413
        its = solver_gmres(B, nse_matrix, x, nse_rhs, P, tol, 500, restart=98)                           # :1415-1418
        inner = dict(shifted=shifted_inv.iterations, nested=nested_inv.iterations)
    else:
        P = PreconditionerBlockIdentity(sizes, mean_q2, correct_pressure_to_zero_mean)                   # :1422-1431
        its = solver_gmres(B, nse_matrix, x, nse_rhs, P, tol, 15000, restart=98)
        inner = {}
    return x, its, inner


def feec_cell_volumes(cell_vertices, n_gauss):
    """|K| of every hexahedron under the trilinear (MappingQ1) map with the n_gauss^3-point Gauss rule -- what
    compute_mean_value integrates for the piecewise-constant pressure.  cell_vertices: [cells][8][3], vertices in
    deal.II's lexicographic order."""
    v = np.asarray(cell_vertices, dtype=np.float64).reshape(-1, 8, 3)
    pts, wts = np.polynomial.legendre.leggauss(n_gauss)
    pts, wts = 0.5 * (pts + 1.0), 0.5 * wts
    vol = np.zeros(v.shape[0])
    for kz, z in enumerate(pts):
        for ky, y in enumerate(pts):
            for kx, x in enumerate(pts):
                dN = np.zeros((8, 3))
                for n in range(8):
                    ix, iy, iz = n & 1, (n >> 1) & 1, (n >> 2) & 1
                    fx, fy, fz = (x if ix else 1 - x), (y if iy else 1 - y), (z if iz else 1 - z)
                    dN[n] = ((1 if ix else -1) * fy * fz, fx * (1 if iy else -1) * fz, fx * fy * (1 if iz else -1))
                J = np.einsum("cnd,ne->cde", v, dN)                # J[c][d][e] = d x_d / d xi_e
                vol += wts[kx] * wts[ky] * wts[kz] * np.abs(np.linalg.det(J))
    return vol


def solve_temperature(B, temperature_matrix, t_preconditioner, temperature_rhs, temperature_solution):
    """boussinesq_model.tpp:1417-1440.  Returns (solution, CG iterations)."""
    x = B.copy(temperature_solution)
    n = x.numel() if hasattr(x, "numel") else len(x)
    tol = 1e-12 * math.sqrt(B.dot(temperature_rhs, temperature_rhs))
    its = solver_cg(B, temperature_matrix, x, temperature_rhs, t_preconditioner, tol, n)
    return x, its
