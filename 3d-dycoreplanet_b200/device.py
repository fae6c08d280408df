"""ctypes binding of the device library (lib/libdcp.so, include/dcp.h) and a host-side mirror of the
reference's model interface for the hot path.

`BoussinesqModel` keeps the member names of Standard::BoussinesqModel
(/root/reference/include/core/boussinesq_model.h:168-180): assemble_nse_system, assemble_nse_preconditioner
(+ build_nse_preconditioner), assemble_temperature_matrix, assemble_temperature_rhs; `SparseMatrix` offers the
deal.II operator concept the linear_algebra/ templates are written against (vmult, vmult_add, m, n)
and `BlockSparseMatrix.block(i, j)`.

There is no CPU fallback: without libdcp.so or without a CUDA device every call raises.
"""
import ctypes
import sys
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None

HOST, DEVICE = 0, 1
MAT_NSE, MAT_NSE_PRECOND, MAT_TEMP_MASS, MAT_TEMP_STIFF, MAT_TEMP = 0, 1, 2, 3, 4
VEC_NSE_RHS, VEC_TEMP_RHS = 0, 1
STRATEGY_SEARCH, STRATEGY_POSITIONS, STRATEGY_OWNER, STRATEGY_STAGED = 0, 1, 2, 3
MAXB = 3

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_lp = ctypes.POINTER(ctypes.c_int64)


class Params(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32), ("cuboid", ctypes.c_int32), ("nse_interval", ctypes.c_int32),
                ("pad", ctypes.c_int32), ("dt", ctypes.c_double), ("inv_re", ctypes.c_double),
                ("inv_pe", ctypes.c_double), ("beta", ctypes.c_double), ("T_ref", ctypes.c_double),
                ("g_scale", ctypes.c_double), ("g_const", ctypes.c_double), ("cor_scale", ctypes.c_double),
                ("omega", ctypes.c_double)]


class ConstraintsDesc(ctypes.Structure):
    _fields_ = [("n_dofs", ctypes.c_int64), ("n_lines", ctypes.c_int64), ("line_dof", c_ip), ("line_ptr", c_ip),
                ("entry_dof", c_ip), ("entry_w", c_dp), ("inhom", c_dp)]


class CsrDesc(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_int64), ("n_cols", ctypes.c_int64), ("rowptr", c_lp), ("col", c_ip)]


class ModelDesc(ctypes.Structure):
    _fields_ = [
        ("dim", ctypes.c_int32), ("family", ctypes.c_int32), ("n_cells", ctypes.c_int64),
        ("nse_n_local", ctypes.c_int32), ("nse_n_blocks", ctypes.c_int32), ("nse_block_size", ctypes.c_int64 * MAXB),
        ("nse_l2g", c_ip), ("nse_local_field", c_ip), ("nse_local_base", c_ip), ("nse_cs", ConstraintsDesc),
        ("temp_n_local", ctypes.c_int32), ("pad0", ctypes.c_int32), ("temp_l2g", c_ip), ("temp_cs", ConstraintsDesc),
        ("nq_nse", ctypes.c_int32), ("nq_temp", ctypes.c_int32), ("ndu", ctypes.c_int32), ("ndp", ctypes.c_int32),
        ("ndt", ctypes.c_int32), ("build_owner_plan", ctypes.c_int32),
        ("phi_u_qn", c_dp), ("dphi_u_qn", c_dp), ("phi_p_qn", c_dp), ("phi_t_qn", c_dp),
        ("phi_u_qt", c_dp), ("phi_t_qt", c_dp), ("dphi_t_qt", c_dp),
        ("geom_qn", c_dp), ("geom_qt", c_dp),
        ("nse_pattern", (CsrDesc * MAXB) * MAXB), ("pre_pattern", (CsrDesc * MAXB) * MAXB), ("temp_pattern", CsrDesc),
        # FEEC family
        ("nq_pre", ctypes.c_int32), ("pad2", ctypes.c_int32), ("nse_sign", c_dp),
        ("feec_phi_w_qn", c_dp), ("feec_curl_w_qn", c_dp), ("feec_phi_u_qn", c_dp),
        ("feec_phi_w_qp", c_dp), ("feec_curl_w_qp", c_dp), ("feec_phi_u_qp", c_dp),
        ("feec_phi_u_qt", c_dp), ("feec_div_u", c_dp), ("geom_qp", c_dp),
        # optional diagnostics inputs
        ("cell_vertices", c_dp), ("n_owned_cells", ctypes.c_int64),
        ("geom_on_device", ctypes.c_int32), ("pad3", ctypes.c_int32),
    ]


class MappingDesc(ctypes.Structure):
    _fields_ = [
        ("dim", ctypes.c_int32), ("nq", ctypes.c_int32), ("extended", ctypes.c_int32), ("n_low", ctypes.c_int32),
        ("n_high", ctypes.c_int32), ("pad", ctypes.c_int32), ("n_cells", ctypes.c_int64),
        ("support_ptr", c_lp), ("support_points", c_dp), ("N_low", c_dp), ("dN_low", c_dp), ("N_high", c_dp),
        ("dN_high", c_dp), ("weights", c_dp),
    ]


EXPORTS = [
    "dcp_ctx_create", "dcp_ctx_destroy", "dcp_ctx_set_stream", "dcp_ctx_synchronize", "dcp_last_error",
    "dcp_ctx_launch_count", "dcp_malloc", "dcp_free", "dcp_memcpy_h2d", "dcp_memcpy_d2h", "dcp_memcpy_h2d_async", "dcp_memcpy_d2h_async", "dcp_copy_fence", "dcp_copy_synchronize", "dcp_model_create",
    "dcp_model_destroy", "dcp_model_set_strategy", "dcp_model_get_strategy", "dcp_model_set_owned", "dcp_gather_f64", "dcp_scatter_f64", "dcp_assemble_nse_system", "dcp_assemble_nse_preconditioner",
    "dcp_assemble_temperature_matrix", "dcp_assemble_temperature_rhs", "dcp_matrix_info", "dcp_matrix_values_device",
    "dcp_matrix_download", "dcp_matrix_upload", "dcp_vector_device", "dcp_vector_download", "dcp_vmult",
    "dcp_vmult_add", "dcp_block_vmult", "dcp_vmult_rows", "dcp_block_vmult_rows", "dcp_jacobi_vmult", "dcp_vec_dot", "dcp_vec_axpy", "dcp_vec_sadd",
    "dcp_vec_scale", "dcp_vec_copy", "dcp_vec_fill", "dcp_vec_shift", "dcp_vec_mgs", "dcp_cg_solve", "dcp_comm_unique_id", "dcp_comm_create", "dcp_comm_adopt", "dcp_comm_info",
    "dcp_comm_destroy", "dcp_halo_create", "dcp_halo_destroy", "dcp_halo_exchange", "dcp_halo_block_vmult",
    "dcp_vec_dot_allreduce", "dcp_allreduce_max", "dcp_velocity_extrema", "dcp_constraints_distribute",
    "dcp_geometry_create", "dcp_ilu_create", "dcp_ilu_refactor", "dcp_ilu_vmult", "dcp_ilu_levels", "dcp_ilu_destroy",
]


class DcpError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_ROOT, "lib", "libdcp.so")


def lib():
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise DcpError(f"{path} missing: the CUDA extension is not built (run __graft_entry__.build() or `make`)")
        L = ctypes.CDLL(path)
        L.dcp_last_error.restype = ctypes.c_char_p
        L.dcp_ctx_launch_count.restype = ctypes.c_int64
        L.dcp_ctx_launch_count.argtypes = [ctypes.c_void_p]
        vp = ctypes.c_void_p
        L.dcp_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        L.dcp_ctx_destroy.argtypes = [vp]
        L.dcp_ctx_set_stream.argtypes = [vp, vp]
        L.dcp_ctx_synchronize.argtypes = [vp]
        L.dcp_malloc.argtypes = [vp, ctypes.c_int64, ctypes.POINTER(vp)]
        L.dcp_free.argtypes = [vp, vp]
        L.dcp_memcpy_h2d.argtypes = [vp, vp, vp, ctypes.c_int64]
        L.dcp_memcpy_d2h.argtypes = [vp, vp, vp, ctypes.c_int64]
        L.dcp_memcpy_h2d_async.argtypes = [vp, vp, vp, ctypes.c_int64]
        L.dcp_memcpy_d2h_async.argtypes = [vp, vp, vp, ctypes.c_int64]
        L.dcp_copy_fence.argtypes = [vp, ctypes.c_int]
        L.dcp_copy_synchronize.argtypes = [vp]
        L.dcp_model_create.argtypes = [vp, ctypes.POINTER(ModelDesc), ctypes.POINTER(vp)]
        L.dcp_model_destroy.argtypes = [vp]
        L.dcp_model_set_strategy.argtypes = [vp, ctypes.c_int]
        L.dcp_model_get_strategy.argtypes = [vp]
        L.dcp_model_set_owned.argtypes = [vp, c_lp, ctypes.c_int64]
        L.dcp_gather_f64.argtypes = [vp, ctypes.c_int64, vp, vp, vp]
        L.dcp_scatter_f64.argtypes = [vp, ctypes.c_int64, vp, vp, vp]
        L.dcp_assemble_nse_system.argtypes = [vp, ctypes.POINTER(Params), vp, vp, ctypes.c_int]
        L.dcp_assemble_nse_preconditioner.argtypes = [vp, ctypes.POINTER(Params)]
        L.dcp_assemble_temperature_matrix.argtypes = [vp, ctypes.POINTER(Params)]
        L.dcp_assemble_temperature_rhs.argtypes = [vp, ctypes.POINTER(Params), vp, vp, ctypes.c_int]
        L.dcp_matrix_info.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_lp, c_lp, c_lp]
        L.dcp_matrix_values_device.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        L.dcp_matrix_download.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.dcp_matrix_upload.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.dcp_vector_device.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), c_lp]
        L.dcp_vector_download.argtypes = [vp, ctypes.c_int, vp]
        L.dcp_vmult.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int]
        L.dcp_vmult_add.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int]
        L.dcp_block_vmult.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_int]
        L.dcp_vmult_rows.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int]
        L.dcp_block_vmult_rows.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_int]
        L.dcp_jacobi_vmult.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int]
        L.dcp_vec_dot.argtypes = [vp, ctypes.c_int64, vp, vp, c_dp]
        L.dcp_vec_mgs.argtypes = [vp, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(vp), vp, c_dp]
        L.dcp_cg_solve.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp,
                                   ctypes.c_double, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                   ctypes.POINTER(ctypes.c_double)]
        L.dcp_vec_axpy.argtypes = [vp, ctypes.c_int64, ctypes.c_double, vp, vp]
        L.dcp_vec_sadd.argtypes = [vp, ctypes.c_int64, ctypes.c_double, ctypes.c_double, vp, vp]
        L.dcp_vec_scale.argtypes = [vp, ctypes.c_int64, ctypes.c_double, vp]
        L.dcp_vec_copy.argtypes = [vp, ctypes.c_int64, vp, vp]
        L.dcp_vec_fill.argtypes = [vp, ctypes.c_int64, ctypes.c_double, vp]
        L.dcp_vec_shift.argtypes = [vp, ctypes.c_int64, ctypes.c_double, vp]
        L.dcp_comm_unique_id.argtypes = [vp]
        L.dcp_comm_create.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        L.dcp_comm_adopt.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        L.dcp_comm_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.dcp_comm_destroy.argtypes = [vp]
        L.dcp_halo_create.argtypes = [vp, ctypes.c_int64, vp, c_lp, vp, c_lp, ctypes.POINTER(vp)]
        L.dcp_halo_destroy.argtypes = [vp]
        L.dcp_halo_exchange.argtypes = [vp, vp]
        L.dcp_halo_block_vmult.argtypes = [vp, ctypes.c_int, vp, vp, vp, ctypes.c_int]
        L.dcp_vec_dot_allreduce.argtypes = [vp, ctypes.c_int, c_lp, c_lp, vp, vp, ctypes.POINTER(ctypes.c_double)]
        L.dcp_allreduce_max.argtypes = [vp, ctypes.c_int, c_dp]
        L.dcp_geometry_create.argtypes = [vp, ctypes.POINTER(MappingDesc), ctypes.POINTER(vp)]
        L.dcp_ilu_create.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        L.dcp_ilu_refactor.argtypes = [vp]
        L.dcp_ilu_vmult.argtypes = [vp, vp, vp, ctypes.c_int]
        L.dcp_ilu_levels.argtypes = [vp, c_lp, c_lp]
        L.dcp_ilu_destroy.argtypes = [vp]
        L.dcp_velocity_extrema.argtypes = [vp, vp, ctypes.c_int, c_dp]
        L.dcp_constraints_distribute.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int]
        _LIB = L
    return _LIB


def check(rc, what=""):
    if rc != 0:
        raise DcpError(f"{what} failed (status {rc}): {lib().dcp_last_error().decode()}")


def params_from(mp):
    """dcp_params from a params.ModelParameters (derivations: SURVEY.md Appendix B)."""
    return Params(dim=mp.space_dimension, cuboid=int(mp.cuboid_geometry), nse_interval=mp.NSE_solver_interval, pad=0,
                  dt=mp.time_step, inv_re=mp.inv_re, inv_pe=mp.inv_pe, beta=mp.expansion_coefficient,
                  T_ref=mp.ref_temperature, g_scale=mp.g_scale, g_const=mp.gravity_constant, cor_scale=mp.cor_scale,
                  omega=mp.omega)


def _ptr(a, ctype):
    if a is None or a.size == 0:
        return ctypes.cast(None, ctype)
    assert a.flags.c_contiguous
    return a.ctypes.data_as(ctype)


def _vec_arg(x):
    """(pointer value, mem flag) of a numpy array (host) or a torch CUDA tensor (device)."""
    if isinstance(x, np.ndarray):
        assert x.dtype == np.float64 and x.flags.c_contiguous
        return ctypes.c_void_p(x.ctypes.data), HOST
    if isinstance(x, int):
        return ctypes.c_void_p(x), DEVICE
    # torch tensor
    assert x.dtype.is_floating_point and x.element_size() == 8 and x.is_contiguous()
    return ctypes.c_void_p(x.data_ptr()), (DEVICE if x.is_cuda else HOST)


class Context:
    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        check(lib().dcp_ctx_create(device, ctypes.byref(self._h)), "dcp_ctx_create")
        self.device = device

    def set_stream(self, cuda_stream_ptr):
        """Run the library's work on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream).  Handle 0 is
        torch's default stream = the legacy default stream, passed as cudaStreamLegacy (1); None = the own stream."""
        if cuda_stream_ptr is None:
            ptr = None
        else:
            ptr = 1 if int(cuda_stream_ptr) == 0 else int(cuda_stream_ptr)
        check(lib().dcp_ctx_set_stream(self._h, ctypes.c_void_p(ptr)), "dcp_ctx_set_stream")

    def synchronize(self):
        check(lib().dcp_ctx_synchronize(self._h), "dcp_ctx_synchronize")

    def download_f64(self, dev_ptr, n):
        """n doubles from a device pointer of this context (dcp_memcpy_d2h)."""
        out = np.empty(n)
        check(lib().dcp_memcpy_d2h(self._h, ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(dev_ptr), 8 * n), "dcp_memcpy_d2h")
        return out

    def free(self, dev_ptr):
        check(lib().dcp_free(self._h, ctypes.c_void_p(dev_ptr)), "dcp_free")

    def launch_count(self):
        return int(lib().dcp_ctx_launch_count(self._h))

    def close(self):
        if self._h:
            lib().dcp_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        if sys.is_finalizing():   # teardown order of native libraries is undefined at exit; the OS reclaims everything
            return
        try:
            self.close()
        except Exception:
            pass


class SparseMatrix:
    """One block: the `LA::SparseMatrix` surface the reference's operator templates use."""

    def __init__(self, model, which, bi, bj):
        self._m, self.which, self.bi, self.bj = model, which, bi, bj
        nr, nc, nz = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(lib().dcp_matrix_info(model._h, which, bi, bj, ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(nz)))
        self._nr, self._nc, self.nnz = nr.value, nc.value, nz.value

    @property
    def n_rows(self):
        return self._nr

    @property
    def n_cols(self):
        return self._nc

    def m(self):
        return self._nr

    def n(self):
        return self._nc

    def vmult(self, dst, src):
        d, md = _vec_arg(dst)
        s, ms = _vec_arg(src)
        assert md == ms
        check(lib().dcp_vmult(self._m._h, self.which, self.bi, self.bj, d, s, md), "vmult")

    def vmult_rows(self, dst, src, rows):
        """vmult restricted to a row class (ROWS_INTERIOR / ROWS_GHOSTED), device vectors only."""
        d, md = _vec_arg(dst)
        s, ms = _vec_arg(src)
        assert md == ms == DEVICE
        check(lib().dcp_vmult_rows(self._m._h, self.which, self.bi, self.bj, d, s, rows), "dcp_vmult_rows")

    def vmult_add(self, dst, src):
        d, md = _vec_arg(dst)
        s, ms = _vec_arg(src)
        assert md == ms
        check(lib().dcp_vmult_add(self._m._h, self.which, self.bi, self.bj, d, s, md), "vmult_add")

    def values(self):
        out = np.zeros(self.nnz)
        if self.nnz:
            check(lib().dcp_matrix_download(self._m._h, self.which, self.bi, self.bj, ctypes.c_void_p(out.ctypes.data)))
        return out

    def set_values(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.nnz
        if self.nnz:
            check(lib().dcp_matrix_upload(self._m._h, self.which, self.bi, self.bj, ctypes.c_void_p(v.ctypes.data)))

    def values_device_ptr(self):
        p = ctypes.c_void_p()
        check(lib().dcp_matrix_values_device(self._m._h, self.which, self.bi, self.bj, ctypes.byref(p)))
        return p.value


class BlockSparseMatrix:
    def __init__(self, model, which, nb):
        self._m, self.which, self.nb = model, which, nb

    def block(self, i, j):
        return SparseMatrix(self._m, self.which, i, j)

    def vmult(self, dst, src):
        d, md = _vec_arg(dst)
        s, ms = _vec_arg(src)
        assert md == ms
        check(lib().dcp_block_vmult(self._m._h, self.which, d, s, md), "block vmult")

    def vmult_rows(self, dst, src, rows):
        d, md = _vec_arg(dst)
        s, ms = _vec_arg(src)
        assert md == ms == DEVICE
        check(lib().dcp_block_vmult_rows(self._m._h, self.which, d, s, rows), "dcp_block_vmult_rows")


ROWS_ALL, ROWS_INTERIOR, ROWS_GHOSTED = 0, 1, 2


PRECOND_IDENTITY, PRECOND_JACOBI, PRECOND_ILU = 0, 1, 2
ERR_NO_CONVERGENCE = 5


def cg_solve(matrix, x, b, tol, max_steps, preconditioner=None, check_every=8):
    """dcp_cg_solve: SolverCG on the device without a host synchronisation per iteration.  `matrix`: SparseMatrix (a
    square block), `preconditioner`: None (identity), PreconditionJacobi or PreconditionILU; x, b: device tensors.
    Returns (converged, last_step, last_residual)."""
    kind, which_p, bp, ilu = PRECOND_IDENTITY, 0, 0, None
    if isinstance(preconditioner, PreconditionJacobi):
        kind, which_p, bp = PRECOND_JACOBI, preconditioner.which, preconditioner.bi
    elif isinstance(preconditioner, PreconditionILU):
        kind, ilu = PRECOND_ILU, preconditioner._h
    elif preconditioner is not None:
        raise TypeError("cg_solve: identity, Jacobi or ILU(0) preconditioners only")
    xd, mx = _vec_arg(x)
    bd, mb = _vec_arg(b)
    assert mx == DEVICE and mb == DEVICE, "cg_solve works on device vectors"
    step, res = ctypes.c_int64(), ctypes.c_double()
    rc = lib().dcp_cg_solve(matrix._m._h, matrix.which, matrix.bi, matrix.bj, kind, which_p, bp, ilu, xd, bd, float(tol),
                            int(min(max_steps, 2 ** 62)), int(check_every), ctypes.byref(step), ctypes.byref(res))
    if rc not in (0, ERR_NO_CONVERGENCE):
        check(rc, "dcp_cg_solve")
    return rc == 0, step.value, res.value


class PreconditionJacobi:
    """LA::PreconditionJacobi on a diagonal block (one sweep, omega = 1)."""

    def __init__(self, model, which, bi):
        self._m, self.which, self.bi = model, which, bi

    def vmult(self, dst, src):
        d, md = _vec_arg(dst)
        s, ms = _vec_arg(src)
        assert md == ms
        check(lib().dcp_jacobi_vmult(self._m._h, self.which, self.bi, d, s, md), "jacobi vmult")


class PreconditionILU:
    """LA::PreconditionILU (Ifpack ILU(0), deal.II defaults) of a diagonal block, factorised on the device."""

    def __init__(self, model, which, bi):
        self._m = model
        self._h = ctypes.c_void_p()
        check(lib().dcp_ilu_create(model._h, which, bi, ctypes.byref(self._h)), "dcp_ilu_create")

    def refactor(self):
        check(lib().dcp_ilu_refactor(self._h), "dcp_ilu_refactor")

    def levels(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        check(lib().dcp_ilu_levels(self._h, ctypes.byref(a), ctypes.byref(b)), "dcp_ilu_levels")
        return a.value, b.value

    def vmult(self, dst, src):
        d, md = _vec_arg(dst)
        s, ms = _vec_arg(src)
        assert md == ms
        check(lib().dcp_ilu_vmult(self._h, d, s, md), "dcp_ilu_vmult")

    def close(self):
        if self._h and self._m._h:      # a closed model has already destroyed its ILU handles
            lib().dcp_ilu_destroy(self._h)
        self._h = ctypes.c_void_p()

    def __del__(self):
        if sys.is_finalizing():   # teardown order of native libraries is undefined at exit; the OS reclaims everything
            return
        try:
            self.close()
        except Exception:
            pass


def _csr_desc(P, name, keep):
    d = CsrDesc()
    if name is None:
        return d
    rp, col = P[name + ".rowptr"], P[name + ".col"]
    d.n_rows, d.n_cols = P.scalar(name + ".n_rows"), P.scalar(name + ".n_cols")
    if len(col) == 0:
        return d
    keep += [rp, col]
    d.rowptr, d.col = _ptr(rp, c_lp), _ptr(col, c_ip)
    return d


def _cs_desc(P, prefix, keep):
    d = ConstraintsDesc()
    ld = P[prefix + ".line_dof"]
    d.n_dofs = len(P[prefix + ".line_of_dof"])
    d.n_lines = len(ld)
    arrs = [ld, P[prefix + ".line_ptr"], P[prefix + ".entry_dof"], P[prefix + ".entry_w"], P[prefix + ".inhom"]]
    keep += arrs
    d.line_dof, d.line_ptr, d.entry_dof = _ptr(arrs[0], c_ip), _ptr(arrs[1], c_ip), _ptr(arrs[2], c_ip)
    d.entry_w, d.inhom = _ptr(arrs[3], c_dp), _ptr(arrs[4], c_dp)
    return d


def geometry_create(ctx, P, rule):
    """Mapping records of quadrature rule `rule` ("qn", "qt" or "qp") evaluated on the device from the cells'
    mapping support points (dcp_geometry_create).  Returns the device pointer (int)."""
    feec = "feec" in P.spec.get("family", "classic")
    md = MappingDesc()
    md.dim, md.nq, md.extended = P.dim, P[f"map.{rule}.w"].size, (1 if feec else 0)
    md.n_low, md.n_high, md.n_cells = P.scalar("map.n_low"), P.scalar("map.n_high"), P.n_cells
    keep = [P["map.ptr"], P["map.points"]] + [P[f"map.{rule}.{k}"] for k in ("N_low", "dN_low", "N_high", "dN_high", "w")]
    md.support_ptr, md.support_points = _ptr(keep[0], c_lp), _ptr(keep[1], c_dp)
    md.N_low, md.dN_low, md.N_high, md.dN_high, md.weights = (_ptr(a, c_dp) for a in keep[2:])
    out = ctypes.c_void_p()
    check(lib().dcp_geometry_create(ctx._h, ctypes.byref(md), ctypes.byref(out)), "dcp_geometry_create")
    return out.value


def model_desc_from_problem(P, owner_plan=False, device_geometry=None):
    """Fill a dcp_model_desc from a harness Problem (the stand-in for deal.II's objects).

    device_geometry: a Context -> the mapping records are evaluated on that device from the support points
    (dcp_geometry_create) and adopted by the model instead of uploading the host records."""
    keep = []
    d = ModelDesc()
    d.build_owner_plan = 1 if owner_plan else 0
    feec = "feec" in P.spec.get("family", "classic")
    d.dim, d.family, d.n_cells = P.dim, (1 if feec else 0), P.n_cells
    d.nse_n_local = P.scalar("nse.n_local")
    if feec:
        d.nse_n_blocks = 3
        d.nse_block_size[0], d.nse_block_size[1], d.nse_block_size[2] = \
            P.scalar("nse.n_w"), P.scalar("nse.n_u"), P.scalar("nse.n_p")
        names = [("nse.l2g", "nse_l2g", c_ip), ("nse.local_field", "nse_local_field", c_ip),
                 ("nse.local_base", "nse_local_base", c_ip), ("temp.l2g", "temp_l2g", c_ip),
                 ("tab.t_qn.phi", "phi_t_qn", c_dp), ("tab.t_qt.phi", "phi_t_qt", c_dp),
                 ("tab.t_qt.dphi", "dphi_t_qt", c_dp), ("geom.qn", "geom_qn", c_dp), ("geom.qt", "geom_qt", c_dp),
                 ("geom.qp", "geom_qp", c_dp), ("nse.sign", "nse_sign", c_dp),
                 ("feec.qn.phi_w", "feec_phi_w_qn", c_dp), ("feec.qn.curl_w", "feec_curl_w_qn", c_dp),
                 ("feec.qn.phi_u", "feec_phi_u_qn", c_dp), ("feec.qp.phi_w", "feec_phi_w_qp", c_dp),
                 ("feec.qp.curl_w", "feec_curl_w_qp", c_dp), ("feec.qp.phi_u", "feec_phi_u_qp", c_dp),
                 ("feec.qt.phi_u", "feec_phi_u_qt", c_dp), ("feec.qn.div_u", "feec_div_u", c_dp)]
        d.nq_pre = P.scalar("q_pre.nq")
        d.ndu, d.ndp, d.ndt = 0, 0, P.scalar("tab.t_qn.nd")
        nb = 3
    else:
        d.nse_n_blocks = 2
        d.nse_block_size[0], d.nse_block_size[1] = P.scalar("nse.n_u"), P.scalar("nse.n_p")
        names = [("nse.l2g", "nse_l2g", c_ip), ("nse.local_field", "nse_local_field", c_ip),
                 ("nse.local_base", "nse_local_base", c_ip), ("temp.l2g", "temp_l2g", c_ip),
                 ("tab.u_qn.phi", "phi_u_qn", c_dp), ("tab.u_qn.dphi", "dphi_u_qn", c_dp),
                 ("tab.p_qn.phi", "phi_p_qn", c_dp), ("tab.t_qn.phi", "phi_t_qn", c_dp),
                 ("tab.u_qt.phi", "phi_u_qt", c_dp), ("tab.t_qt.phi", "phi_t_qt", c_dp),
                 ("tab.t_qt.dphi", "dphi_t_qt", c_dp), ("geom.qn", "geom_qn", c_dp), ("geom.qt", "geom_qt", c_dp)]
        d.ndu, d.ndp, d.ndt = P.scalar("tab.u_qn.nd"), P.scalar("tab.p_qn.nd"), P.scalar("tab.t_qn.nd")
        nb = 2
    for name, field, ct in names:
        if name.startswith("geom.") and device_geometry is not None:
            continue
        a = P[name]
        keep.append(a)
        setattr(d, field, _ptr(a, ct))
    if device_geometry is not None:
        gn = geometry_create(device_geometry, P, "qn")
        gt = gn if P.scalar("geom_shared") else geometry_create(device_geometry, P, "qt")
        d.geom_qn, d.geom_qt = ctypes.cast(gn, c_dp), ctypes.cast(gt, c_dp)
        if feec:
            d.geom_qp = ctypes.cast(geometry_create(device_geometry, P, "qp"), c_dp)
        d.geom_on_device = 1
    d.nse_cs = _cs_desc(P, "nse.cs", keep)
    d.temp_cs = _cs_desc(P, "temp.cs", keep)
    d.temp_n_local = P.scalar("temp.n_local")
    d.nq_nse, d.nq_temp = P.scalar("q_nse.nq"), P.scalar("q_temp.nq")
    for i in range(nb):
        for j in range(nb):
            d.nse_pattern[i][j] = _csr_desc(P, f"nse.b{i}{j}", keep)
            d.pre_pattern[i][j] = _csr_desc(P, f"pre.b{i}{j}", keep)
    d.temp_pattern = _csr_desc(P, "temp.pat", keep)
    if "cell_vertices" in P.names():
        a = P["cell_vertices"]
        keep.append(a)
        d.cell_vertices = _ptr(a, c_dp)
    d.n_owned_cells = P.scalar("n_owned_cells")
    d._keep = keep
    return d


class BoussinesqModel:
    """Device-resident state of the hot path of Standard::BoussinesqModel<dim> for one mesh."""

    def __init__(self, ctx, desc, parameters):
        self.ctx = ctx
        self._h = ctypes.c_void_p()
        check(lib().dcp_model_create(ctx._h, ctypes.byref(desc), ctypes.byref(self._h)), "dcp_model_create")
        self.prm = parameters if isinstance(parameters, Params) else params_from(parameters)
        self.nb = desc.nse_n_blocks
        self.n_nse = sum(desc.nse_block_size[b] for b in range(self.nb))
        self.n_temp = desc.temp_cs.n_dofs
        self.nse_matrix = BlockSparseMatrix(self, MAT_NSE, self.nb)
        self.nse_preconditioner_matrix = BlockSparseMatrix(self, MAT_NSE_PRECOND, self.nb)
        self.temperature_mass_matrix = SparseMatrix(self, MAT_TEMP_MASS, 0, 0)
        self.temperature_stiffness_matrix = SparseMatrix(self, MAT_TEMP_STIFF, 0, 0)
        self.temperature_matrix = SparseMatrix(self, MAT_TEMP, 0, 0)
        self.Mu_plus_A_preconditioner = PreconditionJacobi(self, MAT_NSE_PRECOND, 0)
        self.Mp_preconditioner = PreconditionJacobi(self, MAT_NSE_PRECOND, 1)
        self.T_preconditioner = PreconditionJacobi(self, MAT_TEMP, 0)

    @classmethod
    def from_problem(cls, ctx, P, parameters, owner_plan=False, device_geometry=False):
        return cls(ctx, model_desc_from_problem(P, owner_plan, ctx if device_geometry else None), parameters)

    def set_strategy(self, s):
        check(lib().dcp_model_set_strategy(self._h, s), "dcp_model_set_strategy")

    @property
    def strategy(self):
        return int(lib().dcp_model_get_strategy(self._h))

    def set_owned(self, nse_owned_per_block, temp_owned):
        arr = (ctypes.c_int64 * MAXB)(*(list(nse_owned_per_block) + [0] * (MAXB - len(nse_owned_per_block))))
        check(lib().dcp_model_set_owned(self._h, arr, int(temp_owned)), "dcp_model_set_owned")

    # --- the four assemblers (boussinesq_model.h:168-180) ---------------------------------------
    def assemble_nse_system(self, old_nse_solution, old_temperature_solution):
        a, ma = _vec_arg(old_nse_solution)
        b, mb = _vec_arg(old_temperature_solution)
        assert ma == mb
        check(lib().dcp_assemble_nse_system(self._h, ctypes.byref(self.prm), a, b, ma), "assemble_nse_system")

    def assemble_nse_preconditioner(self):
        check(lib().dcp_assemble_nse_preconditioner(self._h, ctypes.byref(self.prm)), "assemble_nse_preconditioner")

    build_nse_preconditioner = assemble_nse_preconditioner  # Jacobi set-up is part of the device call

    def assemble_temperature_matrix(self):
        check(lib().dcp_assemble_temperature_matrix(self._h, ctypes.byref(self.prm)), "assemble_temperature_matrix")

    def assemble_temperature_rhs(self, old_temperature_solution, nse_solution):
        a, ma = _vec_arg(old_temperature_solution)
        b, mb = _vec_arg(nse_solution)
        assert ma == mb
        check(lib().dcp_assemble_temperature_rhs(self._h, ctypes.byref(self.prm), a, b, ma), "assemble_temperature_rhs")

    # --- passes next to the solves -----------------------------------------------------------------
    def velocity_extrema(self, nse_solution):
        """(get_maximal_velocity(), get_cfl_number()) of boussinesq_model.tpp:1023-1098, rank-local values."""
        a, ma = _vec_arg(nse_solution)
        out = (ctypes.c_double * 2)()
        check(lib().dcp_velocity_extrema(self._h, a, ma, out), "dcp_velocity_extrema")
        return float(out[0]), float(out[1])

    def get_maximal_velocity(self, nse_solution):
        return self.velocity_extrema(nse_solution)[0]

    def get_cfl_number(self, nse_solution):
        return self.velocity_extrema(nse_solution)[1]

    def distribute_nse_constraints(self, x):
        """nse_constraints.distribute(x) (boussinesq_model.tpp:1233), in place."""
        a, ma = _vec_arg(x)
        check(lib().dcp_constraints_distribute(self._h, 0, a, ma), "dcp_constraints_distribute")

    def distribute_temperature_constraints(self, x):
        """temperature_constraints.distribute(x) (boussinesq_model.tpp:1442), in place."""
        a, ma = _vec_arg(x)
        check(lib().dcp_constraints_distribute(self._h, 1, a, ma), "dcp_constraints_distribute")

    # --- results --------------------------------------------------------------------------------
    def vector(self, which):
        n = self.n_nse if which == VEC_NSE_RHS else self.n_temp
        out = np.zeros(n)
        check(lib().dcp_vector_download(self._h, which, ctypes.c_void_p(out.ctypes.data)), "dcp_vector_download")
        return out

    @property
    def nse_rhs(self):
        return self.vector(VEC_NSE_RHS)

    @property
    def temperature_rhs(self):
        return self.vector(VEC_TEMP_RHS)

    def close(self):
        if self._h:
            lib().dcp_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        if sys.is_finalizing():   # teardown order of native libraries is undefined at exit; the OS reclaims everything
            return
        try:
            self.close()
        except Exception:
            pass
