"""One-process-per-GPU plumbing for the partitioned hot path: ghost-dof halo exchange for SpMV.

The reference shards this path over MPI ranks along the p4est curve (`LocallyOwnedCell` filter,
/root/reference/include/core/boussinesq_model.tpp:491-494; owned / relevant index sets :241-252) and every
`LA::SparseMatrix::vmult` hides an Epetra_Import of the off-rank x entries.  Here:

  * assembly needs no collective: every rank integrates its owned cells plus the one-deep ghost layer, so
    each owned row is complete (replaces `compress(VectorOperation::add)`, :513, 736-737);
  * SpMV needs exactly one exchange per product: owned boundary entries are packed with the library's gather
    kernel, moved with grouped point-to-point sends (torch.distributed, NCCL over NVLink on GPUs, gloo in the
    CPU tests) and unpacked with the scatter kernel into the ghost slots of the source vector.

torch is used for device memory, streams and the process group only.  On GPUs the data plane is the library's own
(`Communicator`, `DeviceHalo`: dcp_comm_* / dcp_halo_* / dcp_vec_dot_allreduce in include/dcp.h, NCCL inside libdcp);
torch.distributed only carries the 128-byte NCCL id and the index lists at set-up.  `HaloPlan.exchange` (torch p2p) is
what the CPU (gloo) tests run.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist


class HaloPlan:
    """Owned/ghost maps of one vector layout on this rank.

    keys[i]   : global identity of local dof i (harness `*.dof_key`; a deal.II build would use the global
                dof index)
    owners[i] : owning rank of local dof i
    """

    def __init__(self, keys, owners, rank, world, group=None, device=None):
        self.rank, self.world, self.group = rank, world, group
        keys = np.asarray(keys, dtype=np.int64)
        owners = np.asarray(owners, dtype=np.int32)
        self.n_local = len(keys)
        ghost = np.flatnonzero(owners != rank)
        order = np.argsort(owners[ghost], kind="stable")
        ghost = ghost[order]
        self.recv_idx = ghost.astype(np.int32)                       # local slots, grouped by owner
        self.recv_counts = np.bincount(owners[ghost], minlength=world).astype(np.int64)
        want = {}
        off = 0
        for p in range(world):
            n = int(self.recv_counts[p])
            if n:
                want[p] = keys[ghost[off:off + n]]
            off += n
        # tell every owner which of its dofs we need, in our order
        if world > 1:
            all_want = [None] * world
            dist.all_gather_object(all_want, want, group=group)
        else:
            all_want = [want]
        owned = np.flatnonzero(owners == rank)
        okeys = keys[owned]
        srt = np.argsort(okeys)
        okeys_sorted = okeys[srt]
        send_idx, send_counts = [], np.zeros(world, dtype=np.int64)
        for q in range(world):
            wk = all_want[q].get(rank) if q != rank else None
            if wk is None or len(wk) == 0:
                continue
            pos = np.searchsorted(okeys_sorted, wk)
            if (pos >= len(okeys_sorted)).any() or (okeys_sorted[np.minimum(pos, len(okeys_sorted) - 1)] != wk).any():
                raise RuntimeError(f"rank {rank}: rank {q} asks for dofs this rank does not own")
            send_idx.append(owned[srt[pos]].astype(np.int32))
            send_counts[q] = len(wk)
        self.send_idx = np.concatenate(send_idx) if send_idx else np.zeros(0, dtype=np.int32)
        self.send_counts = send_counts
        self.device = device
        dev = device if device is not None else "cpu"
        self.t_send_idx = torch.from_numpy(self.send_idx).to(dev)
        self.t_recv_idx = torch.from_numpy(self.recv_idx).to(dev)
        self.send_buf = torch.zeros(max(len(self.send_idx), 1), dtype=torch.float64, device=dev)
        self.recv_buf = torch.zeros(max(len(self.recv_idx), 1), dtype=torch.float64, device=dev)

    @property
    def bytes_per_exchange(self):
        return 8 * int(len(self.send_idx))

    def exchange(self, x, ctx=None):
        """Refresh the ghost entries of the local vector `x` (torch tensor, CPU or CUDA) in place."""
        if self.world == 1:
            return
        ns, nr = len(self.send_idx), len(self.recv_idx)
        if x.is_cuda and ctx is not None:
            from . import device as dv
            if ns:
                dv.check(dv.lib().dcp_gather_f64(ctx._h, ns, ctypes.c_void_p(self.t_send_idx.data_ptr()),
                                                 ctypes.c_void_p(x.data_ptr()),
                                                 ctypes.c_void_p(self.send_buf.data_ptr())), "dcp_gather_f64")
        elif ns:
            self.send_buf[:ns] = x[self.t_send_idx.long()]
        ops, so, ro = [], 0, 0
        for p in range(self.world):
            n = int(self.send_counts[p])
            if n:
                ops.append(dist.P2POp(dist.isend, self.send_buf[so:so + n], p, group=self.group))
            so += n
        for p in range(self.world):
            n = int(self.recv_counts[p])
            if n:
                ops.append(dist.P2POp(dist.irecv, self.recv_buf[ro:ro + n], p, group=self.group))
            ro += n
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        if x.is_cuda and ctx is not None:
            from . import device as dv
            if nr:
                dv.check(dv.lib().dcp_scatter_f64(ctx._h, nr, ctypes.c_void_p(self.t_recv_idx.data_ptr()),
                                                  ctypes.c_void_p(self.recv_buf.data_ptr()),
                                                  ctypes.c_void_p(x.data_ptr())), "dcp_scatter_f64")
        elif nr:
            x[self.t_recv_idx.long()] = self.recv_buf[:nr]


class Communicator:
    """The library's NCCL communicator (dcp_comm_create): rank 0 draws the id, the process group broadcasts it."""

    def __init__(self, ctx, rank, world, group=None):
        from . import device as dv
        self.dv, self.ctx, self.rank, self.world = dv, ctx, rank, world
        ident = [None]
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            dv.check(dv.lib().dcp_comm_unique_id(buf), "dcp_comm_unique_id")
            ident = [bytes(buf.raw)]
        if world > 1:
            dist.broadcast_object_list(ident, src=0, group=group)
        self._h = ctypes.c_void_p()
        dv.check(dv.lib().dcp_comm_create(ctx._h, ctypes.c_char_p(ident[0]), rank, world, ctypes.byref(self._h)), "dcp_comm_create")
        self._res = ctypes.c_double()

    def dot(self, x, y, owned_ranges):
        """Inner product over the owned entries (list of (begin, end)), all-reduced: the solvers' MPI_Allreduce."""
        n = len(owned_ranges)
        rb = (ctypes.c_int64 * n)(*[int(b) for b, _ in owned_ranges])
        re = (ctypes.c_int64 * n)(*[int(e) for _, e in owned_ranges])
        self.dv.check(self.dv.lib().dcp_vec_dot_allreduce(self._h, n, rb, re, ctypes.c_void_p(x.data_ptr()),
                                                          ctypes.c_void_p(y.data_ptr()), ctypes.byref(self._res)),
                      "dcp_vec_dot_allreduce")
        return self._res.value

    def max(self, values):
        arr = (ctypes.c_double * len(values))(*[float(v) for v in values])
        self.dv.check(self.dv.lib().dcp_allreduce_max(self._h, len(values), arr), "dcp_allreduce_max")
        return list(arr)

    def close(self):
        if self._h:
            self.dv.lib().dcp_comm_destroy(self._h)
            self._h = None


class DeviceHalo:
    """Ghost exchange and row-distributed products inside libdcp (dcp_halo_create from a HaloPlan's index lists)."""

    def __init__(self, plan, comm):
        from . import device as dv
        self.dv, self.plan, self.comm = dv, plan, comm
        si = np.ascontiguousarray(plan.send_idx, dtype=np.int32)
        ri = np.ascontiguousarray(plan.recv_idx, dtype=np.int32)
        sc = np.ascontiguousarray(plan.send_counts, dtype=np.int64)
        rc = np.ascontiguousarray(plan.recv_counts, dtype=np.int64)
        self._h = ctypes.c_void_p()
        dv.check(dv.lib().dcp_halo_create(comm._h, plan.n_local, si.ctypes.data_as(ctypes.c_void_p),
                                          sc.ctypes.data_as(dv.c_lp), ri.ctypes.data_as(ctypes.c_void_p),
                                          rc.ctypes.data_as(dv.c_lp), ctypes.byref(self._h)), "dcp_halo_create")

    def exchange(self, x):
        self.dv.check(self.dv.lib().dcp_halo_exchange(self._h, ctypes.c_void_p(x.data_ptr())), "dcp_halo_exchange")

    def vmult(self, model, which, dst, src, overlap=False):
        """dst(owned rows) = A src of the row-distributed matrix `which`, ghost refresh of src included."""
        self.dv.check(self.dv.lib().dcp_halo_block_vmult(model._h, which, self._h, ctypes.c_void_p(dst.data_ptr()),
                                                         ctypes.c_void_p(src.data_ptr()), 1 if overlap else 0),
                      "dcp_halo_block_vmult")

    def close(self):
        if self._h:
            self.dv.lib().dcp_halo_destroy(self._h)
            self._h = None


class HaloMatrix:
    """Operator concept (vmult) of a row-distributed matrix through the library's halo product."""

    def __init__(self, model, which, halo, overlap=False):
        self.model, self.which, self.halo, self.overlap = model, which, halo, overlap

    def vmult(self, dst, src):
        self.halo.vmult(self.model, self.which, dst, src, self.overlap)


class DistributedMatrix:
    """`vmult` of a row-distributed matrix: halo exchange of the source, then the owned rows on the device."""

    def __init__(self, matrix, halo, ctx):
        self.matrix, self.halo, self.ctx = matrix, halo, ctx

    def vmult(self, dst, src):
        self.halo.exchange(src, self.ctx)
        self.matrix.vmult(dst, src)


class OverlappedMatrix:
    """`vmult` that hides the ghost exchange behind the rows that do not need it (what Epetra does
    inside Epetra_CrsMatrix::Multiply): the exchange (pack kernel -> NCCL p2p -> unpack kernel) runs on a side
    stream with a context of its own, the main stream computes the rows that read owned columns only, waits for
    the exchange and finishes the rows that read ghost columns.  CUDA tensors only."""

    def __init__(self, matrix, halo, device_index, main_stream):
        from . import device as dv
        self.dv = dv
        self.matrix, self.halo, self.main = matrix, halo, main_stream
        self.side = torch.cuda.Stream(device=device_index)
        self.side_ctx = dv.Context(device_index)
        self.side_ctx.set_stream(self.side.cuda_stream)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()

    def vmult(self, dst, src):
        dv = self.dv
        if self.halo.world == 1:
            self.matrix.vmult(dst, src)
            return
        self.ready.record(self.main)               # src is final on the main stream from here on
        # interior rows first: their kernels run while the host is still issuing the exchange
        self.matrix.vmult_rows(dst, src, dv.ROWS_INTERIOR)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ready)
            self.halo.exchange(src, self.side_ctx)
            self.done.record(self.side)
        self.main.wait_event(self.done)
        self.matrix.vmult_rows(dst, src, dv.ROWS_GHOSTED)

    def close(self):
        self.side_ctx.close()


def global_dot(a, b, owned_mask_or_slices, group=None):
    """Dot product over owned entries + all-reduce (the Krylov solvers' MPI_Allreduce of one double)."""
    s = torch.zeros(1, dtype=torch.float64, device=a.device)
    for sl in owned_mask_or_slices:
        s += torch.dot(a[sl], b[sl])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, group=group)
    return s


class DistributedDeviceBackend:
    """Vector backend of solvers.py for row-distributed device vectors (local layout: owned entries first inside each
    block, ghost slots after them).  Element-wise operations act on the whole local vector (ghost slots hold stale
    values that no owned result depends on); inner products run over the owned entries and are all-reduced inside
    libdcp (dcp_vec_dot_allreduce) -- the MPI_Allreduce of the reference's l2_norm / operator*
    (boussinesq_model.tpp:1165, 1427; inverse_matrix.hpp:99).

    owned_by_length: {local vector length: [(begin, end), ...]} -- which entries of a vector of that length are
    owned (the full block vector and every block sub-vector a solver chain takes dots of)."""

    supports_mgs = False      # inner products need the owned ranges and the all-reduce: the Arnoldi loop stays in solvers.py
    resident_cg = False

    def __init__(self, ctx, comm, owned_by_length):
        from . import solvers
        self._b = solvers.DeviceBackend(ctx)
        self.comm, self.owned = comm, dict(owned_by_length)

    def dot(self, x, y):
        return self.comm.dot(x, y, self.owned[x.numel()])

    def __getattr__(self, name):          # zeros, copy, assign, axpy, sadd, scale, zero, add_scalar, to_numpy, from_numpy
        return getattr(self._b, name)


class HaloBlock:
    """block(i,j).vmult of a row-distributed block matrix: ghost refresh of the source block, then the owned rows."""

    def __init__(self, matrix_block, halo_of_source_block):
        self.block, self.halo = matrix_block, halo_of_source_block

    def vmult(self, dst, src):
        self.halo.exchange(src)
        self.block.vmult(dst, src)

    def vmult_add(self, dst, src):
        self.halo.exchange(src)
        self.block.vmult_add(dst, src)
