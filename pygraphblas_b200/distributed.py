"""1-D row-block partitioning for the multi-GPU hot path.

One process per GPU.  Rank r owns the contiguous row block [bounds[r], bounds[r+1]) of A, cut so
that every block holds about the same number of entries (R-MAT hubs sit at low row ids, so equal-row
splits are badly imbalanced).  `mxv` is local to the block; ONE all-gather of the output slices per
operation gives every rank the full vector for the next one.  The collective runs through
torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np

from .generators import row_block_bounds


def local_block(indptr, indices, values, world, rank):
    """(bounds, local_indptr, local_indices, local_values) of rank's row block."""
    bounds = row_block_bounds(indptr, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    k0, k1 = int(indptr[r0]), int(indptr[r1])
    lptr = (np.asarray(indptr[r0:r1 + 1]) - k0).astype(np.int64)
    return bounds, lptr, indices[k0:k1], (None if values is None else values[k0:k1])


def scatter_ids(n, seed=0x9E3779B1):
    """A fixed pseudo-random relabelling of the n vertices (new id of vertex v).  R-MAT puts its hubs at
    low ids, so contiguous blocks of the ORIGINAL ids are either row-balanced or nnz-balanced, never both;
    contiguous blocks of the scattered ids are both, which lets the blocks have equal length (one plain
    all-gather, no padding) and equal work."""
    rng = np.random.default_rng(seed)
    return rng.permutation(n).astype(np.int64)


def local_block_scattered(indptr, indices, values, world, rank):
    """Rank's block of P A P' (P = scatter_ids): rows with new id in [rank*lb, (rank+1)*lb), lb = ceil(n/world),
    columns relabelled to new ids.  Returns (newid, lb, local_indptr, local_indices, local_values); the input
    vector must be given in the new order (x_new[newid] = x) and the concatenation of the ranks' results IS
    the output vector in the new order."""
    import scipy.sparse as sp
    n = len(indptr) - 1
    newid = scatter_ids(n)
    lb = -(-n // world)
    lb = -(-lb // 16) * 16                 # blocks start at multiples of 16 rows: the library moves slices as 16-byte words
    old_of_new = np.empty(n, np.int64)
    old_of_new[newid] = np.arange(n)
    lo, hi = min(rank * lb, n), min((rank + 1) * lb, n)
    rows = old_of_new[lo:hi]
    data = np.ones(len(indices), np.float32) if values is None else values
    S = sp.csr_matrix((data, indices, indptr), shape=(n, n))[rows]
    S = sp.csr_matrix((S.data, newid[S.indices].astype(np.int32), S.indptr), shape=(len(rows), world * lb))
    S.sort_indices()
    lptr = np.empty(lb + 1, np.int64)
    lptr[:len(rows) + 1] = S.indptr
    lptr[len(rows) + 1:] = S.indptr[-1]
    return newid, lb, lptr, S.indices.astype(np.uint32), (None if values is None else S.data)


def to_scattered(x, newid, total):
    out = np.zeros(total, dtype=x.dtype)
    out[newid] = x
    return out


def allgather_slices(full, local, bounds):
    """All-gather the per-rank output slices `local` (uneven lengths) into the 1-D tensor `full`
    (length bounds[-1]) on every rank.  `full` and `local` are torch tensors on the same device."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    views = [full[int(bounds[r]):int(bounds[r + 1])] for r in range(world)]
    sizes = {v.numel() for v in views}
    if len(sizes) == 1:
        dist.all_gather(views, local)
        return full
    # uneven blocks (the normal case for nnz-balanced splits): every rank sends its slice to every
    # peer and receives theirs, batched into ONE grouped NCCL operation (all-gather-v)
    views[rank].copy_(local)
    ops = []
    for r in range(world):
        if r == rank:
            continue
        if local.numel():
            ops.append(dist.P2POp(dist.isend, local, r))
        if views[r].numel():
            ops.append(dist.P2POp(dist.irecv, views[r], r))
    for w in dist.batch_isend_irecv(ops) if ops else []:
        w.wait()
    return full


# ---------------------------------------------------------------------------------------------------------------
# The exchange step inside the library (csrc/dist.cu): peers write each other's HBM over NVLink with the library's
# own kernels; this module only carries the 64-byte CUDA-IPC handles between the processes (control plane).
def equal_row_blocks(n, world, align=16):
    """Row offsets of `world` contiguous blocks of (almost) equal length, every block starting at a multiple of
    `align` rows (the library moves slices as 16-byte words)."""
    per = -(-n // world)
    per = -(-per // align) * align
    return [min(n, r * per) for r in range(world + 1)]


def exchange_handles(handle, world, rank, gather=None):
    """All ranks' 64-byte handles, in rank order, as one bytes object.  `gather(list_out, obj)` defaults to
    torch.distributed.all_gather_object (works on gloo and nccl process groups alike)."""
    if world == 1:
        return bytes(handle)
    if gather is None:
        import torch.distributed as dist
        gather = dist.all_gather_object
    out = [None] * world
    gather(out, bytes(handle))
    assert all(isinstance(h, (bytes, bytearray)) and len(h) == 64 for h in out), "IPC handles are 64 bytes"
    return b"".join(bytes(h) for h in out)


class Comm:
    """A communicator for replicated vectors of length n (B200_Comm_* of include/b200grb.h).

        comm = Comm(n, FP32)                         # collective: every rank of the process group
        w_slice = A_local.mxv(u, semiring=...)       # local rows
        w = comm.allgather(w_slice, row0)            # Vector view of the replicated result (borrowed)
        p = comm.allreduce(partial, FP32.PLUS_MONOID)
    """

    def __init__(self, n, typ, rank=None, world=None, gather=None):
        from .base import lib, ffi, _check
        from .vector import Vector
        if rank is None or world is None:
            import torch.distributed as dist
            rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        self._lib, self._ffi, self._check, self._Vector = lib, ffi, _check, Vector
        self.rank, self.world, self.n, self.type = rank, world, n, typ
        self._comm = ffi.new("B200_Comm*")
        _check(lib.B200_Comm_create(self._comm, rank, world, n, typ.gb_type))
        h = ffi.new("unsigned char[64]")
        _check(lib.B200_Comm_handle(self._comm[0], h))
        allh = exchange_handles(bytes(ffi.buffer(h, 64)), world, rank, gather)
        _check(lib.B200_Comm_connect(self._comm[0], ffi.from_buffer(allh)))
        self.bounds = equal_row_blocks(n, world)

    def _view(self):
        v = self._ffi.new("GrB_Vector*")
        self._check(self._lib.B200_Comm_result(self._comm[0], v))
        return self._Vector(v, owner=self)      # the communicator owns the handle: never freed through the view

    def allgather(self, slice_vec, row0=None):
        row0 = self.bounds[self.rank] if row0 is None else row0
        self._check(self._lib.B200_Comm_allgather(self._comm[0], slice_vec._vector[0], row0))
        return self._view()

    def allreduce(self, partial, monoid):
        self._check(self._lib.B200_Comm_allreduce(self._comm[0], partial._vector[0], getattr(monoid, "monoid", monoid)))
        return self._view()

    def barrier(self):
        self._check(self._lib.B200_Comm_barrier(self._comm[0]))

    def close(self):
        if self._comm is not None:
            self._lib.B200_Comm_free(self._comm)
            self._comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
