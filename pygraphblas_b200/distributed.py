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
    old_of_new = np.empty(n, np.int64)
    old_of_new[newid] = np.arange(n)
    lo, hi = rank * lb, min((rank + 1) * lb, n)
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
