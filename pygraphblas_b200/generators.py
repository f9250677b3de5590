"""Synthetic graph generators for the benchmark configurations.

Graph500-style R-MAT / Kronecker: a,b,c,d = 0.57,0.19,0.19,0.05, edgefactor tuples per
vertex, numpy.random.default_rng(seed), no vertex permutation, no symmetrisation,
duplicates merged, self-loops kept.
"""
import numpy as np


def _rmat_chunk(args):
    seed_seq, m, scale, a, b, c = args
    rng = np.random.default_rng(seed_seq)
    ab, abc = a + b, a + b + c
    r = np.zeros(m, np.uint32)
    cc = np.zeros(m, np.uint32)
    for _ in range(scale):
        x = rng.random(m, dtype=np.float32)
        rbit = x >= ab                               # quadrants c, d: lower half
        cbit = ((x >= a) & (x < ab)) | (x >= abc)    # quadrants b, d: right half
        r = (r << 1) | rbit.astype(np.uint32)
        cc = (cc << 1) | cbit.astype(np.uint32)
    return (r.astype(np.uint64) << np.uint64(32)) | cc.astype(np.uint64)


def rmat_keys(scale, edgefactor=16, seed=1, a=0.57, b=0.19, c=0.19, chunk=1 << 22, threads=None):
    """uint64 keys (row << 32 | col) of 2^scale * edgefactor directed tuples (with duplicates).
    The tuple stream is cut into fixed chunks, each drawn from its own child of
    SeedSequence(seed), so the result does not depend on the number of worker threads."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n_edges = (1 << scale) * edgefactor
    nchunks = (n_edges + chunk - 1) // chunk
    seeds = np.random.SeedSequence(seed).spawn(nchunks)
    jobs = [(seeds[k], min(chunk, n_edges - k * chunk), scale, a, b, c) for k in range(nchunks)]
    threads = threads or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(_rmat_chunk, jobs))
    return np.concatenate(parts) if len(parts) > 1 else parts[0]


def rmat_csr(scale, edgefactor=16, seed=1):
    """(n, indptr int64, indices uint32) of the deduplicated R-MAT pattern, columns sorted."""
    n = 1 << scale
    key = np.sort(rmat_keys(scale, edgefactor, seed))
    keep = np.empty(len(key), np.bool_)
    keep[0] = True
    np.not_equal(key[1:], key[:-1], out=keep[1:])
    key = key[keep]
    del keep
    r = (key >> np.uint64(32)).astype(np.int64)
    indices = (key & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    del key
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=n), out=indptr[1:])
    return n, indptr, indices


def row_block_bounds(indptr, nparts):
    """nnz-balanced contiguous row ranges: bounds[p]..bounds[p+1] is part p."""
    nnz = int(indptr[-1])
    n = len(indptr) - 1
    targets = (np.arange(1, nparts) * nnz) // nparts
    cuts = np.searchsorted(indptr, targets, side="left")
    bounds = np.concatenate(([0], np.minimum(cuts, n), [n])).astype(np.int64)
    return np.maximum.accumulate(bounds)
