"""world_size-2 CPU (gloo) test of the multi-GPU host logic: nnz-balanced row blocks + one all-gather of
the output slices reproduce the single-process SpMV.  The per-rank arithmetic here is the CPU oracle's
OpenMP port (the CUDA kernels need a GPU); what is under test is the partition / gather plumbing that
bench.py uses unchanged with NCCL."""
import ctypes
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from pygraphblas_b200.distributed import local_block, allgather_slices
    from pygraphblas_b200.generators import rmat_csr
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, indptr, indices = rmat_csr(11, 8, seed=3)
    rng = np.random.default_rng(0)
    vals = (rng.integers(1, 9, len(indices)) / 4.0).astype(np.float32)
    u = (rng.integers(0, 9, n) / 4.0).astype(np.float32)
    bounds, lptr, lidx, lval = local_block(indptr, indices, vals, world, rank)
    lrows = len(lptr) - 1
    L = orc.lib()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    full = torch.from_numpy(u.copy())
    for _ in range(3):                       # three chained steps: the gathered w is the next u
        w = np.zeros(lrows, np.float32); pres = np.zeros(lrows, np.uint8)
        uu = full.numpy()
        L.fast_spmv_plus_times_f32(ctypes.c_int64(lrows), p(lptr), p(np.ascontiguousarray(lidx)), p(np.ascontiguousarray(lval)), p(uu), p(w), p(pres))
        allgather_slices(full, torch.from_numpy(w), bounds)
    if rank == 0:
        np.save(out, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _worker_scattered(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from pygraphblas_b200.distributed import local_block_scattered, to_scattered
    from pygraphblas_b200.generators import rmat_csr
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, indptr, indices = rmat_csr(11, 8, seed=3)
    rng = np.random.default_rng(0)
    vals = (rng.integers(1, 9, len(indices)) / 4.0).astype(np.float32)
    u = (rng.integers(0, 9, n) / 4.0).astype(np.float32)
    newid, lb, lptr, lidx, lval = local_block_scattered(indptr, indices, vals, world, rank)
    full = torch.from_numpy(to_scattered(u, newid, world * lb))
    L = orc.lib()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for _ in range(3):
        w = np.zeros(lb, np.float32); pres = np.zeros(lb, np.uint8)
        uu = full.numpy().copy()
        L.fast_spmv_plus_times_f32(ctypes.c_int64(lb), p(lptr), p(np.ascontiguousarray(lidx)), p(np.ascontiguousarray(lval)), p(uu), p(w), p(pres))
        dist.all_gather_into_tensor(full, torch.from_numpy(w))        # the call bench.py issues with NCCL
    if rank == 0:
        np.save(out, full.numpy()[newid])                             # back to the original vertex order
    dist.barrier()
    dist.destroy_process_group()


def test_scattered_equal_blocks_world2(tmp_path):
    """bench.py's multi-GPU layout: pseudo-random vertex relabelling -> equal-length, equal-work blocks ->
    one all_gather_into_tensor per step."""
    import torch.multiprocessing as mp
    from pygraphblas_b200.generators import rmat_csr
    from pygraphblas_b200.distributed import local_block_scattered
    import scipy.sparse as sp
    out = str(tmp_path / "w2.npy")
    mp.spawn(_worker_scattered, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    n, indptr, indices = rmat_csr(11, 8, seed=3)
    rng = np.random.default_rng(0)
    vals = (rng.integers(1, 9, len(indices)) / 4.0).astype(np.float32)
    u = (rng.integers(0, 9, n) / 4.0).astype(np.float32)
    A = sp.csr_matrix((vals.astype(np.float64), indices, indptr), shape=(n, n))
    ref = u.astype(np.float64)
    for _ in range(3):
        ref = A @ ref
    assert np.allclose(got, ref, rtol=1e-5)
    work = [len(local_block_scattered(indptr, indices, vals, 8, r)[3]) for r in range(8)]
    assert max(work) <= 1.35 * (sum(work) / 8)


def test_row_block_allgather_world2(tmp_path):
    import torch.multiprocessing as mp
    from pygraphblas_b200.generators import rmat_csr, row_block_bounds
    import scipy.sparse as sp
    out = str(tmp_path / "w.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    n, indptr, indices = rmat_csr(11, 8, seed=3)
    rng = np.random.default_rng(0)
    vals = (rng.integers(1, 9, len(indices)) / 4.0).astype(np.float32)
    u = (rng.integers(0, 9, n) / 4.0).astype(np.float32)
    A = sp.csr_matrix((vals.astype(np.float64), indices, indptr), shape=(n, n))
    ref = u.astype(np.float64)
    for _ in range(3):
        ref = A @ ref
    assert np.allclose(got, ref, rtol=1e-5)
    # the split is nnz-balanced and covers every row exactly once
    b = row_block_bounds(indptr, 8)
    assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
    per = np.diff(indptr[b])
    assert per.max() <= 1.3 * per.mean() + indptr[1:].max()


# ------------------------------------------------------------------ the in-library exchange (csrc/dist.cu): host-side control plane
def _worker_handles(rank, world, port, out):
    import torch.distributed as dist
    from pygraphblas_b200.distributed import exchange_handles
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank + 1]) * 64                       # stands for this rank's 64-byte CUDA-IPC handle
    allh = exchange_handles(mine, world, rank)
    ok = len(allh) == 64 * world and all(allh[64 * r:64 * (r + 1)] == bytes([r + 1]) * 64 for r in range(world))
    np.save(out + f".{rank}.npy", np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_ipc_handle_exchange_world2(tmp_path):
    """Every rank ends up with all ranks' 64-byte handles in rank order (what B200_Comm_connect takes)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "h")
    mp.spawn(_worker_handles, args=(2, _free_port(), out), nprocs=2, join=True)
    assert all(bool(np.load(out + f".{r}.npy")[0]) for r in range(2))


def test_equal_row_blocks_are_aligned_and_cover():
    from pygraphblas_b200.distributed import equal_row_blocks, local_block_scattered
    from pygraphblas_b200.generators import rmat_csr
    for n, world in ((1 << 22, 8), (1000, 3), (17, 2), (5, 8)):
        b = equal_row_blocks(n, world)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1 and all(x <= y for x, y in zip(b, b[1:]))
        assert all(x % 16 == 0 or x == n for x in b[:-1])        # the library moves slices as 16-byte words (empty trailing blocks start at n)
    n, indptr, indices = rmat_csr(10, 8, seed=3)
    for world in (2, 3, 8):
        parts = [local_block_scattered(indptr, indices, None, world, r) for r in range(world)]
        lb = parts[0][1]
        assert lb % 16 == 0 and world * lb >= n
        assert sum(len(p[3]) for p in parts) == len(indices)          # every entry lands in exactly one block


def test_comm_refuses_without_a_gpu():
    """The exchange runs only over GPU peer memory: without a device the communicator refuses (no CPU fallback)."""
    import pygraphblas_b200 as gb
    if gb.have_device():
        pytest.skip("needs a machine without a CUDA device")
    from pygraphblas_b200.distributed import Comm
    with pytest.raises(gb.Panic):
        Comm(1024, gb.FP32, 0, 1)
