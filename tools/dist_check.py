#!/usr/bin/env python
"""Multi-GPU exchange check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py [scale]

1. all-gather form (mxv): rank r owns a row block of A; w = A u is formed by local GrB_mxv + the library's
   peer push; the replicated result on EVERY rank must equal the single-process product (presence bit-exact, values
   bit-exact for MIN_PLUS, <= 1e-6 relative for FP32 PLUS_TIMES), over many back-to-back steps (double buffering).
2. all-reduce form (vxm / T0 on the row-split A): every rank folds its rows' contributions into a full-length partial;
   the rank-ordered monoid fold over peer memory must equal the single-process product.
3. device time of both collectives on a scale-22-sized FP32 vector.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("B200GRB_DEVICE", str(local))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pygraphblas_b200 as gb
    from pygraphblas_b200 import Matrix, Vector, FP32, descriptor
    from pygraphblas_b200.generators import rmat_csr
    from pygraphblas_b200.distributed import Comm, equal_row_blocks
    import scipy.sparse as sp
    lib, ffi = gb.lib, gb.ffi
    sp_ = ffi.new("void**"); lib.B200_get_stream(sp_)
    stream = torch.cuda.ExternalStream(int(ffi.cast("uintptr_t", sp_[0])), device=torch.device("cuda", local))

    n, indptr, indices = rmat_csr(scale, 16, seed=1)
    rng = np.random.default_rng(2)
    vals = (rng.random(len(indices), dtype=np.float32) + np.float32(0.5)).astype(np.float32)
    u0 = rng.random(n, dtype=np.float32)
    S = sp.csr_matrix((vals, indices, indptr), shape=(n, n))
    bounds = equal_row_blocks(n, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    Sl = S[r0:r1]
    A = Matrix.from_csr(Sl.indptr.astype(np.int64), Sl.indices.astype(np.uint32), Sl.data, r1 - r0, n, FP32)
    comm = Comm(n, FP32, rank, world)
    ok = True

    # ---- 1. all-gather form
    rows_nonempty = np.diff(indptr) > 0
    for name, sr, exact in (("PLUS_TIMES", FP32.PLUS_TIMES, False), ("MIN_PLUS", FP32.MIN_PLUS, True)):
        for k in range(6):
            uk = (u0 * np.float32(1 + k)).astype(np.float32)
            u = Vector.from_numpy(uk)
            wl = A.mxv(u, semiring=sr)
            w = comm.allgather(wl, r0)
            x, p = w.to_numpy()
            if name == "PLUS_TIMES":
                ref = (S.astype(np.float64) @ uk.astype(np.float64))
                good = np.array_equal(p != 0, rows_nonempty) and np.allclose(x[rows_nonempty], ref[rows_nonempty], rtol=1e-6 * 30, atol=0)
                err = float(np.max(np.abs(x[rows_nonempty] - ref[rows_nonempty]) / np.abs(ref[rows_nonempty])))
            else:
                # single-process product through the same library on this rank (bit-exact: min is order independent)
                Af = Matrix.from_csr(indptr, indices, vals, n, n, FP32)
                xs, ps = Af.mxv(u, semiring=sr).to_numpy()
                good = np.array_equal(p, ps) and np.array_equal(x[ps != 0], xs[ps != 0])
                err = 0.0
            ok &= bool(good)
            if rank == 0:
                print(f"allgather {name} step {k}: {'ok' if good else 'MISMATCH'} (max rel err {err:.2e})", flush=True)
    # many back-to-back steps without host synchronisation: the last result must still be right
    u = Vector.from_numpy(u0)
    for k in range(200):
        w = comm.allgather(A.mxv(u, semiring=FP32.PLUS_TIMES), r0)
    x, p = w.to_numpy()
    ref = S.astype(np.float64) @ u0.astype(np.float64)
    good = np.array_equal(p != 0, rows_nonempty) and np.allclose(x[rows_nonempty], ref[rows_nonempty], rtol=3e-5, atol=0)
    ok &= bool(good)
    if rank == 0:
        print(f"allgather 200 back-to-back steps: {'ok' if good else 'MISMATCH'}", flush=True)

    # ---- 2. all-reduce form: w = A' u, rank r contributes A[r0:r1, :]' u[r0:r1]
    for name, sr, mon in (("PLUS_TIMES", FP32.PLUS_TIMES, FP32.PLUS_MONOID), ("MIN_PLUS", FP32.MIN_PLUS, FP32.MIN_MONOID)):
        ul = Vector.from_numpy(u0[r0:r1])
        part = A.mxv(ul, semiring=sr, desc=descriptor.T0)            # full-length partial (n)
        w = comm.allreduce(part, mon)
        x, p = w.to_numpy()
        Af = Matrix.from_csr(indptr, indices, vals, n, n, FP32)
        xs, ps = Af.mxv(Vector.from_numpy(u0), semiring=sr, desc=descriptor.T0).to_numpy()
        if name == "MIN_PLUS":
            good = np.array_equal(p, ps) and np.array_equal(x[ps != 0], xs[ps != 0])
        else:
            good = np.array_equal(p, ps) and np.allclose(x[ps != 0], xs[ps != 0], rtol=3e-5, atol=0)
        ok &= bool(good)
        if rank == 0:
            print(f"allreduce {name}: {'ok' if good else 'MISMATCH'}", flush=True)

    # ---- 3. device time of the collectives on a scale-22 sized vector
    n22 = 1 << 22
    comm2 = Comm(n22, FP32, rank, world)
    b2 = equal_row_blocks(n22, world)
    sl = Vector.from_numpy(np.ones(b2[rank + 1] - b2[rank], np.float32))
    full = Vector.from_numpy(np.ones(n22, np.float32))
    for label, fn in (("allgather", lambda: comm2.allgather(sl, b2[rank])), ("allreduce", lambda: comm2.allreduce(full, FP32.PLUS_MONOID))):
        for _ in range(5):
            fn()
        lib.B200_device_synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50):
            fn()
        e1.record(stream)
        lib.B200_device_synchronize(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{label} of a 2^22 FP32 vector over {world} GPUs: {float(t.item()) * 1e3:.1f} us per call (max over ranks)", flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "OK" if int(t.item()) == 1 else "FAILED", flush=True)
    comm.close(); comm2.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
