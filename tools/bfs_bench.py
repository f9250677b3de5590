#!/usr/bin/env python
"""BASELINE.json configs[2] shape on one GPU: full BFS from the max-out-degree vertex with
`q<!visited, replace> = q' lor.land A` per level (Vector.vxm, descriptor RC; the frontier update
visited |= q is an accumulating mxv with the identity matrix, i.e. the hot path again).
Reports per-level times, the heaviest step in GEdge/s (edges of the graph / step time) and checks the
reached set against scipy.  python tools/bfs_bench.py [scale]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp
from scipy.sparse.csgraph import breadth_first_order
import torch
import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, Vector, BOOL, descriptor
from bench import cached_graph

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n, indptr, indices = cached_graph(scale)
nnz = len(indices)
A = Matrix.from_csr(indptr, indices, None, n, n, BOOL)
I = Matrix.from_csr(np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.uint32), None, n, n, BOOL)
src = int(np.argmax(np.diff(indptr)))
sp_ = gb.ffi.new("void**"); gb.lib.B200_get_stream(sp_)
stream = torch.cuda.ExternalStream(int(gb.ffi.cast("uintptr_t", sp_[0])))

def bfs():
    q = Vector.sparse(BOOL, n); q[src] = True
    visited = Vector.sparse(BOOL, n); visited[src] = True
    times, sizes = [], []
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        q = q.vxm(A, mask=visited, desc=descriptor.RC, semiring=BOOL.LOR_LAND)
        e1.record(stream)
        nq = q.nvals
        gb.lib.B200_device_synchronize(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1)); sizes.append(nq)
        if nq == 0:
            break
        I.mxv(q, out=visited, accum=BOOL.LOR, semiring=BOOL.LOR_LAND)
    return visited, times, sizes

bfs()                                   # warm-up: builds the cached transpose
t0 = time.perf_counter(); visited, times, sizes = bfs(); wall = time.perf_counter() - t0
reached = visited.nvals
order = breadth_first_order(sp.csr_matrix((np.ones(nnz, np.bool_), indices, indptr), shape=(n, n)), src, directed=True, return_predecessors=False)
print(f"scale {scale}: n={n} nnz={nnz} src={src} levels={len(sizes)-1} reached={reached} scipy_reached={len(order)} parity={reached == len(order)}")
print("frontier sizes:", sizes)
print("step ms:", [round(t, 3) for t in times])
tmax = max(times)
print(f"heaviest step {tmax:.3f} ms = {nnz / tmax / 1e6:.1f} GEdge/s (graph edges / step time); whole BFS {sum(times):.2f} ms device, {wall*1e3:.1f} ms wall -> {nnz / sum(times) / 1e6:.1f} GTEPS-equivalent")
