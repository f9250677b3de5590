#!/usr/bin/env python
"""Run the masked triangle-count SpGEMM (BASELINE.json configs[3]) a few times -- a short target for ncu.
    python tools/prof_spgemm.py [scale] [reps] [unmasked]"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp
import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, INT64, FP32, descriptor
from bench import cached_graph

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "masked"
n, indptr, indices = cached_graph(scale)
if mode.startswith("masked"):
    S = sp.csr_matrix((np.ones(len(indices), np.int8), indices, indptr), shape=(n, n))
    Ls = sp.tril(S + S.T, -1).tocsr(); Ls.sort_indices()
    L = Matrix.from_csr(Ls.indptr.astype(np.int64), Ls.indices.astype(np.uint32), np.ones(Ls.nnz, np.int64), n, n, INT64)
    for dname in (("S",) if mode == "masked_S" else ("S", "ST1")):
        for _ in range(reps):
            gb.lib.B200_device_synchronize(); t0 = time.perf_counter()
            C = L.mxm(L, mask=L, semiring=INT64.PLUS_PAIR, desc=getattr(descriptor, dname))
            gb.lib.B200_device_synchronize(); dt = time.perf_counter() - t0
            print(f"masked {dname} scale {scale}: {dt*1e3:.2f} ms wall (incl. freeing the previous result), nnz_out {C.nvals}", flush=True)
else:
    vals = np.ones(len(indices), np.float32)
    A = Matrix.from_csr(indptr, indices, vals, n, n, FP32)
    for _ in range(reps):
        gb.lib.B200_device_synchronize(); t0 = time.perf_counter()
        C = A.mxm(A, semiring=FP32.PLUS_SECOND)
        gb.lib.B200_device_synchronize(); dt = time.perf_counter() - t0
        f = gb.ffi.new("uint64_t*"); o = gb.ffi.new("uint64_t*"); gb.lib.B200_last_mxm_stats(f, o)
        print(f"unmasked scale {scale}: {dt*1e3:.2f} ms, products {f[0]}, nnz_out {o[0]}, {o[0]/dt/1e6:.1f} Mnnz/s", flush=True)
