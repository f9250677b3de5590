#!/usr/bin/env python
"""Sweep the SpMV hot-table configuration on the BASELINE configs[1] graph (one process, one graph build)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, Vector, FP32
from bench import cached_graph, spmv_inputs

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n, indptr, indices = cached_graph(scale)
vals, u0 = spmv_inputs(scale, len(indices), n)
A = Matrix.from_csr(indptr, indices, vals, n, n, FP32)
u = Vector.from_numpy(u0)
w = Vector.sparse(FP32, n)
sp_ = gb.ffi.new("void**"); gb.lib.B200_get_stream(sp_)
stream = torch.cuda.ExternalStream(int(gb.ffi.cast("uintptr_t", sp_[0])))

def run(label, env):
    for k in ("B200GRB_SPMV_ITEMS",):
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(5):
        A.mxv(u, semiring=FP32.PLUS_TIMES, out=w)
    gb.lib.B200_device_synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(50):
        A.mxv(u, semiring=FP32.PLUS_TIMES, out=w)
    e1.record(stream)
    gb.lib.B200_device_synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"{label:32s} {ms*1e3:8.1f} us  {len(indices)/ms/1e6:7.1f} GEdge/s", flush=True)

run("plain tile kernel", {"B200GRB_NO_HOT": "1"})
run("plain kernel, relabelled columns", {"B200GRB_RELABEL_ONLY": "1"})
for groups in (4, 2):
    for kb in (16, 32, 64, 128, 160):
        run(f"hot groups={groups} table={kb}KB", {"B200GRB_HOT_GROUPS": str(groups), "B200GRB_HOT_KB": str(kb)})
