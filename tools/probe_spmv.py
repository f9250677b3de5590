#!/usr/bin/env python
"""Which part of the SpMV tile kernel costs what: same graph, semirings that drop the gather / the value stream."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, Vector, FP32
from bench import cached_graph, spmv_inputs

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n, indptr, indices = cached_graph(scale)
vals, u0 = spmv_inputs(len(indices), n)
A = Matrix.from_csr(indptr, indices, vals, n, n, FP32)
u = Vector.from_numpy(u0)
w = Vector.sparse(FP32, n)
sp_ = gb.ffi.new("void**"); gb.lib.B200_get_stream(sp_)
stream = torch.cuda.ExternalStream(int(gb.ffi.cast("uintptr_t", sp_[0])))

def run(label, sr, env):
    keep_run = os.environ.get("B200GRB_SPMV_RUN")
    for k in ("B200GRB_SPMV_ITEMS", "B200GRB_SPMV_HOT", "B200GRB_HOT_GROUPS", "B200GRB_SPMV_DEBUG", "B200GRB_SPMV_RUN", "B200GRB_SPMV_PIPE"):
        os.environ.pop(k, None)
    if keep_run is not None and "B200GRB_SPMV_RUN" not in env:
        os.environ["B200GRB_SPMV_RUN"] = keep_run
    os.environ.update(env)
    gb.lib.B200_reload_tunables()
    for _ in range(5):
        A.mxv(u, semiring=sr, out=w)
    gb.lib.B200_device_synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(30):
        A.mxv(u, semiring=sr, out=w)
    e1.record(stream)
    gb.lib.B200_device_synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"{label:44s} {ms*1e3:8.1f} us", flush=True)

if len(sys.argv) > 2 and sys.argv[2] == "sweep":
    for rep in range(3):
        for pipe in ("0", "1"):
            for kb in ("64", "96", "128"):
                run(f"rep {rep} pipe={pipe} table cap {kb}KB PLUS_TIMES", FP32.PLUS_TIMES, {"B200GRB_SPMV_HOT": kb, "B200GRB_SPMV_PIPE": pipe})
    run("plain run kernel (no table, no staging)", FP32.PLUS_TIMES, {"B200GRB_SPMV_HOT": "0"})
    run("default MIN_PLUS", FP32.MIN_PLUS, {})
    run("default PLUS_SECOND", FP32.PLUS_SECOND, {})
    run("PLUS_FIRST (no gather)", FP32.PLUS_FIRST, {})
    sys.exit(0)
run("default (hot table, TMA staged) PLUS_TIMES", FP32.PLUS_TIMES, {})
run("run kernel PLUS_SECOND", FP32.PLUS_SECOND, {})
run("run kernel PLUS_FIRST (no gather)", FP32.PLUS_FIRST, {})
run("run kernel PLUS_PAIR (col only)", FP32.PLUS_PAIR, {})
run("run kernel MIN_PLUS", FP32.MIN_PLUS, {})
for kb in ("0", "32", "64", "96", "128", "160"):
    run(f"run kernel + hot table {kb}KB PLUS_TIMES", FP32.PLUS_TIMES, {"B200GRB_SPMV_HOT": kb})
run("run-time operators: PLUS_MINUS", FP32.PLUS_MINUS, {})
