#!/usr/bin/env python
"""Stress of the overlapped host copies (B200_Vector_set_dense / export_dense with where = 2) around GrB_mxv, the way bench.py's
pipelined e2e leg drives them: bursts of steps with three buffers in flight, every step's input scaled by a different power of
two so that a stale or half-written export is visible (PLUS_TIMES is linear; scaling by 2^k is exact in fp32), the last three
exports of every burst compared bit for bit with the serial result.
    python tools/e2e_stress.py [scale] [bursts] [steps_per_burst]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, Vector, FP32
from bench import cached_graph, spmv_inputs

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
bursts = int(sys.argv[2]) if len(sys.argv) > 2 else 60
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 11
lib, ffi = gb.lib, gb.ffi
n, indptr, indices = cached_graph(scale)
vals, u0 = spmv_inputs(len(indices), n)
A = Matrix.from_csr(indptr, indices, vals, n, n, FP32)
NB, NF = 3, 4
pin = lambda k, dt: torch.empty(k, dtype=dt).pin_memory().numpy()
u_pin = [pin(n, torch.float32) for _ in range(NF)]
for f in range(NF):
    u_pin[f][:] = u0 * np.float32(2.0 ** f)
w_pin = [pin(n, torch.float32) for _ in range(NB)]
p_pin = [pin(n, torch.uint8) for _ in range(NB)]
up = [ffi.cast("void*", x.ctypes.data) for x in u_pin]
wp = [ffi.cast("void*", x.ctypes.data) for x in w_pin]
pp = [ffi.cast("uint8_t*", x.ctypes.data) for x in p_pin]

# the serial result (blocking copies)
us, ws = Vector.from_numpy(u0), Vector.sparse(FP32, n)
A.mxv(us, semiring=FP32.PLUS_TIMES, out=ws)
w_ref, p_ref = ws.to_numpy()
w_ref, p_ref = w_ref.copy(), p_ref.copy()
keep = p_ref != 0
# run-to-run determinism of the kernels themselves (no copies in flight): 40 more calls, bit for bit
nd = 0
for _ in range(40):
    A.mxv(us, semiring=FP32.PLUS_TIMES, out=ws)
    w2, p2 = ws.to_numpy()
    nd += int(not (np.array_equal(p2, p_ref) and np.array_equal(w2[keep], w_ref[keep])))
print(f"serial repeats differing from the first call: {nd} of 40", flush=True)


def run(label):
    ue = [Vector.from_numpy(u_pin[0]) for _ in range(NB)]
    we = [Vector.sparse(FP32, n) for _ in range(NB)]
    bad, step_no = [], 0
    for burst in range(bursts):
        lib.B200_device_synchronize()
        first = step_no
        for i in range(steps):
            b, f = step_no % NB, step_no % NF
            assert lib.B200_Vector_set_dense(ue[b]._vector[0], up[f], ffi.NULL, 2) == 0
            A.mxv(ue[b], semiring=FP32.PLUS_TIMES, out=we[b])
            assert lib.B200_Vector_export_dense(we[b]._vector[0], wp[b], pp[b], 2) == 0
            if i >= 2:
                assert lib.GrB_Vector_wait(we[(step_no - 2) % NB]._vector) == 0
            step_no += 1
        lib.B200_device_synchronize(); torch.cuda.synchronize()
        for s in range(step_no - NB, step_no):                # the last three exports of the burst
            b, f = s % NB, s % NF
            want = w_ref * np.float32(2.0 ** f)
            pd = int(np.count_nonzero(p_pin[b] != p_ref))
            vd = np.flatnonzero(keep & (w_pin[b] != want))
            if pd or len(vd):
                got = w_pin[b][vd]
                other = {g: int(np.count_nonzero(got == w_ref[vd] * np.float32(2.0 ** g))) for g in range(NF)}
                bad.append({"burst": burst, "step": s, "step_in_burst": s - first, "buffer": b, "factor": f, "presence_diffs": pd, "value_diffs": int(len(vd)),
                            "zeros": int(np.count_nonzero(got == 0)), "equal_to_factor": other, "first": [int(x) for x in vd[:3]], "last": [int(x) for x in vd[-3:]]})
    print(f"{label}: {bursts} bursts x {steps} steps, {len(bad)} bad exports of {bursts * NB} checked", flush=True)
    for x in bad[:6]:
        print("   ", x, flush=True)
    return len(bad)


total = 0
for label, env in (("hot-table kernel (TMA-staged runs), T in w's buffers", {"B200GRB_SPMV_HOT": "-1", "B200GRB_MXV_INPLACE": "1"}),
                   ("plain run kernel (no TMA), T in w's buffers", {"B200GRB_SPMV_HOT": "0", "B200GRB_MXV_INPLACE": "1"}),
                   ("hot-table kernel, T in fresh buffers", {"B200GRB_SPMV_HOT": "-1", "B200GRB_MXV_INPLACE": "0"})):
    os.environ.update(env); lib.B200_reload_tunables()
    total += run(label)
print("E2E_STRESS", "clean" if total == 0 else f"{total} bad exports")
