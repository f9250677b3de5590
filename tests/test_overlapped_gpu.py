"""GPU tests of the overlapped host copies (B200_Vector_set_dense / export_dense with where = 2, GrB_Vector_wait) around GrB_mxv --
the path bench.py's end-to-end leg takes through the reference's Matrix.mxv.

test_pipelined_exports_are_bit_exact is the regression test of a write-after-read hazard on the TMA stage of the hot-table SpMV
kernel (spmv_run.cuh): the bulk copy of a warp's next run could land before a lane's shared-memory loads of the current run had
returned, folding words of the next run into this run's rows -- one wrong hub-row sum in about 1 % of the launches at scale 22,
and only when copies kept the memory system busy (tools/e2e_stress.py, profiles/r02_e2e_stress_*.txt)."""
import functools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pinned(n, dtype):
    import torch
    return torch.empty(n, dtype=dtype).pin_memory().numpy()


@functools.lru_cache(maxsize=2)
def _setup(scale, edgefactor=16):
    import torch
    import pygraphblas_b200 as gb
    from pygraphblas_b200 import Matrix, FP32
    from pygraphblas_b200.generators import rmat_csr
    if not gb.have_device():
        pytest.skip("no CUDA device")
    n, indptr, indices = rmat_csr(scale, edgefactor, seed=11)
    rng = np.random.default_rng(4)
    vals = (rng.random(len(indices), dtype=np.float32) + np.float32(0.5)).astype(np.float32)
    u0 = rng.random(n, dtype=np.float32)
    return gb, torch, Matrix.from_csr(indptr, indices, vals, n, n, FP32), n, u0


def test_overlapped_import_export_round_trip():
    gb, torch, A, n, u0 = _setup(12)
    from pygraphblas_b200 import Vector, FP32
    lib, ffi = gb.lib, gb.ffi
    src, dst, pres = _pinned(n, torch.float32), _pinned(n, torch.float32), _pinned(n, torch.uint8)
    src[:] = u0
    v = Vector.from_numpy(np.zeros(n, np.float32))
    assert lib.B200_Vector_set_dense(v._vector[0], ffi.cast("void*", src.ctypes.data), ffi.NULL, 2) == 0
    assert lib.B200_Vector_export_dense(v._vector[0], ffi.cast("void*", dst.ctypes.data), ffi.cast("uint8_t*", pres.ctypes.data), 2) == 0
    assert lib.GrB_Vector_wait(v._vector) == 0
    assert np.array_equal(dst, u0) and np.all(pres == 1)
    # a product between the two copies, against the blocking form of the same calls
    w = Vector.sparse(FP32, n)
    A.mxv(v, semiring=FP32.PLUS_TIMES, out=w)
    assert lib.B200_Vector_export_dense(w._vector[0], ffi.cast("void*", dst.ctypes.data), ffi.cast("uint8_t*", pres.ctypes.data), 2) == 0
    assert lib.GrB_Vector_wait(w._vector) == 0
    x, p = w.to_numpy()
    assert np.array_equal(p, pres) and np.array_equal(x[p != 0], dst[p != 0])


@pytest.mark.parametrize("inplace", ["1", "0"])
def test_pipelined_exports_are_bit_exact(inplace, monkeypatch):
    """Three steps in flight, every step's u scaled by a different power of two (PLUS_TIMES is linear and the scaling exact), the
    last three exports of every burst compared bit for bit with the serial result: a stale, half-written or mis-folded export shows."""
    gb, torch, A, n, u0 = _setup(20)
    from pygraphblas_b200 import Vector, FP32
    lib, ffi = gb.lib, gb.ffi
    monkeypatch.setenv("B200GRB_MXV_INPLACE", inplace)
    lib.B200_reload_tunables()
    try:
        NB, NF, bursts, steps = 3, 4, 36, 11
        u_pin = [_pinned(n, torch.float32) for _ in range(NF)]
        for f in range(NF):
            u_pin[f][:] = u0 * np.float32(2.0 ** f)
        w_pin = [_pinned(n, torch.float32) for _ in range(NB)]
        p_pin = [_pinned(n, torch.uint8) for _ in range(NB)]
        up = [ffi.cast("void*", x.ctypes.data) for x in u_pin]
        wp = [ffi.cast("void*", x.ctypes.data) for x in w_pin]
        pp = [ffi.cast("uint8_t*", x.ctypes.data) for x in p_pin]
        us, ws = Vector.from_numpy(u0), Vector.sparse(FP32, n)
        A.mxv(us, semiring=FP32.PLUS_TIMES, out=ws)
        w_ref, p_ref = ws.to_numpy()
        w_ref, p_ref = w_ref.copy(), p_ref.copy()
        keep = p_ref != 0
        ue = [Vector.from_numpy(u_pin[0]) for _ in range(NB)]
        we = [Vector.sparse(FP32, n) for _ in range(NB)]
        bad, s = [], 0
        for _ in range(bursts):
            lib.B200_device_synchronize()
            for i in range(steps):
                b, f = s % NB, s % NF
                assert lib.B200_Vector_set_dense(ue[b]._vector[0], up[f], ffi.NULL, 2) == 0
                A.mxv(ue[b], semiring=FP32.PLUS_TIMES, out=we[b])
                assert lib.B200_Vector_export_dense(we[b]._vector[0], wp[b], pp[b], 2) == 0
                if i >= 2:
                    assert lib.GrB_Vector_wait(we[(s - 2) % NB]._vector) == 0
                s += 1
            lib.B200_device_synchronize()
            for t in range(s - NB, s):
                b, f = t % NB, t % NF
                want = w_ref * np.float32(2.0 ** f)
                if not (np.array_equal(p_pin[b], p_ref) and np.array_equal(w_pin[b][keep], want[keep])):
                    bad.append((t, int(np.count_nonzero(p_pin[b] != p_ref)), np.flatnonzero(keep & (w_pin[b] != want))[:4].tolist()))
        assert not bad, f"{len(bad)} of {bursts * NB} exports differ from the serial result: {bad[:4]}"
    finally:
        monkeypatch.delenv("B200GRB_MXV_INPLACE", raising=False)
        lib.B200_reload_tunables()
