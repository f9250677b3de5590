"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/*.h
declares, and the host-side plumbing (no compute) behaves like the reference's binding expects."""
import ctypes
import numpy as np
import pytest

import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, Vector, INT64, FP32, FP64, BOOL, UINT8, INT8, descriptor, Accum
from pygraphblas_b200._ffi import declared_symbols, LIB_PATH, ffi, lib


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) > 2000
    dll = ctypes.CDLL(LIB_PATH)
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing[:20]
    for hot in ("GrB_mxm", "GrB_mxv", "GrB_vxm"):
        assert hot in names


def test_info_codes_match_reference_numbering():
    # /root/reference/pygraphblas/base.py:189-203
    assert [lib.GrB_SUCCESS, lib.GrB_NO_VALUE, lib.GrB_UNINITIALIZED_OBJECT, lib.GrB_INVALID_OBJECT, lib.GrB_NULL_POINTER,
            lib.GrB_INVALID_VALUE, lib.GrB_INVALID_INDEX, lib.GrB_DOMAIN_MISMATCH, lib.GrB_DIMENSION_MISMATCH,
            lib.GrB_OUTPUT_NOT_EMPTY, lib.GrB_OUT_OF_MEMORY, lib.GrB_INSUFFICIENT_SPACE, lib.GrB_INDEX_OUT_OF_BOUNDS,
            lib.GrB_PANIC] == list(range(14))


def test_matrix_plumbing_roundtrip():
    m = Matrix.from_lists([2, 0, 1, 0], [0, 1, 2, 1], [3, 1, 2, 7])      # duplicate (0,1): later wins
    assert m.type is INT64 and m.shape == (3, 3) and m.nvals == 3
    assert m.to_lists() == [[0, 1, 2], [1, 2, 0], [7, 2, 3]]              # row-major sorted
    m[1, 1] = 5
    assert m[1, 1] == 5 and m.nvals == 4
    del m[1, 1]
    assert m.nvals == 3 and m.get(1, 1) is None
    with pytest.raises(gb.NoValue):
        m[2, 2]
    with pytest.raises(gb.InvalidIndex):
        m[5, 0] = 1
    d = m.dup()
    d[0, 0] = 1
    assert m.nvals == 3 and d.nvals == 4 and not m.iseq(d) and m.iseq(m.dup())
    m.clear()
    assert m.nvals == 0


def test_default_dimensions_are_hypersparse():
    m = Matrix.sparse(INT8)                                              # tests/test_matrix.py:41-44
    assert m.nrows == m.ncols == 1 << 60 and m.nvals == 0
    m[1 << 40, 1 << 50] = 3
    assert m.to_lists() == [[1 << 40], [1 << 50], [3]]
    v = Vector.sparse(FP64)
    assert v.size == 1 << 60
    v[1 << 59] = 2.5
    assert v.to_lists() == [[1 << 59], [2.5]]


def test_vector_plumbing_and_typecast():
    v = Vector.from_lists([3, 1], [1.5, -2.25], size=5)
    assert v.type is FP64 and v.size == 5 and v.to_lists() == [[1, 3], [-2.25, 1.5]]
    u = Vector.sparse(UINT8, 4)
    assert lib.GrB_Vector_setElement_INT64(u._vector[0], 300, 0) == 0   # C-style wrap on setElement typecast
    x = ffi.new("int64_t*")
    assert lib.GrB_Vector_extractElement_INT64(x, u._vector[0], 0) == 0 and x[0] == 300 % 256
    assert lib.GrB_Vector_setElement_FP64(u._vector[0], -3.7, 1) == 0   # float -> uint8 saturates at 0
    assert u[1] == 0
    assert lib.GrB_Vector_setElement_FP64(u._vector[0], 3.7, 2) == 0    # truncation
    assert u[2] == 3


def test_build_rules():
    m = Matrix.sparse(INT64, 3, 3)
    I = ffi.new("GrB_Index[]", [0, 0, 1]); J = ffi.new("GrB_Index[]", [1, 1, 2]); X = ffi.new("int64_t[]", [4, 5, 6])
    assert lib.GrB_Matrix_build_INT64(m._matrix[0], I, J, X, 3, lib.GrB_PLUS_INT64) == 0
    assert m.to_lists() == [[0, 1], [1, 2], [9, 6]]
    assert lib.GrB_Matrix_build_INT64(m._matrix[0], I, J, X, 3, lib.GrB_PLUS_INT64) == lib.GrB_OUTPUT_NOT_EMPTY
    m2 = Matrix.sparse(INT64, 2, 2)
    assert lib.GrB_Matrix_build_INT64(m2._matrix[0], I, J, X, 3, ffi.NULL) == lib.GrB_INDEX_OUT_OF_BOUNDS


def test_operator_objects_and_ztype():
    assert INT64.PLUS_TIMES.ztype is INT64 and FP32.LOR_EQ.ztype is BOOL and BOOL.LOR_LAND.ztype is BOOL
    assert INT64.min_plus is INT64.MIN_PLUS and INT64.min.name == "MIN_INT64"
    assert len(gb.semirings) > 1300 and len(gb.monoids) > 60 and len(gb.binaryops) > 300
    # promotion: tests/test_matrix.py:1017-1028
    assert gb.promote(FP32, FP64) is FP64 and gb.promote(FP32, UINT8) is FP32 and gb.promote(INT8, UINT8) is INT8
    assert gb.promote(BOOL, INT8) is INT8


def test_descriptors():
    # tests/test_descriptor.py:6-10
    assert descriptor.T1 != descriptor.T0
    assert descriptor.T1 in descriptor.CT1
    assert descriptor.CT1 == (descriptor.C & descriptor.T1)
    assert descriptor.T0 not in descriptor.RC and descriptor.T0 in descriptor.RCT0
    d = descriptor.Descriptor()
    d[lib.GrB_MASK] = lib.GrB_COMP
    d[lib.GrB_MASK] = lib.GrB_STRUCTURE
    assert d[lib.GrB_MASK] == lib.GrB_COMP + lib.GrB_STRUCTURE
    assert descriptor.Default[lib.GrB_INP0] == lib.GxB_DEFAULT
    # freeing builtin / NULL descriptors is a no-op (descriptor.py:76-78,148)
    p = ffi.new("GrB_Descriptor*", lib.GrB_DESC_T0)
    assert lib.GrB_Descriptor_free(p) == 0 and lib.GrB_Descriptor_free(ffi.new("GrB_Descriptor*")) == 0


def test_argument_errors_are_checked_before_compute():
    m = Matrix.from_lists([0, 1, 2], [1, 2, 0], [1, 2, 3])
    v = Vector.from_lists([0, 1], [1, 2], size=2)
    with pytest.raises(gb.DimensionMismatch):
        m.mxv(v)
    with pytest.raises(gb.DimensionMismatch):
        m.mxm(Matrix.sparse(INT64, 2, 2))
    with pytest.raises(TypeError):
        m @ 3                                        # tests/test_matrix.py:287-288
    with pytest.raises(TypeError):
        v @ v


@pytest.mark.skipif(gb.have_device(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_device():
    m = Matrix.from_lists([0, 1, 2], [1, 2, 0], [1, 2, 3])
    v = Vector.from_lists([0, 1, 2], [2, 3, 4])
    for call in (lambda: m.mxv(v), lambda: v.vxm(m), lambda: m.mxm(m), lambda: m.transpose()):
        with pytest.raises(gb.Panic, match="no CPU fallback"):
            call()


def test_user_defined_operators_are_refused():
    @ffi.callback("void(void*, const void*, const void*)")
    def fn(z, x, y):
        pass
    op = ffi.new("GrB_BinaryOp*")
    assert lib.GrB_BinaryOp_new(op, fn, lib.GrB_FP32, lib.GrB_FP32, lib.GrB_FP32) == 0
    mon = ffi.new("GrB_Monoid*")
    assert lib.GrB_Monoid_new_FP32(mon, op[0], 0.0) == 0
    sr = ffi.new("GrB_Semiring*")
    assert lib.GrB_Semiring_new(sr, mon[0], op[0]) == 0
    m = Matrix.from_lists([0], [0], [1.0], typ=FP32)
    out = Matrix.sparse(FP32, 1, 1)
    # tests/test_udt.py:89-140 style semirings: refused, never run on the CPU
    assert lib.GrB_mxm(out._matrix[0], ffi.NULL, ffi.NULL, sr[0], m._matrix[0], m._matrix[0], ffi.NULL) == lib.GrB_INVALID_VALUE


def test_calls_from_a_thread_pool():
    """The reference drives `lib` from a ThreadPool (demo/dnn/challenge.py:48-51): entry points take the library lock
    and make the library's device current for the calling thread; distinct objects from distinct threads must work."""
    from concurrent.futures import ThreadPoolExecutor
    import pygraphblas_b200 as gb
    from pygraphblas_b200 import Matrix, Vector

    def work(k):
        m = Matrix.from_lists([0, 1, 2], [1, 2, 0], [k, k + 1, k + 2])
        v = Vector.from_lists([0, 2], [k, -k], 3)
        m[2, 2] = 7
        assert m.nvals == 4 and v.nvals == 2 and m[0, 1] == k and m.dup().to_lists() == m.to_lists()
        try:
            w = m.mxv(v)
            assert gb.have_device() and w.size == 3
        except gb.base.Panic:
            assert not gb.have_device()
        return k

    with ThreadPoolExecutor(8) as ex:
        assert sorted(ex.map(work, range(64))) == list(range(64))


def test_monoid_new_checks_operator_and_identity():
    """A monoid over a builtin operator must be associative / commutative with the identity the kernels use
    (accumulators are initialised from the operator): MINUS is refused, a wrong identity is refused, the right one accepted."""
    mon = ffi.new("GrB_Monoid*")
    assert lib.GrB_Monoid_new_INT64(mon, lib.GrB_PLUS_INT64, 0) == lib.GrB_SUCCESS
    assert lib.GrB_Monoid_free(mon) == lib.GrB_SUCCESS
    assert lib.GrB_Monoid_new_FP32(mon, lib.GrB_MIN_FP32, float("inf")) == lib.GrB_SUCCESS
    assert lib.GrB_Monoid_free(mon) == lib.GrB_SUCCESS
    assert lib.GrB_Monoid_new_INT64(mon, lib.GrB_MINUS_INT64, 0) == lib.GrB_DOMAIN_MISMATCH
    assert lib.GrB_Monoid_new_INT64(mon, lib.GrB_PLUS_INT64, 1) == lib.GrB_INVALID_VALUE
    assert lib.GrB_Monoid_new_INT8(mon, lib.GrB_MAX_INT8, -128) == lib.GrB_SUCCESS
    assert lib.GrB_Monoid_free(mon) == lib.GrB_SUCCESS
