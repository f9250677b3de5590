"""CPU: argument validation of the compute entry points happens before any device work and uses the GraphBLAS error
codes the reference maps to exceptions (base.py:189-203): NULL / invalid handles, dimension mismatches, index lists
out of bounds, user-defined operators -- and only then the no-GPU Panic."""
import pytest
import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, Vector, INT64, FP32, BOOL, lib, ffi
from pygraphblas_b200.base import (DimensionMismatch, IndexOutOfBound, NullPointer, UninitializedObject, InvalidValue, Panic)

NULL = ffi.NULL


def raises_or_panics(exc, fn):
    """fn must raise exc; with a GPU present some of these calls would simply succeed only if the arguments were valid,
    so exc is required either way."""
    with pytest.raises(exc):
        fn()


def test_dimension_mismatch_comes_before_the_device_check():
    v3, v4 = Vector.from_lists([0], [1], 3), Vector.from_lists([0], [1], 4)
    m34 = Matrix.from_lists([0], [0], [1], 3, 4)
    raises_or_panics(DimensionMismatch, lambda: v3.eadd(v4))
    raises_or_panics(DimensionMismatch, lambda: v3.emult(v4))
    raises_or_panics(DimensionMismatch, lambda: v3.apply(INT64.AINV, out=v4))
    raises_or_panics(DimensionMismatch, lambda: v3.assign(v4))
    raises_or_panics(DimensionMismatch, lambda: v3.assign_scalar(1, mask=v4))
    raises_or_panics(DimensionMismatch, lambda: v4.extract([0, 1], out=v3))
    raises_or_panics(DimensionMismatch, lambda: m34.mxv(v3))
    raises_or_panics(DimensionMismatch, lambda: v4.vxm(m34))
    raises_or_panics(DimensionMismatch, lambda: m34.mxm(m34))
    raises_or_panics(DimensionMismatch, lambda: m34.eadd(Matrix.sparse(INT64, 4, 3)))
    raises_or_panics(DimensionMismatch, lambda: m34.tril().select(">", 0, out=Matrix.sparse(INT64, 4, 3)) if gb.have_device() else m34.select(">", 0, out=Matrix.sparse(INT64, 4, 3)))
    raises_or_panics(DimensionMismatch, lambda: m34.reduce_vector(out=v4))
    raises_or_panics(DimensionMismatch, lambda: m34.transpose(out=Matrix.sparse(INT64, 3, 4)))


def test_index_lists_are_checked():
    v = Vector.from_lists([0], [1], 5)
    raises_or_panics(IndexOutOfBound, lambda: v.extract([0, 7]))
    raises_or_panics(IndexOutOfBound, lambda: v.assign_scalar(1, [9]))
    raises_or_panics(IndexOutOfBound, lambda: v.assign(Vector.from_lists([0], [1], 2), [1, 5]))


def test_null_and_invalid_handles():
    v = Vector.from_lists([0], [1], 3)
    h = v._vector[0]
    assert lib.GrB_Vector_eWiseAdd_BinaryOp(h, NULL, NULL, NULL, h, h, NULL) == lib.GrB_NULL_POINTER
    assert lib.GrB_Vector_eWiseAdd_BinaryOp(NULL, NULL, NULL, INT64.PLUS.get_op(), h, h, NULL) == lib.GrB_NULL_POINTER
    assert lib.GrB_Vector_apply(h, NULL, NULL, NULL, h, NULL) == lib.GrB_NULL_POINTER
    assert lib.GrB_Vector_assign(h, NULL, NULL, h, NULL, 0, NULL) == lib.GrB_NULL_POINTER          # NULL index list
    x = ffi.new("int64_t*")
    assert lib.GrB_Vector_reduce_INT64(x, NULL, NULL, h, NULL) in (lib.GrB_NULL_POINTER, lib.GrB_PANIC)
    assert lib.GxB_Vector_select(h, NULL, NULL, NULL, h, NULL, NULL) == lib.GrB_NULL_POINTER
    M = Matrix.from_lists([0], [0], [1], 2, 2)
    m = M._matrix
    assert lib.GxB_Matrix_select(m[0], NULL, NULL, lib.GxB_GT_THUNK, m[0], NULL, NULL) == lib.GrB_INVALID_VALUE   # thunk missing
    assert lib.GrB_Matrix_apply(m[0], NULL, NULL, NULL, m[0], NULL) == lib.GrB_NULL_POINTER
    assert lib.GrB_mxv(h, NULL, NULL, NULL, m[0], h, NULL) == lib.GrB_NULL_POINTER
    dead = Vector.from_lists([0], [1], 3)
    dh = dead._vector[0]
    lib.GrB_Vector_free(dead._vector)
    assert lib.GrB_Vector_eWiseMult_BinaryOp(h, NULL, NULL, INT64.TIMES.get_op(), h, dh, NULL) in (lib.GrB_NULL_POINTER, lib.GrB_UNINITIALIZED_OBJECT)


def test_valid_calls_panic_without_a_gpu_and_only_then():
    if gb.have_device():
        pytest.skip("a GPU is present: these calls compute")
    v = Vector.from_lists([0, 1], [1, 2], 3)
    m = Matrix.from_lists([0, 1], [1, 0], [1, 2], 3, 3)
    for fn in (lambda: v.eadd(v), lambda: v.emult(v), lambda: -v, lambda: v + 1, lambda: v.assign_scalar(1), lambda: v[0:1],
               lambda: v.reduce_int(), lambda: v.select(">", 0), lambda: m.tril(), lambda: m.apply(INT64.ABS), lambda: m.reduce_int(),
               lambda: m.reduce_vector(), lambda: m.eadd(m), lambda: m.emult(m), lambda: m.mxv(v), lambda: v.vxm(m), lambda: m.mxm(m),
               lambda: m.transpose()):
        with pytest.raises(Panic) as e:
            fn()
        assert "no CPU fallback" in str(e.value)
