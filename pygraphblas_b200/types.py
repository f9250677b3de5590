"""GraphBLAS builtin types and their operator namespaces.

Mirrors the surface of /root/reference/pygraphblas/types.py for the hot path:
one class-like object per builtin type (types.py:179-345) carrying its binary
operators, monoids and semirings as attributes in both cases (`INT64.PLUS_TIMES`,
`INT64.min_plus`, `INT64.min`, `BOOL.LOR_LAND`; semiring.py:36-44, binaryop.py:44-47),
`promote` (types.py:484-500) and the default semiring rule (types.py:156-158, 198-200).
User-defined and complex types are out of scope.
"""
import numpy as np
from .base import lib, ffi


class Type:
    """One builtin GraphBLAS type."""

    def __init__(self, name, c_type, dtype, gb_type):
        self.name = self.__name__ = name
        self.c_type = c_type
        self.dtype = np.dtype(dtype)
        self.gb_type = gb_type
        self.ptr = c_type + "*"
        fn = lambda n: getattr(lib, f"GrB_{n}_{name}")
        self._Matrix_setElement = fn("Matrix_setElement")
        self._Matrix_extractElement = fn("Matrix_extractElement")
        self._Matrix_extractTuples = fn("Matrix_extractTuples")
        self._Matrix_build = fn("Matrix_build")
        self._Vector_setElement = fn("Vector_setElement")
        self._Vector_extractElement = fn("Vector_extractElement")
        self._Vector_extractTuples = fn("Vector_extractTuples")
        self._Vector_build = fn("Vector_build")
        self._Vector_assignScalar = fn("Vector_assign")                 # types.py:103 of the reference
        self._Vector_apply_BinaryOp1st = fn("Vector_apply_BinaryOp1st")
        self._Vector_apply_BinaryOp2nd = fn("Vector_apply_BinaryOp2nd")
        self._Vector_reduce = fn("Vector_reduce")
        self._Matrix_reduce = fn("Matrix_reduce")
        self._Matrix_apply_BinaryOp1st = fn("Matrix_apply_BinaryOp1st")
        self._Matrix_apply_BinaryOp2nd = fn("Matrix_apply_BinaryOp2nd")

    def _default_addop(self):                                           # types.py:157-160
        return self.LOR if self is BOOL else self.PLUS

    def _default_multop(self):
        return self.LAND if self is BOOL else self.TIMES

    def __repr__(self):
        return f"<Type {self.name}>"

    def _default_semiring(self):
        return self.LOR_LAND if self is BOOL else self.PLUS_TIMES

    def from_value(self, v):
        return bool(v) if self is BOOL else (float(v) if self.dtype.kind == "f" else int(v))


BOOL = Type("BOOL", "_Bool", np.bool_, lib.GrB_BOOL)
INT8 = Type("INT8", "int8_t", np.int8, lib.GrB_INT8)
INT16 = Type("INT16", "int16_t", np.int16, lib.GrB_INT16)
INT32 = Type("INT32", "int32_t", np.int32, lib.GrB_INT32)
INT64 = Type("INT64", "int64_t", np.int64, lib.GrB_INT64)
UINT8 = Type("UINT8", "uint8_t", np.uint8, lib.GrB_UINT8)
UINT16 = Type("UINT16", "uint16_t", np.uint16, lib.GrB_UINT16)
UINT32 = Type("UINT32", "uint32_t", np.uint32, lib.GrB_UINT32)
UINT64 = Type("UINT64", "uint64_t", np.uint64, lib.GrB_UINT64)
FP32 = Type("FP32", "float", np.float32, lib.GrB_FP32)
FP64 = Type("FP64", "double", np.float64, lib.GrB_FP64)

ALL_TYPES = (BOOL, INT8, INT16, INT32, INT64, UINT8, UINT16, UINT32, UINT64, FP32, FP64)
_by_name = {t.name: t for t in ALL_TYPES}
_by_handle = {t.gb_type: t for t in ALL_TYPES}

# types.py:468-481
_promotion_order = (FP64, FP32, INT64, UINT64, INT32, UINT32, INT16, UINT16, INT8, UINT8)


def by_name(name):
    return _by_name[name]


def from_handle(gb_type):
    return _by_handle[gb_type]


def from_python(value):
    """Type inferred from a python scalar the way Matrix.from_lists does (matrix.py:306-313)."""
    if isinstance(value, (bool, np.bool_)):
        return BOOL
    if isinstance(value, (int, np.integer)):
        return INT64
    if isinstance(value, (float, np.floating)):
        return FP64
    raise TypeError(f"no GraphBLAS type for {type(value)!r}")


def promote(left, right):
    if left is right:
        return left
    if left is BOOL:
        return right
    if right is BOOL:
        return left
    for t in _promotion_order:
        if left is t or right is t:
            return t
    raise TypeError(f"inconvertable types {left!r} and {right!r}")


def _dtype_lookup(dtype):
    dtype = np.dtype(dtype)
    for t in ALL_TYPES:
        if t.dtype == dtype:
            return t
    raise TypeError(f"no GraphBLAS type for dtype {dtype}")
