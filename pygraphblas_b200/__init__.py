"""pygraphblas_b200 -- the B200-native mxm / mxv / vxm hot path behind the pygraphblas API.

Layout (only what the hot path needs):
    csrc/            CUDA (sm_100a) kernels + the C-ABI library libb200grb.so
    _ffi.py          CFFI binding of include/b200grb.h (`lib`, `ffi`, `initialize`)
    base/types/ops/descriptor/matrix/vector.py
                     host-side mirror of the reference's operator / container interface
                     for this path (same names and semantics as /root/reference/pygraphblas)

Usage is the reference's:
    from pygraphblas_b200 import Matrix, Vector, INT64, BOOL, descriptor, Accum
    C = A.mxm(B, semiring=INT64.PLUS_TIMES, mask=M, desc=descriptor.RC)
"""
from .base import (lib, ffi, have_device, GraphBLASException, NoValue, UninitializedObject, InvalidObject, NullPointer,
                   InvalidValue, InvalidIndex, DomainMismatch, DimensionMismatch, OutputNotEmpty, OutOfMemory,
                   InsufficientSpace, IndexOutOfBound, Panic)
from . import types
from .types import (BOOL, INT8, INT16, INT32, INT64, UINT8, UINT16, UINT32, UINT64, FP32, FP64, promote)
from .ops import BinaryOp, Accum, Monoid, Semiring, current_semiring, current_accum, binaryops, monoids, semirings
from . import descriptor
from .matrix import Matrix
from .vector import Vector
from .scalar import Scalar

__all__ = ["lib", "ffi", "have_device", "Matrix", "Vector", "types", "descriptor", "Accum", "BinaryOp", "Monoid", "Semiring",
           "BOOL", "INT8", "INT16", "INT32", "INT64", "UINT8", "UINT16", "UINT32", "UINT64", "FP32", "FP64", "promote"]
