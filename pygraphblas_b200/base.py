"""Error mapping and library access for the host-side mirror.

Mirrors /root/reference/pygraphblas/base.py: the GrB_Info -> exception table
(base.py:189-203) and the `_check` helper (base.py:206-210, matrix.py:43-51).
"""
from ._ffi import ffi, lib, initialize, is_initialized

NULL = ffi.NULL

if not is_initialized():
    initialize(blocking=False, memory_manager="c")


class GraphBLASException(Exception):
    pass


class NoValue(GraphBLASException):
    pass


class UninitializedObject(GraphBLASException):
    pass


class InvalidObject(GraphBLASException):
    pass


class NullPointer(GraphBLASException):
    pass


class InvalidValue(GraphBLASException):
    pass


class InvalidIndex(GraphBLASException):
    pass


class DomainMismatch(GraphBLASException):
    pass


class DimensionMismatch(GraphBLASException):
    pass


class OutputNotEmpty(GraphBLASException):
    pass


class OutOfMemory(GraphBLASException):
    pass


class InsufficientSpace(GraphBLASException):
    pass


class IndexOutOfBound(GraphBLASException):
    pass


class Panic(GraphBLASException):
    pass


_error_codes = {
    1: NoValue, 2: UninitializedObject, 3: InvalidObject, 4: NullPointer, 5: InvalidValue,
    6: InvalidIndex, 7: DomainMismatch, 8: DimensionMismatch, 9: OutputNotEmpty, 10: OutOfMemory,
    11: InsufficientSpace, 12: IndexOutOfBound, 13: Panic,
}


def _check(res):
    """Raise the exception class the reference raises for a non-zero GrB_Info."""
    if res != lib.GrB_SUCCESS:
        raise _error_codes[res](ffi.string(lib.B200_last_error()).decode("utf8", "replace"))


def have_device():
    """True when a CUDA device is usable; without one every mxm/mxv/vxm raises Panic."""
    return bool(lib.B200_have_device())
