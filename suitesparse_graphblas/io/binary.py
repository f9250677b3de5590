"""binread / binwrite with the signatures the reference calls (/root/reference/pygraphblas/matrix.py:494, 939):

    matrix = binary.binread(bin_file, opener)            -> a `GrB_Matrix*` handle the reference wraps with Matrix(handle)
    binary.binwrite(self._matrix, filename, comments, opener)

The file layout is pygraphblas_b200/io.py's (grb_read / grb_write); the matrix is handed to the library with one bulk build."""
from pathlib import Path

import numpy as np

from pygraphblas_b200._ffi import ffi, lib
from pygraphblas_b200 import io as _io
from pygraphblas_b200 import types as _types


def _check(info, what):
    if info != lib.GrB_SUCCESS:
        raise RuntimeError(f"{what} failed: {ffi.string(lib.B200_last_error()).decode()}")


def binread(filename, opener=Path.open):
    I, J, V, nrows, ncols, typ = _io.grb_read(Path(filename), opener)
    A = ffi.new("GrB_Matrix*")
    _check(lib.GrB_Matrix_new(A, typ.gb_type, nrows, ncols), "GrB_Matrix_new")
    I = np.ascontiguousarray(I, np.uint64); J = np.ascontiguousarray(J, np.uint64); V = np.ascontiguousarray(V, typ.dtype)
    _check(typ._Matrix_build(A[0], ffi.cast("GrB_Index*", I.ctypes.data), ffi.cast("GrB_Index*", J.ctypes.data), ffi.cast(typ.ptr, V.ctypes.data), len(I), ffi.NULL),
           "GrB_Matrix_build")
    return A


def binwrite(A, filename, comments="", opener=Path.open):
    t = ffi.new("GrB_Type*")
    _check(lib.GxB_Matrix_type(t, A[0]), "GxB_Matrix_type")
    typ = _types.from_handle(t[0])
    nr, nc, nv = ffi.new("GrB_Index*"), ffi.new("GrB_Index*"), ffi.new("GrB_Index*")
    _check(lib.GrB_Matrix_nrows(nr, A[0]), "GrB_Matrix_nrows"); _check(lib.GrB_Matrix_ncols(nc, A[0]), "GrB_Matrix_ncols"); _check(lib.GrB_Matrix_nvals(nv, A[0]), "GrB_Matrix_nvals")
    n = int(nv[0])
    I, J, V = np.empty(max(n, 1), np.uint64), np.empty(max(n, 1), np.uint64), np.empty(max(n, 1), typ.dtype)
    _check(typ._Matrix_extractTuples(ffi.cast("GrB_Index*", I.ctypes.data), ffi.cast("GrB_Index*", J.ctypes.data), ffi.cast(typ.ptr, V.ctypes.data), nv, A[0]),
           "GrB_Matrix_extractTuples")
    _io.grb_write(Path(filename), I[:n], J[:n], V[:n], int(nr[0]), int(nc[0]), typ, comments, opener)
