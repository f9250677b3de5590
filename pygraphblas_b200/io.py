"""Matrix-Market and delimited-text ingest / egress.

The reference reads Matrix-Market files entry by entry through `setElement`
(/root/reference/pygraphblas/matrix.py:377-409, third-party `mmparse`) and tab-separated files the same way
(:411-475).  Here a file is parsed in bulk with numpy and handed to the library as one `build`, the same bulk
path the benchmark graphs take; the coordinate format, the `integer` / `real` / `pattern` fields, `general` /
`symmetric` / `skew-symmetric` storage and the `%%GraphBLAS GrB_<TYPE>` hint line the reference writes are
understood.  Host-side code only: nothing here touches the GPU.
"""
import io
import numpy as np
from . import types

_FIELD_TYPE = {"integer": "INT64", "real": "FP64", "double": "FP64", "pattern": "BOOL"}


def _open(f, mode):
    if hasattr(f, "read") or hasattr(f, "write"):
        return f, False
    return open(f, mode), True


def mm_read(mm_file):
    """-> (I, J, V, nrows, ncols, typ) of a coordinate Matrix-Market file (0-based indices, duplicates kept)."""
    f, close = _open(mm_file, "r")
    try:
        head = f.readline().split()
        if len(head) < 5 or head[0] != "%%MatrixMarket" or head[1].lower() != "matrix" or head[2].lower() != "coordinate":
            raise ValueError("not a coordinate Matrix-Market file")
        field, storage = head[3].lower(), head[4].lower()
        if field not in _FIELD_TYPE:
            raise ValueError(f"Matrix-Market field '{field}' is not supported (complex types are out of scope)")
        typ = types.by_name(_FIELD_TYPE[field])
        line = f.readline()
        while line.startswith("%"):
            if line.startswith("%%GraphBLAS"):
                name = line.split()[1].replace("GrB_", "").replace("GxB_", "")
                typ = types.by_name(name)
            line = f.readline()
        nrows, ncols, nnz = (int(x) for x in line.split())
        body = f.read()
    finally:
        if close:
            f.close()
    if nnz == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, typ.dtype), nrows, ncols, typ
    cols = 2 if field == "pattern" else 3
    data = np.loadtxt(io.StringIO(body), dtype=np.float64 if typ.dtype.kind == "f" else np.int64, ndmin=2, usecols=range(cols), max_rows=nnz)
    I = data[:, 0].astype(np.int64) - 1
    J = data[:, 1].astype(np.int64) - 1
    V = np.ones(len(I), typ.dtype) if field == "pattern" else data[:, 2].astype(typ.dtype)
    if storage in ("symmetric", "skew-symmetric"):
        off = I != J
        sign = -1 if storage == "skew-symmetric" else 1
        I, J, V = np.concatenate([I, J[off]]), np.concatenate([J, I[off]]), np.concatenate([V, (sign * V[off]).astype(typ.dtype)])
    elif storage != "general":
        raise ValueError(f"Matrix-Market storage '{storage}' is not supported")
    return I.astype(np.uint64), J.astype(np.uint64), V, nrows, ncols, typ


def mm_write(matrix, mm_file):
    """Write `matrix` in coordinate format with the `%%GraphBLAS` type line; symmetric matrices store their lower
    triangle only (tests/test_matrix.py:329-346 of the reference pins the layout)."""
    I, J, V = matrix.to_arrays()
    I, J = I.astype(np.int64), J.astype(np.int64)
    typ = matrix.type
    field = "pattern" if typ is types.BOOL else ("real" if typ.dtype.kind == "f" else "integer")
    symmetric = matrix.nrows == matrix.ncols
    if symmetric:
        a = {(int(i), int(j)): v for i, j, v in zip(I, J, V)}
        symmetric = all(a.get((j, i)) == v for (i, j), v in a.items())
    if symmetric:
        keep = I >= J
        I, J, V = I[keep], J[keep], V[keep]
    f, close = _open(mm_file, "w")
    try:
        f.write(f"%%MatrixMarket matrix coordinate {field} {'symmetric' if symmetric else 'general'}\n")
        f.write(f"%%GraphBLAS GrB_{typ.name}\n")
        f.write(f"{matrix.nrows} {matrix.ncols} {len(I)}\n")
        for i, j, v in zip(I, J, V):
            if field == "pattern":
                f.write(f"{i + 1} {j + 1}\n")
            elif field == "real":
                f.write(f"{i + 1} {j + 1} {float(v)!r}\n")
            else:
                f.write(f"{i + 1} {j + 1} {int(v)}\n")
    finally:
        if close:
            f.close()


def delimited_read(path, typ, one_based=True, delimiter="\t"):
    """-> (I, J, V) of a `row<delim>col<delim>value` text file (matrix.py:411-475 of the reference)."""
    data = np.loadtxt(path, dtype=np.float64 if typ.dtype.kind == "f" else np.int64, delimiter=delimiter, ndmin=2)
    if data.shape[1] != 3:
        raise TypeError("File can contain only 3 columns: row, col and val")
    off = 1 if one_based else 0
    return (data[:, 0].astype(np.int64) - off).astype(np.uint64), (data[:, 1].astype(np.int64) - off).astype(np.uint64), data[:, 2].astype(typ.dtype)
