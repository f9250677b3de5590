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


# ---------------------------------------------------------------------------------------------------------------
# `.grb` binary files: the LAGraph binwrite layout that suitesparse_graphblas.io.binary reads and writes for the
# reference (/root/reference/pygraphblas/matrix.py:489-497 binread, :935-942 binwrite; fixture docs/test_binfile.grb).
# Layout (little endian), reconstructed from the fixture and checked against docs/test_mm.mm, which holds the same matrix:
#   512-byte text header ("SuiteSparse:GraphBLAS matrix\n<version>\nnrows: ..\nncols: ..\nnvec: ..\nnvals: ..\nformat: ..\nsize: ..\ntype: ..\n")
#   int32 fmt (0 = by row, 1 = by column), int32 kind (1 hypersparse, 2 sparse, 4 bitmap, 8 full), double hyper_switch,
#   uint64 nrows, uint64 ncols, int64 nonempty, uint64 nvec, uint64 nvals, int32 typecode, uint64 typesize   (packed)
#   then, by kind:  Ap[nvec+1] u64 (sparse, hyper) | Ah[nvec] u64 (hyper) | Ab[nrows*ncols] i8 (bitmap) | Ai[nvals] u64 (sparse, hyper)
#                   | Ax[nvals or nrows*ncols] of the type
_GRB_TYPECODES = ["BOOL", "INT8", "INT16", "INT32", "INT64", "UINT8", "UINT16", "UINT32", "UINT64", "FP32", "FP64"]
_GRB_HEADER = 512
_KIND_NAMES = {1: "HYPER", 2: "SPARSE", 4: "BITMAP", 8: "FULL"}


def grb_read(bin_file, opener=None):
    """-> (I, J, V, nrows, ncols, typ) of a `.grb` file (every storage kind, by row or by column)."""
    import struct
    if opener is not None:
        with opener(bin_file, "rb") as f:
            data = f.read()
    elif hasattr(bin_file, "read"):
        data = bin_file.read()
    else:
        with open(bin_file, "rb") as f:
            data = f.read()
    if not data.startswith(b"SuiteSparse:GraphBLAS matrix"):
        raise ValueError("not a SuiteSparse:GraphBLAS binary matrix file")
    fmt, kind, _hyper, nrows, ncols, _nonempty, nvec, nvals, typecode, typesize = struct.unpack_from("<iidQQqQQiQ", data, _GRB_HEADER)
    if not 0 <= typecode < len(_GRB_TYPECODES):
        raise ValueError(f"type code {typecode} is not supported (complex and user-defined types are out of scope)")
    typ = types.by_name(_GRB_TYPECODES[typecode])
    if typ.dtype.itemsize != typesize:
        raise ValueError("type size does not match the type code")
    off = _GRB_HEADER + struct.calcsize("<iidQQqQQiQ")

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(data, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a

    vdim, vlen = (nrows, ncols) if fmt == 0 else (ncols, nrows)          # vectors = rows when stored by row
    if kind in (1, 2):
        Ap = take("<u8", nvec + 1)
        Ah = take("<u8", nvec) if kind == 1 else np.arange(nvec, dtype=np.uint64)
        Ai = take("<u8", nvals)
        Ax = take(typ.dtype, nvals)
        major = np.repeat(Ah, np.diff(Ap).astype(np.int64))
        minor = Ai
    elif kind == 4:
        Ab = take("i1", vdim * vlen)
        Ax = take(typ.dtype, vdim * vlen)
        pos = np.flatnonzero(Ab)
        major, minor, Ax = (pos // vlen).astype(np.uint64), (pos % vlen).astype(np.uint64), Ax[pos]
    elif kind == 8:
        Ax = take(typ.dtype, vdim * vlen)
        pos = np.arange(vdim * vlen)
        major, minor = (pos // vlen).astype(np.uint64), (pos % vlen).astype(np.uint64)
    else:
        raise ValueError(f"unknown storage kind {kind}")
    I, J = (major, minor) if fmt == 0 else (minor, major)
    return np.ascontiguousarray(I), np.ascontiguousarray(J), np.ascontiguousarray(Ax), int(nrows), int(ncols), typ


def grb_write(bin_file, I, J, V, nrows, ncols, typ, comments="", opener=None):
    """Write tuples as a `.grb` file in the SPARSE, by-row form (entries sorted row-major)."""
    import struct
    I = np.asarray(I, dtype=np.uint64); J = np.asarray(J, dtype=np.uint64); V = np.asarray(V, dtype=typ.dtype)
    order = np.lexsort((J, I))
    I, J, V = I[order], J[order], V[order]
    nvals = len(I)
    Ap = np.zeros(nrows + 1, np.uint64)
    np.cumsum(np.bincount(I.astype(np.int64), minlength=nrows), out=Ap[1:])
    head = ("SuiteSparse:GraphBLAS matrix\nv5.1 (libb200grb)\nnrows:  %-18d\nncols:  %-18d\nnvec:   %-18d\nnvals:  %-18d\nformat: SPARSER \n"
            "size:   %-18d\ntype:   GrB_%-72s\n%s\n" % (nrows, ncols, nrows, nvals, typ.dtype.itemsize, typ.name, comments[:200]))
    head = head.encode()[:_GRB_HEADER - 2].ljust(_GRB_HEADER - 2, b" ") + b"\n\x00"
    blob = head + struct.pack("<iidQQqQQiQ", 0, 2, 0.0625, nrows, ncols, -1, nrows, nvals, _GRB_TYPECODES.index(typ.name), typ.dtype.itemsize)
    blob += Ap.astype("<u8").tobytes() + J.astype("<u8").tobytes() + V.tobytes()
    if opener is not None:
        with opener(bin_file, "wb") as f:
            f.write(blob)
    elif hasattr(bin_file, "write"):
        bin_file.write(blob)
    else:
        with open(bin_file, "wb") as f:
            f.write(blob)
