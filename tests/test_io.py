"""CPU: Matrix-Market / delimited-text ingest and egress (pygraphblas_b200/io.py) against the reference's own
known answers: tests/test_matrix.py:329-357 and the doctest of matrix.py:378-393 (docs/test_mm.mm, whose 12 entries
are written out below)."""
import numpy as np
from pygraphblas_b200 import Matrix, INT8, INT64, FP64, BOOL

TEST_MM = """%%MatrixMarket matrix coordinate integer general
%%GraphBLAS GrB_INT64
7 7 12
1 2 0
1 4 1
2 5 2
2 7 3
3 6 4
4 1 5
4 3 6
5 6 7
6 3 8
7 3 9
7 4 10
7 5 11
"""


def test_mm_read_write_known_answers(tmp_path):
    mmf = tmp_path / "mmwrite_test.mm"
    m = Matrix.from_lists([0, 1, 2], [0, 1, 2], [2, 3, 4])
    with mmf.open("w") as f:
        m.to_mm(f)
    assert mmf.open().readlines() == ["%%MatrixMarket matrix coordinate integer symmetric\n", "%%GraphBLAS GrB_INT64\n", "3 3 3\n",
                                      "1 1 2\n", "2 2 3\n", "3 3 4\n"]
    assert Matrix.from_mm(mmf).iseq(m)
    p = tmp_path / "test_mm.mm"
    p.write_text(TEST_MM)
    M = Matrix.from_mm(p)
    assert M.type is INT64 and M.shape == (7, 7)
    assert M.to_lists() == [[0, 0, 1, 1, 2, 3, 3, 4, 5, 6, 6, 6], [1, 3, 4, 6, 5, 0, 2, 5, 2, 2, 3, 4], list(range(12))]


def test_mm_symmetric_pattern_real(tmp_path):
    p = tmp_path / "s.mm"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% a comment\n3 3 3\n1 1 1.5\n2 1 2.5\n3 2 -1\n")
    M = Matrix.from_mm(p)
    assert M.type is FP64 and M.to_lists() == [[0, 0, 1, 1, 2], [0, 1, 0, 2, 1], [1.5, 2.5, 2.5, -1.0, -1.0]]
    q = tmp_path / "p.mm"
    q.write_text("%%MatrixMarket matrix coordinate pattern general\n2 3 2\n1 3\n2 1\n")
    P = Matrix.from_mm(q)
    assert P.type is BOOL and P.shape == (2, 3) and P.to_lists() == [[0, 1], [2, 0], [True, True]]
    out = tmp_path / "o.mm"
    with out.open("w") as f:
        M.to_mm(f)
    assert Matrix.from_mm(out).iseq(M)
    g = Matrix.from_lists([0, 1], [1, 0], [1.0, 2.0], 2, 2, FP64)           # not symmetric: general
    with out.open("w") as f:
        g.to_mm(f)
    assert "general" in out.read_text().splitlines()[0] and Matrix.from_mm(out).iseq(g)


def test_tsv_read_known_answer(tmp_path):
    """tests/test_matrix.py:349-357"""
    f = tmp_path / "tsv_test.mm"
    f.write_text("1\t1\t2\n2\t2\t3\n3\t3\t4\n")
    n = Matrix.from_tsv(f, INT8, 3, 3)
    assert n.type is INT8 and n.to_lists() == [[0, 1, 2], [0, 1, 2], [2, 3, 4]]


def test_convenience_constructors_known_answers():
    """matrix.py:574-594 (identity doctest), vector.py:370-382 (from_1_to_n), tests/test_matrix.py:1060-1068 (scipy round trip)"""
    import scipy.sparse as sp
    from pygraphblas_b200 import Vector, UINT8
    assert Matrix.identity(UINT8, 3, 42).to_lists() == [[0, 1, 2], [0, 1, 2], [42, 42, 42]]
    assert Matrix.identity(FP64, 2).to_lists() == [[0, 1], [0, 1], [1.0, 1.0]]
    assert Vector.from_1_to_n(3).to_lists() == [[0, 1, 2], [1, 2, 3]]
    m = Matrix.from_lists([0, 1, 2], [1, 2, 0], [1, 2, 3])
    s = m.to_scipy_sparse()
    assert sp.isspmatrix_csr(s) and s.dtype == np.int64 and (s.toarray() == np.array([[0, 1, 0], [0, 0, 2], [3, 0, 0]])).all()
    assert Matrix.from_scipy_sparse(s).iseq(m) and m.to_scipy_sparse("coo").nnz == 3


def test_grb_binary_reader_on_the_reference_fixture(tmp_path):
    """`.grb` files (suitesparse_graphblas.io.binary of the binding stub, /root/reference/pygraphblas/matrix.py:489-497):
    the reference's own fixture docs/test_binfile.grb (committed base64, tests/golden/) holds the matrix of docs/test_mm.mm."""
    import base64, io, os
    from pygraphblas_b200 import io as gio
    raw = base64.b64decode(open(os.path.join(os.path.dirname(__file__), "golden", "test_binfile.grb.b64")).read())
    I, J, V, nrows, ncols, typ = gio.grb_read(io.BytesIO(raw))
    assert (nrows, ncols, typ) == (7, 7, INT64)
    want = [tuple(int(x) for x in l.split()) for l in TEST_MM.splitlines()[3:]]
    assert sorted(zip((I + 1).tolist(), (J + 1).tolist(), V.tolist())) == sorted(want)
    # write -> read round trip in the SPARSE form, through the module the reference imports
    from suitesparse_graphblas.io import binary
    from pygraphblas_b200 import lib, ffi
    f = tmp_path / "fixture.grb"
    f.write_bytes(raw)
    A = binary.binread(f)
    nv = ffi.new("GrB_Index*"); assert lib.GrB_Matrix_nvals(nv, A[0]) == 0 and nv[0] == 12
    g = tmp_path / "out.grb"
    binary.binwrite(A, g, comments="round trip")
    I2, J2, V2, nr2, nc2, t2 = gio.grb_read(g)
    assert (nr2, nc2, t2) == (7, 7, INT64)
    assert sorted(zip(I2.tolist(), J2.tolist(), V2.tolist())) == sorted(zip(I.tolist(), J.tolist(), V.tolist()))
    lib.GrB_Matrix_free(A)
