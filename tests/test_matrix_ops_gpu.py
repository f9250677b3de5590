"""GPU parity of the matrix-side consumers of the hot path (matrix_ops.cu) -- SURVEY.md section 8(f)3:
(1) the known answers of the reference's own unit tests (/root/reference/tests/test_matrix.py, cited per case;
    inputs and expected outputs transcribed as data) through the host-side mirror of its API;
(2) random matrices against scipy / numpy for select, apply, reduce, eWiseAdd and eWiseMult with masks,
    accumulators and transposed operands;
(3) triangle counting end to end the way demo/Triangle-Counting.ipynb:581-582 writes it:
    L = A.tril(-1); C = L.mxm(L, mask=L, PLUS_PAIR); C.reduce_int() -- all four steps in HBM."""
import itertools
import numpy as np
import pytest
import scipy.sparse as sp

import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, Vector, BOOL, INT8, INT64, FP32, FP64, descriptor, lib

pytestmark = pytest.mark.gpu


def L3(m):
    I, J, X = m.to_arrays()
    return [I.tolist(), J.tolist(), X.tolist()]


def test_ref_matrix_eadd_sub_emult():
    """tests/test_matrix.py:137-205"""
    I = list(range(10))
    v = Matrix.from_lists(I, I, I); v[0, 1] = 1
    w = Matrix.from_lists(I, I, I); w[1, 0] = 1
    ref = Matrix.from_lists(I, I, list(range(0, 20, 2))); ref[0, 1] = 1; ref[1, 0] = 1
    assert v.eadd(w).iseq(ref) and (v + w).iseq(ref)
    assert v.eadd(w, INT64.SECOND).iseq(w.eadd(v, INT64.FIRST))
    sub = Matrix.from_lists(I, I, [0] * 10); sub[0, 1] = 1; sub[1, 0] = 1
    assert (v - w).iseq(sub)
    V = list(range(1, 11))
    a, b = Matrix.from_lists(I, I, V), Matrix.from_lists(I, I, V)
    assert a.emult(b).iseq(Matrix.from_lists(I, I, [x * x for x in V]))
    assert a.emult(b, INT64.SECOND).iseq(b)
    assert (a / b).iseq(Matrix.from_lists(I, I, [1] * 10))


def test_ref_matrix_reduce():
    """tests/test_matrix.py:208-246"""
    v = Matrix.sparse(BOOL, 10, 10)
    assert not v.reduce_bool()
    v[3, 3] = True; v[4, 4] = False
    assert v.reduce_bool() is True and v.reduce_bool(BOOL.LAND_MONOID) is False
    i = Matrix.sparse(INT8, 10, 10)
    assert i.reduce_int() == 0
    i[3, 3] = 3; i[4, 4] = 4
    assert i.reduce_int() == 7 and i.reduce_int(INT8.TIMES_MONOID) == 12
    f = Matrix.sparse(FP64, 10, 10)
    assert f.reduce_float() == 0.0
    f[3, 3] = 3.3; f[4, 4] = 4.4
    assert f.reduce_float() == 7.7 and f.reduce_float(FP64.TIMES_MONOID) == pytest.approx(14.52, rel=1e-15)
    m = Matrix.from_lists(list(range(10)), list(range(10)), list(range(10)))
    assert m.reduce_vector().iseq(Vector.from_list(list(range(10))))


def test_ref_apply_select():
    """tests/test_matrix.py:536-545, 580-656"""
    v = Matrix.from_lists([0, 1, 2], [0, 1, 2], [2, 3, 4])
    assert L3(v.apply(INT64.AINV)) == [[0, 1, 2], [0, 1, 2], [-2, -3, -4]]
    v = Matrix.from_lists([0, 1, 2], [0, 1, 2], [0, 0, 3])
    for w in (v.select(lib.GxB_NONZERO), v.select("!=0"), v.select("!=", 0), v.select(">", 0)):
        assert L3(w) == [[2], [2], [3]]
    assert L3(v.select("<", 3)) == [[0, 1], [0, 1], [0, 0]]
    assert v.select(">=", 0).iseq(v) and v.select(">=0").iseq(v)
    I, J = map(list, zip(*itertools.product(range(3), repeat=2)))
    m = Matrix.from_lists(I, J, list(range(9)), 3, 3)
    assert L3(m.tril()) == [[0, 1, 1, 2, 2, 2], [0, 0, 1, 0, 1, 2], [0, 3, 4, 6, 7, 8]]
    assert L3(m.triu()) == [[0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2], [0, 1, 2, 4, 5, 8]]
    assert L3(m.diag()) == [[0, 1, 2], [0, 1, 2], [0, 4, 8]]
    assert L3(m.offdiag()) == [[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1], [1, 2, 3, 5, 6, 7]]
    assert L3(m.nonzero()) == [[0, 0, 1, 1, 1, 2, 2, 2], [1, 2, 0, 1, 2, 0, 1, 2], [1, 2, 3, 4, 5, 6, 7, 8]]
    assert L3(-m)[2] == [0, -1, -2, -3, -4, -5, -6, -7, -8] and L3(abs(-m))[2] == list(range(9))
    f = Matrix.from_lists([0, 1, 2], [0, 1, 2], [0.0, 1.0, 2.0], 3, 3)
    assert L3(f.apply(FP64.MINV))[2] == [float("inf"), 1.0, 0.5]
    assert L3(Matrix.from_lists([0, 1], [0, 1], [4, 2]).apply_first(2, INT8.PLUS))[2] == [6, 4]          # :909-912
    assert L3(Matrix.from_lists([0, 1], [0, 1], [5, 1]).apply_second(INT8.MINUS, 2))[2] == [3, -1]        # :914-917


def _rand(rng, m, n, dens, dtype, lo=-4, hi=5):
    A = sp.random(m, n, density=dens, format="csr", random_state=int(rng.integers(1 << 30)), dtype=np.float64)
    A.data = rng.integers(lo, hi, A.nnz).astype(dtype)
    A.sort_indices()
    return A


def _csr_equal(M, S, exact=True):
    p, j, x = M.to_csr()
    S = S.tocsr(); S.sort_indices()
    assert np.array_equal(p, S.indptr) and np.array_equal(j, S.indices)
    assert np.array_equal(x, S.data.astype(x.dtype)) if exact else np.allclose(x, S.data, rtol=1e-6)


def _keep(S, pred):
    """scipy matrix with the entries where pred(i, j, x) holds, explicit zeros preserved"""
    C = S.tocoo()
    k = pred(C.row, C.col, C.data)
    return sp.csr_matrix((C.data[k], (C.row[k], C.col[k])), shape=S.shape)


@pytest.mark.parametrize("seed", range(12))
def test_random_select_apply_reduce(seed):
    rng = np.random.default_rng(6000 + seed)
    m, n = (int(rng.integers(1, 60)), int(rng.integers(1, 60))) if seed % 4 else (3000, 2500)
    typ, dt = [(INT64, np.int64), (FP32, np.float32), (INT8, np.int8), (FP64, np.float64)][seed % 4]
    S = _rand(rng, m, n, 0.3 if m < 100 else 0.01, dt)
    if m >= 100:                                                    # one hub row
        hub = sp.csr_matrix((rng.integers(1, 4, n).astype(dt), (np.zeros(n, int), np.arange(n))), shape=(m, n))
        S = (S + hub).tocsr(); S.sort_indices()
        S.data = S.data.astype(dt)
    A = Matrix.from_scipy(S, typ)
    k = int(rng.integers(-3, 4))
    _csr_equal(A.tril(k), _keep(S, lambda i, j, x: j - i <= k))
    _csr_equal(A.triu(k), _keep(S, lambda i, j, x: j - i >= k))
    _csr_equal(A.offdiag(), _keep(S, lambda i, j, x: i != j))
    _csr_equal(A.nonzero(), _keep(S, lambda i, j, x: x != 0))
    _csr_equal(A.select(">", 1), _keep(S, lambda i, j, x: x > 1))
    _csr_equal(A.select("<=0"), _keep(S, lambda i, j, x: x <= 0))
    _csr_equal(A.select(lib.GxB_TRIL, 0, out=Matrix.sparse(typ, n, m), desc=descriptor.T0), _keep(S.T.tocsr(), lambda i, j, x: j <= i))
    Z = S.copy(); Z.data = np.abs(Z.data)
    _csr_equal(A.apply(typ.ABS), Z)
    Z = S.copy(); Z.data = (Z.data * dt(3)).astype(dt)
    _csr_equal(A.apply_second(typ.TIMES, 3), Z)
    Z = S.copy(); Z.data = (dt(1) - Z.data).astype(dt)
    _csr_equal(A.apply_first(1, typ.MINUS), Z)
    # reductions
    assert A.reduce_int() == int(S.data.astype(np.int64).sum())
    if S.nnz:
        assert A.reduce_float(FP64.MAX_MONOID) == float(S.data.max()) and A.reduce_float(FP64.MIN_MONOID) == float(S.data.min())
    wide = np.int64 if np.dtype(dt).kind == "i" else np.float64
    rows = np.asarray(S.astype(wide).sum(axis=1)).ravel().astype(dt)          # integer sums wrap in the matrix type
    nz = np.diff(S.indptr) > 0
    w = A.reduce_vector()
    x, p = w.to_numpy()
    assert np.array_equal(p != 0, nz) and np.allclose(x[nz].astype(np.float64), rows[nz].astype(np.float64), rtol=1e-5)
    cols = np.asarray(S.astype(wide).sum(axis=0)).ravel().astype(dt)
    cz = np.diff(S.tocsc().indptr) > 0
    x, p = A.reduce_vector(desc=descriptor.T0).to_numpy()
    assert np.array_equal(p != 0, cz) and np.allclose(x[cz].astype(np.float64), cols[cz].astype(np.float64), rtol=1e-5)
    x, p = A.reduce_vector(typ.MAX_MONOID).to_numpy()
    assert np.array_equal(x[nz], np.array([S.data[S.indptr[r]:S.indptr[r + 1]].max() for r in np.nonzero(nz)[0]], dtype=dt))


@pytest.mark.parametrize("seed", range(12))
def test_random_eadd_emult(seed):
    rng = np.random.default_rng(6500 + seed)
    m, n = (int(rng.integers(1, 50)), int(rng.integers(1, 50))) if seed % 3 else (2000, 1500)
    typ, dt = [(INT64, np.int64), (FP64, np.float64), (INT8, np.int8)][seed % 3]
    dens = 0.3 if m < 100 else 0.01
    Sa, Sb = _rand(rng, m, n, dens, dt, 1, 6), _rand(rng, m, n, dens, dt, 1, 6)
    A, B = Matrix.from_scipy(Sa, typ), Matrix.from_scipy(Sb, typ)
    _csr_equal(A.eadd(B), (Sa + Sb).tocsr())
    pa, pb = Sa.copy(), Sb.copy(); pa.data[:] = 1; pb.data[:] = 1
    both = pa.multiply(pb).tocsr()                                   # pattern intersection
    _csr_equal(A.emult(B), Sa.multiply(Sb).tocsr().astype(dt))
    _csr_equal(A.emult(B, typ.FIRST), Sa.multiply(both).tocsr().astype(dt))
    # union with MAX: both present -> max, one present -> that one
    Mx = Sa.maximum(Sb).tocsr().astype(dt)
    _csr_equal(A.eadd(B, typ.MAX), Mx)
    # transposed second operand, accumulate into an existing C, structural mask
    St = _rand(rng, n, m, dens, dt, 1, 6)
    Bt = Matrix.from_scipy(St, typ)
    _csr_equal(A.emult(Bt, typ.TIMES, desc=descriptor.T1), Sa.multiply(St.T.tocsr()).tocsr().astype(dt))
    C0 = _rand(rng, m, n, dens, dt, 1, 6)
    C = Matrix.from_scipy(C0, typ)
    A.emult(B, typ.TIMES, out=C, accum=typ.PLUS)
    _csr_equal(C, (C0 + Sa.multiply(Sb)).tocsr().astype(dt))
    Mk = _rand(rng, m, n, dens, dt, 1, 2)
    C = Matrix.from_scipy(C0, typ)
    A.eadd(B, typ.MAX, out=C, mask=Matrix.from_scipy(Mk, typ), desc=descriptor.R)
    pm = Mk.copy(); pm.data[:] = 1
    _csr_equal(C, Mx.multiply(pm).tocsr().astype(dt))


def test_triangle_count_end_to_end():
    """demo/Triangle-Counting.ipynb:581-582: select -> masked mxm -> reduce, against scipy and the karate-club count."""
    from pygraphblas_b200.generators import rmat_csr
    n, indptr, indices = rmat_csr(13, 8, seed=4)
    S = sp.csr_matrix((np.ones(len(indices), np.int64), indices, indptr), shape=(n, n))
    S = ((S + S.T) > 0).astype(np.int64).tocsr(); S.setdiag(0); S.eliminate_zeros(); S.sort_indices()
    A = Matrix.from_scipy(S, INT64)
    Lm = A.tril(-1)
    C = Lm.mxm(Lm, mask=Lm, semiring=INT64.PLUS_PAIR)
    Ls = sp.tril(S, -1).tocsr()
    assert C.reduce_int() == int((Ls @ Ls).multiply(Ls).sum())
    # same count through a transposed second operand and through the upper triangle
    U = A.triu(1)
    assert Lm.mxm(U, mask=Lm, semiring=INT64.PLUS_PAIR, desc=descriptor.T1).reduce_int() == C.reduce_int()
    import networkx as nx
    K = nx.to_scipy_sparse_array(nx.karate_club_graph(), format="csr", dtype=np.int64)
    K.data[:] = 1
    Ak = Matrix.from_scipy(sp.csr_matrix(K), INT64)
    Lk = Ak.tril(-1)
    assert Lk.mxm(Lk, mask=Lk, semiring=INT64.PLUS_PAIR).reduce_int() == 45          # demo/Triangle-Counting.ipynb:33,56


# ------------------------------------------------------------------ GrB_Matrix_assign_<T> (matrix_assign.cu)
def _mdict(m):
    I, J, X = m.to_arrays()
    return {(int(i), int(j)): int(x) for i, j, x in zip(I, J, X)}


def _assign_scalar_model(C, M, accum, x, I, J, mask_struct, mask_comp, replace):
    """GrB_assign of a scalar, C API 1.3 section 4.3.7.6, on dicts: Z = C with the region I x J set to x (or accum(C, x));
    then the mask -- which spans all of C -- and GrB_REPLACE."""
    Z = dict(C)
    for k in itertools.product(I, J):
        Z[k] = (C[k] + x) if (accum and k in C) else x
    out = {}
    for k in set(C) | set(Z):
        if M is None:
            m = not mask_comp
        else:
            m = k in M and (mask_struct or M[k] != 0)
            m = (not m) if mask_comp else m
        if m:
            if k in Z:
                out[k] = Z[k]
        elif not replace and k in C:
            out[k] = C[k]
    return out


@pytest.mark.parametrize("seed", range(24))
def test_matrix_assign_scalar_against_model(seed):
    """C<M>(I,J) = accum(C(I,J), x) through the C ABI, over GrB_ALL / explicit lists / GxB_RANGE / GxB_STRIDE index forms,
    value / structural / complemented masks, GrB_REPLACE and a PLUS accumulator."""
    rng = np.random.default_rng(300 + seed)
    nr, nc = int(rng.integers(3, 14)), int(rng.integers(3, 14))
    def rand(dens):
        k = int(round(nr * nc * dens))
        flat = rng.choice(nr * nc, size=k, replace=False)
        return Matrix.from_lists(flat // nc, flat % nc, rng.integers(-3, 4, k), nr, nc, INT64)
    C = rand([0.0, 0.3, 0.6][seed % 3])
    M = rand(0.5) if seed % 2 else None
    dname = ["", "S", "C", "R", "RC", "RSC"][seed % 6]
    desc = getattr(descriptor, dname) if dname else None
    accum = INT64.PLUS if seed % 4 >= 2 else None
    ffi = gb.ffi
    form = seed % 4
    if form == 0:
        I, Iarg, ni = list(range(nr)), lib.GrB_ALL, 0
    elif form == 1:
        I = sorted(set(rng.integers(0, nr, 3).tolist())); Iarg = ffi.new("GrB_Index[]", I); ni = len(I)
    elif form == 2:
        lo, hi = sorted(rng.integers(0, nr, 2).tolist()); I = list(range(lo, hi + 1)); Iarg = ffi.new("GrB_Index[]", [lo, hi]); ni = lib.GxB_RANGE
    else:
        lo, hi = 0, nr - 1; I = list(range(lo, hi + 1, 2)); Iarg = ffi.new("GrB_Index[]", [lo, hi, 2]); ni = lib.GxB_STRIDE
    if seed % 3 == 0:
        J, Jarg, nj = list(range(nc)), lib.GrB_ALL, 0
    else:
        J = rng.integers(0, nc, 4).tolist(); Jarg = ffi.new("GrB_Index[]", J); nj = len(J)      # duplicates allowed
    expect = _assign_scalar_model(_mdict(C), _mdict(M) if M else None, accum is not None, 7, I, sorted(set(J)),
                                  "S" in dname, "C" in dname, "R" in dname)
    info = lib.GrB_Matrix_assign_INT64(C._matrix[0], M._matrix[0] if M else ffi.NULL, accum.get_op() if accum else ffi.NULL, 7,
                                       Iarg, ni, Jarg, nj, desc.get_desc() if desc else ffi.NULL)
    assert info == 0, ffi.string(lib.B200_last_error())
    assert _mdict(C) == expect, (seed, dname, form)


def test_dense_fill_like_the_reference_test_pow():
    """Matrix.dense (/root/reference/pygraphblas/matrix.py:183 -> assign_scalar :179) as tests/test_matrix.py:858-864 uses it."""
    m = Matrix.sparse(gb.UINT8, 10, 10)
    assert lib.GrB_Matrix_assign_UINT8(m._matrix[0], gb.ffi.NULL, gb.ffi.NULL, 1, lib.GrB_ALL, 0, lib.GrB_ALL, 0, gb.ffi.NULL) == 0
    assert m.nvals == 100 and set(m.to_arrays()[2].tolist()) == {1}
    assert (m @ m).iseq(m ** 2) and set((m ** 3).to_arrays()[2].tolist()) == {100}


# ------------------------------------------------------------------ GrB_Matrix_extract / GxB_*_diag / kronecker (matrix_assign.cu) against scipy
def _sp(m):
    I, J, X = m.to_arrays()
    return sp.csr_matrix((X.astype(np.float64), (I.astype(np.int64), J.astype(np.int64))), shape=(m.nrows, m.ncols))


def _rand_int_matrix(rng, nr, nc, dens):
    k = int(round(nr * nc * dens))
    flat = rng.choice(nr * nc, size=k, replace=False)
    return Matrix.from_lists(flat // nc, flat % nc, rng.integers(1, 9, k), nr, nc, INT64)


@pytest.mark.parametrize("seed", range(12))
def test_matrix_extract_against_scipy(seed):
    """C = A(I, J) for GrB_ALL, ranges, sorted lists, unsorted lists with duplicates, and INP0 = TRAN."""
    rng = np.random.default_rng(700 + seed)
    nr, nc = int(rng.integers(4, 30)), int(rng.integers(4, 30))
    A = _rand_int_matrix(rng, nr, nc, 0.3)
    tran = seed % 4 == 3
    S = _sp(A).T.tocsr() if tran else _sp(A)
    an, am = S.shape
    ffi = gb.ffi
    if seed % 3 == 0:
        I, Iarg, ni = np.arange(an), lib.GrB_ALL, 0
    elif seed % 3 == 1:
        I = rng.integers(0, an, 5); Iarg = ffi.new("GrB_Index[]", I.tolist()); ni = 5                  # duplicates allowed
    else:
        lo, hi = sorted(rng.integers(0, an, 2).tolist()); I = np.arange(lo, hi + 1); Iarg = ffi.new("GrB_Index[]", [lo, hi]); ni = lib.GxB_RANGE
    if seed % 4 == 0:
        J, Jarg, nj = np.arange(am), lib.GrB_ALL, 0
    elif seed % 4 == 1:
        J = np.sort(rng.choice(am, size=min(4, am), replace=False)); Jarg = ffi.new("GrB_Index[]", J.tolist()); nj = len(J)
    else:
        J = rng.integers(0, am, 6); Jarg = ffi.new("GrB_Index[]", J.tolist()); nj = 6                    # unsorted, duplicates
    C = Matrix.sparse(INT64, len(I), len(J))
    desc = descriptor.T0.get_desc() if tran else ffi.NULL
    assert lib.GrB_Matrix_extract(C._matrix[0], ffi.NULL, ffi.NULL, A._matrix[0], Iarg, ni, Jarg, nj, desc) == 0, ffi.string(lib.B200_last_error())
    ref = S[I][:, J].tocsr(); ref.sort_indices()
    got = _sp(C); got.sort_indices()
    assert (got != ref).nnz == 0 and got.nnz == ref.nnz


def test_diag_and_kronecker_against_scipy():
    rng = np.random.default_rng(9)
    v = Vector.from_lists([0, 2, 3], [5, 7, 9], 5, INT64)
    for k in (0, 2, -1):
        n = 5 + abs(k)
        C = Matrix.sparse(INT64, n, n)
        assert lib.GxB_Matrix_diag(C._matrix[0], v._vector[0], k, gb.ffi.NULL) == 0
        ref = sp.diags([np.array([5, 0, 7, 9, 0], np.float64)], [k], shape=(n, n)).tocsr(); ref.eliminate_zeros()
        assert (_sp(C) != ref).nnz == 0
        w = Vector.sparse(INT64, 5)
        assert lib.GxB_Vector_diag(w._vector[0], C._matrix[0], k, gb.ffi.NULL) == 0
        assert w.to_lists() == [[0, 2, 3], [5, 7, 9]]
    A, B = _rand_int_matrix(rng, 4, 5, 0.5), _rand_int_matrix(rng, 3, 6, 0.4)
    C = Matrix.sparse(INT64, 12, 30)
    assert lib.GrB_Matrix_kronecker_BinaryOp(C._matrix[0], gb.ffi.NULL, gb.ffi.NULL, INT64.TIMES.get_op(), A._matrix[0], B._matrix[0], gb.ffi.NULL) == 0
    assert (_sp(C) != sp.kron(_sp(A), _sp(B)).tocsr()).nnz == 0
