"""GPU parity of the vector loop-glue operations (vector_ops.cu) -- SURVEY.md section 8(f)1:
(1) the known answers of the reference's own unit tests (/root/reference/tests/test_vector.py, cited per
    case; inputs and expected outputs transcribed as data), run through the host-side mirror of its API;
(2) random cases over every builtin type with mask / accumulator / replace / index lists against the dict
    model oracle/vecmodel.py;  (3) a whole PageRank and a level-BFS written like the reference's
    (gap/prmark.py:8-30, demo/Introduction-to-GraphBLAS-with-Python.ipynb bfs) against numpy / scipy."""
import numpy as np
import pytest

import pygraphblas_b200 as gb
from pygraphblas_b200 import Matrix, Vector, BOOL, INT8, INT64, UINT8, UINT64, FP32, FP64, descriptor
from oracle import vecmodel as vm
from oracle.pymodel import DT
import util

pytestmark = pytest.mark.gpu


def L(v):
    I, X = v.to_arrays()
    return [I.tolist(), X.tolist()]


# ------------------------------------------------------------------ (1) the reference's known answers
def test_ref_vector_eadd():
    """tests/test_vector.py:98-163"""
    V = list(range(2, 10))
    v = Vector.from_lists(V, V); v[0] = 1
    w = Vector.from_lists(V, V); w[1] = 1
    ref = Vector.from_lists(V, list(range(4, 20, 2))); ref[0] = 1; ref[1] = 1
    s1 = v.eadd(w)
    assert s1.iseq(ref) and (v | w).iseq(s1)
    s3 = v.dup(); s3 |= w
    assert s3.iseq(s1)
    sub = Vector.from_list([1, 1] + [0] * 8)
    assert (v - w).iseq(sub)
    d2 = v.dup(); d2 -= w
    assert d2.iseq(sub)
    idx = [0, 2, 3, 4, 5, 6, 7, 8, 9]
    assert (1 - v).iseq(Vector.from_lists(idx, [0, -1, -2, -3, -4, -5, -6, -7, -8]))
    assert (v - 1).iseq(Vector.from_lists(idx, [0, 1, 2, 3, 4, 5, 6, 7, 8]))
    assert (1 + v).iseq(Vector.from_lists(idx, [2, 3, 4, 5, 6, 7, 8, 9, 10]))
    assert (v + 1).iseq(Vector.from_lists(idx, [2, 3, 4, 5, 6, 7, 8, 9, 10]))
    w = v.dup(); w -= 1
    assert w.iseq(Vector.from_lists(idx, [0, 1, 2, 3, 4, 5, 6, 7, 8]))
    w = v.dup(); w += 1
    assert w.iseq(Vector.from_lists(idx, [2, 3, 4, 5, 6, 7, 8, 9, 10]))
    w = v.dup(); w += v
    assert w.iseq(Vector.from_lists(idx, [2, 4, 6, 8, 10, 12, 14, 16, 18]))


def test_ref_vector_emult():
    """tests/test_vector.py:166-195"""
    V = list(range(1, 11))
    v, w = Vector.from_list(V), Vector.from_list(V)
    m1 = v.emult(w)
    assert m1.iseq(Vector.from_list([x * x for x in V])) and (v & w).iseq(m1)
    m3 = v.dup(); m3 &= w
    assert m3.iseq(m1)
    assert v.emult(w, INT64.PLUS).iseq(Vector.from_list([x + x for x in V]))
    assert (v / w).iseq(Vector.from_list([1] * 10))
    d2 = v.dup(); d2 /= w
    assert d2.iseq(Vector.from_list([1] * 10))
    w = v.dup(); w *= v
    assert w.iseq(Vector.from_lists(list(range(10)), [1, 4, 9, 16, 25, 36, 49, 64, 81, 100]))


def test_ref_pattern_reduce():
    """tests/test_vector.py:198-240"""
    v = Vector.sparse(INT64, 3); v[0] = 0; v[2] = 42
    p = v.pattern()
    assert p.type is BOOL and L(p) == [[0, 2], [True, True]]
    assert L(v.pattern(INT8)) == [[0, 2], [1, 1]]
    b = Vector.sparse(BOOL, 10)
    assert not b.reduce_bool()
    b[3] = True
    assert b.reduce_bool()
    i = Vector.sparse(INT64, 10)
    assert i.reduce_int() == 0 and type(i.reduce_int()) is int
    i[3] = 3; i[4] = 4
    assert i.reduce_int() == 7
    f = Vector.sparse(FP64, 10)
    assert f.reduce_float() == 0.0
    f[3] = 3.3; f[4] = 4.4
    assert f.reduce_float() == 7.7


def test_ref_slice_assign():
    """tests/test_vector.py:243-297, 518-526 (slices have an INCLUSIVE stop in the reference)"""
    v = Vector.from_list(list(range(10)))
    w = v[:9]
    assert w.size == 10 and L(w) == [list(range(10)), list(range(10))]
    w = v[1:8]
    assert w.size == 8 and L(w) == [list(range(8)), list(range(1, 9))]
    w = v[1:]
    assert w.size == 9 and L(w) == [list(range(9)), list(range(1, 10))]
    w = v[1:9:2]
    assert w.size == 5 and L(w) == [[0, 1, 2, 3, 4], [1, 3, 5, 7, 9]]
    w = v[7:1:-2]
    assert w.size == 4 and L(w) == [[0, 1, 2, 3], [7, 5, 3, 1]]
    w = v[[2, 3, 5, 7]]
    assert w.size == 4 and L(w) == [[0, 1, 2, 3], [2, 3, 5, 7]]
    v = Vector.sparse(INT64, 10)
    w = Vector.from_lists(list(range(10)), list(range(10)))
    v[:] = w
    assert v.iseq(w)
    v[1:] = w[9:1:-1]
    assert L(v) == [list(range(10)), [0, 9, 8, 7, 6, 5, 4, 3, 2, 1]]
    w[9:1:-1] = v[9:1:-1]
    assert w.iseq(v)
    v[:] = 3
    assert L(v) == [list(range(10)), [3] * 10]
    v[1:] = 0
    assert L(v) == [list(range(10)), [3] + [0] * 9]
    t = Vector.from_lists(list(range(10)), list(range(1, 11)))          # from_1_to_n(10)
    assert L(t[1:9:3]) == [[0, 1, 2], [2, 5, 8]] and L(t[9:1:-3]) == [[0, 1, 2], [10, 7, 4]]


def test_ref_apply_and_scalar_ops():
    """tests/test_vector.py:318-328, 414-420, 439-515, 553-559"""
    v = Vector.from_lists([0, 1, 2], [2.0, 4.0, 8.0])
    assert L(v.apply(INT64.AINV)) == [[0, 1, 2], [-2.0, -4.0, -8.0]]
    assert L(~v) == [[0, 1, 2], [0.5, 0.25, 0.125]]
    u, w = Vector.from_lists([1], [5], typ=UINT64), Vector.from_lists([1], [9], typ=UINT64)
    assert u.eadd(w, UINT64.BOR)[1] == 5 | 9
    m = Vector.from_lists([0, 1], [4, 2])
    assert L(m.apply_first(2, INT8.PLUS)) == [[0, 1], [6, 4]]
    m = Vector.from_lists([0, 1], [5, 1])
    assert L(m.apply_second(INT8.MINUS, 2)) == [[0, 1], [3, -1]]
    assert L(m + 3) == [[0, 1], [8, 4]] and L(3 + m) == [[0, 1], [8, 4]]
    assert L(m - 3) == [[0, 1], [2, -2]] and L(3 - m) == [[0, 1], [-2, 2]]
    assert L(m * 3) == [[0, 1], [15, 3]] and L(3 * m) == [[0, 1], [15, 3]]
    x = m.dup(); x += 3
    assert L(x) == [[0, 1], [8, 4]]
    x = m.dup(); x -= 3
    assert L(x) == [[0, 1], [2, -2]]
    x = m.dup(); x *= 3
    assert L(x) == [[0, 1], [15, 3]]
    d = Vector.from_lists([0, 1], [15, 3])
    assert L(d / 3) == [[0, 1], [5, 1]]
    assert L(15 / Vector.from_lists([0, 1], [3, 5])) == [[0, 1], [5, 3]]
    d /= 3
    assert L(d) == [[0, 1], [5, 1]]
    assert L(-Vector.from_lists([0, 1], [0, 2])) == [[0, 1], [0, -2]]
    assert L(abs(Vector.from_lists([0, 1], [0, -2]))) == [[0, 1], [0, 2]]


# ------------------------------------------------------------------ (2) random cases against the dict model
def _dict(d):
    dt = DT[d["type"]]
    return {int(i): dt(x) for i, x in zip(d["I"], d["X"])}


def _same(vec, model, typ):
    I, X = vec.to_arrays()
    assert vec.type.name == typ
    assert I.tolist() == sorted(model), (I.tolist(), sorted(model))
    exp = np.array([model[k] for k in sorted(model)], dtype=DT[typ])
    if typ in ("FP32", "FP64"):
        assert np.allclose(X, exp, rtol=1e-6, atol=0, equal_nan=True), (X, exp)
    else:
        assert np.array_equal(X, exp), (X, exp)


DESCS = [{}, {"replace": True}, {"mask_comp": True}, {"mask_struct": True}, {"replace": True, "mask_comp": True},
         {"replace": True, "mask_struct": True, "mask_comp": True}]
DNAME = ["", "R", "C", "S", "RC", "RSC"]
BINOPS = ["PLUS", "MINUS", "TIMES", "MIN", "MAX", "FIRST", "SECOND", "DIV", "LOR", "LAND", "EQ", "GT", "ISNE", "PAIR"]


@pytest.mark.parametrize("seed", range(60))
def test_random_ewise_apply_assign(seed):
    rng = np.random.default_rng(4000 + seed)
    n = int(rng.integers(1, 80)) if seed % 6 else int(rng.integers(3000, 9000))
    T = util.ALL_T
    ut, vt, wt, mt = (T[int(rng.integers(len(T)))] for _ in range(4))
    u = util.rand_vec(rng, ut, n, [1.0, 0.5, 0.1][seed % 3])
    v = util.rand_vec(rng, vt, n, [0.4, 1.0, 0.7][seed % 3])
    w = util.rand_vec(rng, wt, n, 0.5)
    mask = util.rand_vec(rng, mt, n, 0.5) if seed % 2 else None
    k = seed % len(DESCS) if mask is not None else 0
    dm, dg = DESCS[k], (getattr(descriptor, DNAME[k]) if DNAME[k] else None)
    accum = (["PLUS", "MIN", "SECOND", "TIMES"][seed % 4], wt) if seed % 3 == 1 else None
    gacc = getattr(util.g_type(wt), accum[0]) if accum else None
    mm = (_dict(mask), mt) if mask is not None else None
    gu, gv, gm = util.g_vec(u), util.g_vec(v), (util.g_vec(mask) if mask is not None else None)
    optype = T[int(rng.integers(len(T)))]
    op = BINOPS[seed % len(BINOPS)]
    gop = getattr(util.g_type(optype), op)
    kind = seed % 5
    gw = util.g_vec(w)
    if kind == 0:      # eWiseAdd
        gu.eadd(gv, gop, out=gw, mask=gm, accum=gacc, desc=dg)
        Tm, zt = vm.ewise("add", op, optype, _dict(u), ut, _dict(v), vt)
        exp = vm.write(_dict(w), wt, mm, accum, Tm, zt, dm)
    elif kind == 1:    # eWiseMult
        gu.emult(gv, gop, out=gw, mask=gm, accum=gacc, desc=dg)
        Tm, zt = vm.ewise("mult", op, optype, _dict(u), ut, _dict(v), vt)
        exp = vm.write(_dict(w), wt, mm, accum, Tm, zt, dm)
    elif kind == 2:    # apply: unary, bind first / second
        sub = seed % 3
        if sub == 0:
            uop = ["IDENTITY", "AINV", "ABS", "LNOT", "ONE", "MINV"][seed % 6]
            gu.apply(getattr(util.g_type(optype), uop), out=gw, mask=gm, accum=gacc, desc=dg)
            Tm, zt = vm.apply(uop, optype, _dict(u))
        else:
            # the bound scalar travels through the typed entry point of the VECTOR's type (vector.py:1293-1299, 1345-1351)
            if ut == "BOOL":
                s = bool(rng.integers(0, 2))
            elif ut in ("FP32", "FP64"):
                s = float(rng.integers(-6, 7)) / 2
            elif ut.startswith("U"):
                s = int(rng.integers(0, 4))
            else:
                s = int(rng.integers(-3, 4))
            if sub == 1:
                gu.apply_first(s, gop, out=gw, mask=gm, accum=gacc, desc=dg)
            else:
                gu.apply_second(gop, s, out=gw, mask=gm, accum=gacc, desc=dg)
            Tm, zt = vm.bind(op, optype, DT[ut](s), ut, _dict(u), first=(sub == 1))
        exp = vm.write(_dict(w), wt, mm, accum, Tm, zt, dm)
    elif kind == 3:    # assign scalar / vector over GrB_ALL, a stride and an index list
        sub = seed % 3
        idx = None if sub == 0 else (list(range(1, n, 3)) if sub == 1 else sorted(rng.choice(n, max(1, n // 4), replace=False).tolist()))
        index = None if sub == 0 else (slice(1, n - 1, 3) if sub == 1 else idx)
        region = None if idx is None else set(idx)
        if sub == 1 and n < 2:
            pytest.skip("stride needs two positions")
        if seed % 2:
            s = int(rng.integers(-3, 4))
            gw.assign_scalar(s, index, mask=gm, accum=gacc, desc=dg)
            where = range(n) if idx is None else idx
            Tm, zt = {int(i): DT[wt](vm.cast(np.int64(s), wt)) for i in where}, wt
        else:
            src = util.rand_vec(rng, ut, n if idx is None else len(idx), 0.6)
            gw.assign(util.g_vec(src), index, mask=gm, accum=gacc, desc=dg)
            sd = _dict(src)
            Tm, zt = ({k: x for k, x in sd.items()} if idx is None else {idx[k]: x for k, x in sd.items()}), ut
        exp = vm.write(_dict(w), wt, mm, accum, Tm, zt, dm, region=region)
    else:              # extract + reduce
        idx = sorted(rng.choice(n, max(1, n // 3), replace=False).tolist()) if seed % 2 else None
        out = gu.extract(idx)
        ud = _dict(u)
        _same(out, ud if idx is None else {q: ud[i] for q, i in enumerate(idx) if i in ud}, ut)
        for rt, mon, ident in (("INT64", "PLUS", 0), ("FP64", "MAX", -np.inf), ("BOOL", "LOR", False)):
            got = gu._reduce(util.g_type(rt), getattr(util.g_type(rt), mon + "_MONOID"))
            want = vm.reduce(mon, rt, ud, ident)
            assert got == want or (isinstance(got, float) and np.isclose(got, float(want))), (rt, got, want)
        return
    _same(gw, exp, wt)


# ------------------------------------------------------------------ (3) whole algorithms, written like the reference's
def test_pagerank_like_prmark():
    """gap/prmark.py:8-30 of the reference, line for line, on an R-MAT graph; against numpy power iteration."""
    from pygraphblas_b200.generators import rmat_csr
    import scipy.sparse as sp
    n, indptr, indices = rmat_csr(12, 8, seed=5)
    A = Matrix.from_csr(indptr, indices, np.ones(len(indices), np.float32), n, n, FP32)
    deg = np.diff(indptr).astype(np.float32)
    nz = np.nonzero(deg)[0]
    d = Vector.from_lists(nz.tolist(), deg[nz].tolist(), n, FP32)
    damping, itermax = 0.85, 30
    r = Vector.sparse(FP32, n); t = Vector.sparse(FP32, n)
    d.assign_scalar(damping, accum=FP32.DIV)
    r[:] = 1.0 / n
    teleport = (1 - damping) / n
    for _ in range(itermax):
        t, r = r, t
        w = t / d
        r[:] = teleport
        A.mxv(w, out=r, accum=FP32.PLUS, semiring=FP32.PLUS_SECOND, desc=descriptor.T0)
        t -= r
        t.apply(FP32.ABS, out=t)
        if t.reduce_float() <= 1e-7:
            break
    S = sp.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(n, n))
    dd = np.where(deg > 0, deg / damping, 1.0).astype(np.float64)
    x = np.full(n, 1.0 / n)
    for _ in range(itermax):
        wv = np.where(deg > 0, x / dd, 0.0)
        x = teleport + S.T @ wv
    got = r.to_numpy()[0].astype(np.float64)
    assert np.allclose(got, x, rtol=2e-4, atol=1e-9)


def test_level_bfs_like_the_notebook():
    """demo/Introduction-to-GraphBLAS-with-Python.ipynb `bfs`: levels in a UINT8 vector that is also the
    (complemented, valued) mask of the next vxm; reduce_bool ends the loop."""
    from pygraphblas_b200.generators import rmat_csr
    import scipy.sparse as sp
    from scipy.sparse.csgraph import breadth_first_order
    n, indptr, indices = rmat_csr(11, 8, seed=9)
    A = Matrix.from_csr(indptr, indices, None, n, n, BOOL)
    start = int(np.argmax(np.diff(indptr)))
    v = Vector.sparse(UINT8, n); q = Vector.sparse(BOOL, n)
    q[start] = True
    level = 1
    while q.reduce_bool() and level <= n:
        v.assign_scalar(level, mask=q)
        v.vxm(A, mask=v, out=q, desc=descriptor.RC, semiring=BOOL.LOR_LAND)
        level += 1
    S = sp.csr_matrix((np.ones(len(indices), np.bool_), indices, indptr), shape=(n, n))
    order, pred = breadth_first_order(S, start, directed=True, return_predecessors=True)
    ref = np.zeros(n, np.int64); ref[start] = 1
    for node in order[1:]:
        ref[node] = ref[pred[node]] + 1
    x, p = v.to_numpy()
    assert np.array_equal(np.nonzero(p)[0], np.nonzero(ref)[0])
    assert np.array_equal(x[p != 0].astype(np.int64), ref[ref > 0])
