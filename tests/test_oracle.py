"""CPU tests that PIN THE ORACLE: against every golden vector the reference's tests hold for
the mxm/mxv/vxm path, against an independent pure-Python model, against scipy, and against
the known karate-club triangle count."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from oracle import pymodel
import util


def test_oracle_matches_reference_goldens(goldens):
    assert len(goldens["cases"]) >= 30
    for case in goldens["cases"]:
        res = util.oracle_run(case)
        ok = util.same_mat(res, case["expect"]) if case["op"] == "mxm" else util.same_vec(res, case["expect"])
        assert ok, f"{case['id']} ({case['source']}): got {res}, expected {case['expect']}"


def _dict(d):
    if "nrows" in d:
        return {(i, j): orc.DTYPES[d["type"]](x) for i, j, x in zip(d["I"], d["J"], d["X"])}
    return {(i, 0): orc.DTYPES[d["type"]](x) for i, x in zip(d["I"], d["X"])}


@pytest.mark.parametrize("seed", range(40))
def test_oracle_matches_python_model(seed):
    rng = np.random.default_rng(seed)
    typ = util.ALL_T[seed % len(util.ALL_T)]
    m, k, n = rng.integers(1, 9, 3)
    A = util.rand_mat(rng, typ, int(k) if seed % 3 == 1 else int(m), int(m) if seed % 3 == 1 else int(k), 0.4)
    B = util.rand_mat(rng, rng.choice(util.ALL_T), int(k), int(n), 0.4)
    desc = "T0" if seed % 3 == 1 else ""
    extra = ["", "C", "R", "RC", "S", "RSC", "SC", "RS"][seed % 8]
    desc = extra + desc
    ctype = rng.choice(util.ALL_T)
    C = util.rand_mat(rng, ctype, int(m), int(n), 0.3)
    M = util.rand_mat(rng, rng.choice(util.ALL_T), int(m), int(n), 0.5) if seed % 2 == 0 else None
    srs = util.semirings_for(typ)
    sr = srs[seed % len(srs)]
    ztype = orc.semiring_ztype(sr)
    accum = None
    if seed % 4 >= 2:
        accum = (["PLUS", "MIN", "MAX", "SECOND", "FIRST", "TIMES"][seed % 6], ctype)
    f = orc.parse_desc(desc)
    expect = pymodel.mxm(_dict(C), ctype, _dict(M) if M else None, M["type"] if M else None, accum, sr,
                         _dict(A), A["type"], _dict(B), B["type"], f)
    got = orc.mxm(util.o_mat(C), util.o_mat(M) if M else None, accum, sr, util.o_mat(A), util.o_mat(B), desc).todict()
    assert set(got) == set(expect), (seed, sr, desc, got, expect)
    for kk in got:
        assert got[kk] == expect[kk].item(), (seed, sr, desc, kk, got[kk], expect[kk])


@pytest.mark.parametrize("desc", ["C", "RC", "SC", "RSC"])
def test_oracle_complemented_null_mask(desc):
    """C<!NULL>: nothing is let through; REPLACE clears C (both restatements agree, C API 1.3 section 4.3)."""
    rng = np.random.default_rng(5)
    A, B, C = util.rand_mat(rng, "INT32", 6, 6, 0.5), util.rand_mat(rng, "INT32", 6, 6, 0.5), util.rand_mat(rng, "INT32", 6, 6, 0.4)
    sr = ("PLUS", "TIMES", "INT32")
    for accum in (None, ("PLUS", "INT32")):
        got = orc.mxm(util.o_mat(C), None, accum, sr, util.o_mat(A), util.o_mat(B), desc).todict()
        expect = pymodel.mxm(_dict(C), "INT32", None, None, accum, sr, _dict(A), "INT32", _dict(B), "INT32", orc.parse_desc(desc))
        assert set(got) == set(expect) == (set() if "R" in desc else set(_dict(C)))
        assert all(got[k] == _dict(C)[k] for k in got)


def test_oracle_plus_times_matches_scipy():
    # BASELINE.json configs[0]: 1024 x 1024 random 1% CSR, PLUS_TIMES_FP64 mxv
    A = sp.random(1024, 1024, density=0.01, format="csr", dtype=np.float64, random_state=0)
    A.sort_indices()
    u = np.random.default_rng(0).random(1024)
    coo = A.tocoo()
    Ao = orc.SpMat("FP64", 1024, 1024, coo.row, coo.col, coo.data)
    uo = orc.SpVec("FP64", 1024, np.arange(1024), u)
    w = orc.mxv(orc.SpVec("FP64", 1024), None, None, ("PLUS", "TIMES", "FP64"), Ao, uo)
    ref = A @ u
    nz = np.diff(A.indptr) > 0
    assert np.array_equal(w.I, np.nonzero(nz)[0])
    assert np.allclose(w.X, ref[nz], rtol=1e-12, atol=0)
    # SpGEMM against scipy too
    B = sp.random(300, 200, density=0.03, format="csr", dtype=np.float64, random_state=1)
    Cs = (A[:200, :300] @ B).tocoo()
    a2 = A[:200, :300].tocoo()
    b2 = B.tocoo()
    Co = orc.mxm(orc.SpMat("FP64", 200, 200), None, None, ("PLUS", "TIMES", "FP64"),
                 orc.SpMat("FP64", 200, 300, a2.row, a2.col, a2.data), orc.SpMat("FP64", 300, 200, b2.row, b2.col, b2.data))
    ref = sp.csr_matrix(Cs).todok()
    assert Co.nvals == len(ref)
    for (i, j), x in Co.todict().items():
        assert abs(ref[i, j] - x) <= 1e-12 * abs(x)


def test_oracle_karate_triangles(goldens):
    import networkx as nx
    G = nx.karate_club_graph()
    n = G.number_of_nodes()
    rows, cols = [], []
    for a, b in G.edges():
        lo, hi = min(a, b), max(a, b)
        rows.append(hi); cols.append(lo)            # strict lower triangle L
    L = orc.SpMat("INT64", n, n, rows, cols, np.ones(len(rows), np.int64))
    # demo/Triangle-Counting.ipynb:581-582: L.mxm(L, mask=L) with PLUS_PAIR, then reduce
    C = orc.mxm(orc.SpMat("INT64", n, n), L, None, ("PLUS", "PAIR", "INT64"), L, L, "")
    assert int(C.X.sum()) == goldens["known_answers"]["karate_triangles"]["value"] == 45
    # dot form C<L> = L * L' (descriptor ST1, demo/TriangleCentrality.ipynb:596) counts the same triangles
    C2 = orc.mxm(orc.SpMat("INT64", n, n), L, None, ("PLUS", "PAIR", "INT64"), L, L, "ST1")
    assert int(C2.X.sum()) == 45


def test_fast_kernels_match_oracle():
    """The typed OpenMP kernels used as the bench CPU baseline agree with the generic oracle."""
    import ctypes
    from pygraphblas_b200.generators import rmat_csr
    L = orc.lib()
    n, indptr, indices = rmat_csr(10, 8, seed=5)
    nnz = len(indices)
    rng = np.random.default_rng(3)
    vals = (rng.random(nnz, dtype=np.float32) + 0.5).astype(np.float32)
    u = rng.random(n, dtype=np.float32)
    w = np.zeros(n, np.float32); pres = np.zeros(n, np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.fast_spmv_plus_times_f32(ctypes.c_int64(n), p(indptr), p(indices), p(vals), p(u), p(w), p(pres))
    rows = np.repeat(np.arange(n), np.diff(indptr))
    Ao = orc.SpMat("FP32", n, n, rows, indices, vals)
    wo = orc.mxv(orc.SpVec("FP32", n), None, None, ("PLUS", "TIMES", "FP32"), Ao, orc.SpVec("FP32", n, np.arange(n), u))
    assert np.array_equal(np.nonzero(pres)[0], wo.I)
    assert np.allclose(w[pres != 0], wo.X, rtol=1e-5)
    # the partitioned / first-touch variant of the same kernel (bench.py reports the faster of the two)
    L.fast_spmv_plan_f32.restype = ctypes.c_void_p
    for nt in (1, 3, 8):
        plan = ctypes.c_void_p(L.fast_spmv_plan_f32(ctypes.c_int64(n), p(indptr), p(indices), p(vals), ctypes.c_int(nt)))
        w2 = np.full(n, -1, np.float32); pres2 = np.full(n, 9, np.uint8)
        L.fast_spmv_plan_run_f32(plan, p(u), p(w2), p(pres2))
        L.fast_spmv_plan_free_f32(plan)
        assert np.array_equal(pres2, pres) and np.allclose(w2, w, rtol=1e-6, atol=0), nt
    # masked plus_pair (triangle kernel), both formulations
    Ls = sp.tril(sp.csr_matrix((np.ones(nnz), indices, indptr), shape=(n, n)) + sp.csr_matrix((np.ones(nnz), indices, indptr), shape=(n, n)).T, -1).tocsr()
    Ls.sort_indices()
    lp, lj = Ls.indptr.astype(np.int64), Ls.indices.astype(np.uint32)
    cval = np.zeros(len(lj), np.int64); chas = np.zeros(len(lj), np.uint8)
    L.fast_masked_saxpy_plus_pair_i64(ctypes.c_int64(n), ctypes.c_int64(n), p(lp), p(lj), p(lp), p(lj), p(lp), p(lj), p(cval), p(chas))
    lrows = np.repeat(np.arange(n), np.diff(lp))
    Lo = orc.SpMat("INT64", n, n, lrows, lj, np.ones(len(lj), np.int64))
    Co = orc.mxm(orc.SpMat("INT64", n, n), Lo, None, ("PLUS", "PAIR", "INT64"), Lo, Lo, "S")
    assert int(cval.sum()) == int(Co.X.sum())
    assert int(chas.sum()) == Co.nvals
    cval2 = np.zeros(len(lj), np.int64); chas2 = np.zeros(len(lj), np.uint8)
    L.fast_masked_dot_plus_pair_i64(ctypes.c_int64(n), p(lp), p(lj), p(lp), p(lj), p(lp), p(lj), p(cval2), p(chas2))
    Cd = orc.mxm(orc.SpMat("INT64", n, n), Lo, None, ("PLUS", "PAIR", "INT64"), Lo, Lo, "ST1")
    assert int(cval2.sum()) == int(Cd.X.sum()) and int(chas2.sum()) == Cd.nvals
