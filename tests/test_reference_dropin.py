"""The drop-in boundary, exercised with the UNMODIFIED reference package (CPU container only: the
reference tree does not travel to the GPU box).  `suitesparse_graphblas/` at the repository root is the
binding stub of INTEGRATION.md; with it ahead on PYTHONPATH, /root/reference/pygraphblas imports and runs
on libb200grb.so.  Without a GPU every call that computes must refuse with Panic ("no CPU fallback") -- never
compute on the host -- while the handle plumbing around them works."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pygraphblas")), reason="reference tree not present")


def _run(code):
    env = dict(os.environ, PYTHONPATH=f"{ROOT}:{REF}")
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env, cwd="/tmp", timeout=600)


def test_unmodified_reference_imports_and_plumbing_works():
    r = _run("""
        import pygraphblas as gb
        from pygraphblas import Matrix, Vector, Scalar, INT64, BOOL, FP32, descriptor, lib
        assert gb.__file__.startswith("/root/reference/"), gb.__file__
        assert lib.GxB_IMPLEMENTATION_MAJOR == 5                    # base.py:37-46 version constants
        m = Matrix.from_lists([0, 1, 2], [1, 2, 0], [1, 2, 3])       # tests/test_matrix.py:250
        assert (m.nrows, m.ncols, m.nvals) == (3, 3, 3) and m.type is INT64
        assert m.to_lists() == [[0, 1, 2], [1, 2, 0], [1, 2, 3]]
        assert m.dup().to_lists() == m.to_lists()
        v = Vector.from_lists([0, 1, 2], [2, 3, 4])
        assert v.dup().to_lists() == [[0, 1, 2], [2, 3, 4]]
        assert INT64.PLUS_TIMES.ztype is INT64 and INT64.min_plus is INT64.MIN_PLUS and BOOL.LOR_LAND.ztype is BOOL
        assert descriptor.T1 in descriptor.CT1 and descriptor.CT1 == (descriptor.C & descriptor.T1)   # tests/test_descriptor.py:6-10
        assert Scalar.from_value(3)[0] == 3
        assert Matrix.sparse(INT64).nrows == 1 << 60                   # matrix.py:167-170
        print("HAVE_DEVICE", lib.B200_have_device())
        for call in (lambda: m.mxv(v), lambda: v.vxm(m), lambda: m.mxm(m), lambda: m @ m,
                     lambda: m.iseq(m.dup()), lambda: m.reduce_int(), lambda: v + v, lambda: m.tril(), lambda: v.apply(INT64.AINV)):
            try:
                out = call()
                assert lib.B200_have_device()
            except gb.base.Panic as e:
                assert not lib.B200_have_device() and b"no CPU fallback" in e.args[0]
        print("OK")
    """)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_reference_own_tests_run_against_the_library():
    """Run the reference's own unit tests.  Those that only need handle plumbing must pass; everything that
    computes may fail ONLY with the no-GPU Panic (or, for the entry points the library does not implement,
    with the stub's InvalidValue)."""
    env = dict(os.environ, PYTHONPATH=f"{ROOT}:{REF}")
    files = [f"{REF}/tests/test_{n}.py" for n in ("matrix", "vector", "descriptor", "scalar", "types", "base")]
    r = subprocess.run([sys.executable, "-m", "pytest", "-c", "/dev/null", "--rootdir", "/tmp", "-q", "-p", "no:cacheprovider"] + files,
                       capture_output=True, text=True, env=env, cwd="/tmp", timeout=900)
    out = r.stdout
    passed = int(__import__("re").search(r"(\d+) passed", out).group(1)) if " passed" in out else 0
    assert passed >= 35, out[-3000:]          # 40 of the reference's 121 tests need nothing but handle plumbing
    for line in out.splitlines():
        if line.startswith("FAILED") and any(k in line for k in ("test_mxm", "test_mxv", "test_vxm", "test_RC")):
            assert "Panic" in line, line
    # nothing computed on the host: every failure is a refusal, not a wrong answer
    bad = [l for l in out.splitlines() if l.startswith("FAILED") and "AssertionError" in l]
    assert not bad, bad


def test_slice_to_index_list_matches_the_reference():
    """Vector._index (the mirror) against the reference's own base._build_range (base.py:216-252) on a grid of
    slices and lists: same GrB_ALL / GxB_RANGE / GxB_STRIDE / GxB_BACKWARDS encoding, same result length."""
    r = _run("""
        import itertools
        import pygraphblas as ref
        from pygraphblas.base import _build_range, lib as rlib, ffi as rffi
        import pygraphblas_b200 as gb
        v = gb.Vector.sparse(gb.INT64, 10)
        vals = [None, 0, 1, 3, 8, 9]
        steps = [None, 1, 2, 3, -1, -2, -3]
        n = 0
        for a, b, c in itertools.product(vals, vals, steps):
            sl = slice(a, b, c)
            I0, ni0, sz0 = _build_range(sl, 9)
            I1, ni1, sz1 = v._index(sl, 10)
            if I0 == rlib.GrB_ALL:
                assert I1 == gb.lib.GrB_ALL and sz1 == 10, (sl,)
                continue
            assert int(ni0) == int(ni1), (sl, ni0, ni1)
            k = 2 if int(ni0) == int(rlib.GxB_RANGE) else 3
            assert [int(I0[q]) for q in range(k)] == [int(I1[q]) for q in range(k)], sl
            assert sz0 == sz1, (sl, sz0, sz1)
            n += 1
        I0, ni0, sz0 = _build_range([2, 3, 5, 7], 9)
        I1, ni1, sz1 = v._index([2, 3, 5, 7], 10)
        assert (ni0, sz0) == (ni1, sz1) == (4, 4) and [int(I1[q]) for q in range(4)] == I0
        print("OK", n)
    """)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_type_promotion_and_default_operators_match_the_reference():
    """types.promote (types.py:484-500), the default semiring / add / mult operators per type (types.py:156-160)
    and every semiring's ztype, mirror against the reference's own objects."""
    r = _run("""
        import pygraphblas as ref
        import pygraphblas_b200 as gb
        names = ["BOOL", "INT8", "INT16", "INT32", "INT64", "UINT8", "UINT16", "UINT32", "UINT64", "FP32", "FP64"]
        for a in names:
            for b in names:
                assert ref.types.promote(getattr(ref, a), getattr(ref, b)).__name__ == gb.types.promote(getattr(gb, a), getattr(gb, b)).name, (a, b)
            ra, ga = getattr(ref, a), getattr(gb, a)
            assert ra._default_semiring().name == ga._default_semiring().name, a
            assert ra._default_addop().name == ga._default_addop().name and ra._default_multop().name == ga._default_multop().name, a
        n = 0
        for name, sr in gb.ops.semirings.items():
            rs = getattr(getattr(ref, sr.type), f"{sr.pls}_{sr.mul}", None)
            if rs is None:
                continue
            assert rs.ztype.__name__ == sr.ztype.name, name
            n += 1
        assert n > 1000, n
        print("OK", n)
    """)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_output_inference_of_the_hot_calls_matches_the_reference():
    """Rows a1-a4: what Matrix.mxm / Matrix.mxv / Vector.vxm decide on the host before the FFI call -- type and
    shape of an implicit output, the semiring actually passed, accumulator / descriptor from the context managers
    (matrix.py:2553-2584, 2693-2726, vector.py:942-971, matrix.py:2380-2399).  Both packages run with the three
    hot entry points replaced by a recorder, so nothing computes; the recorded calls must agree."""
    r = _run("""
        import itertools
        import pygraphblas as ref
        import pygraphblas.matrix, pygraphblas.vector
        import pygraphblas_b200 as gb
        import pygraphblas_b200.matrix, pygraphblas_b200.vector

        class Recorder:
            def __init__(self, real, pkg):
                self._real, self._pkg, self.calls = real, pkg, []
            def __getattr__(self, name):
                if name in ("GrB_mxm", "GrB_mxv", "GrB_vxm"):
                    def rec(out, mask, accum, semiring, a, b, desc):
                        self.calls.append((name, out, mask != self._real_null(), accum, semiring, desc))
                        return 0
                    return rec
                return getattr(self._real, name)
            def _real_null(self):
                return self._pkg.base.NULL if hasattr(self._pkg, "base") else None

        rrec = Recorder(ref.lib, ref); grec = Recorder(gb.lib, gb)
        ref.matrix.lib = rrec; ref.vector.lib = rrec
        gb.matrix.lib = grec; gb.vector.lib = grec

        def name_of(pkg, kind, handle):
            ffi = pkg.ffi if hasattr(pkg, "ffi") else pkg.base.ffi
            if handle == ffi.NULL:
                return None
            p = ffi.new("char**")
            k = {"binop": 0, "semiring": 2}[kind]
            assert gb.lib.B200_object_name(gb.ffi.cast("const char**", p), k, gb.ffi.cast("void*", handle)) == 0
            return gb.ffi.string(gb.ffi.cast("char*", p[0])).decode()

        def desc_bits(pkg, d):
            ffi = pkg.ffi if hasattr(pkg, "ffi") else pkg.base.ffi
            if d == ffi.NULL:
                return (0, 0, 0, 0)
            out = []
            for f in (gb.lib.GrB_OUTP, gb.lib.GrB_MASK, gb.lib.GrB_INP0, gb.lib.GrB_INP1):
                v = gb.ffi.new("GrB_Desc_Value*")
                assert gb.lib.GxB_Desc_get(gb.ffi.cast("GrB_Descriptor", d), f, v) == 0
                out.append(int(v[0]))
            return tuple(out)

        def summarize(pkg, rec, result):
            name, out, has_mask, accum, semiring, desc = rec.calls[-1]
            typ = result.type.__name__ if hasattr(result.type, "__name__") else result.type.name
            shape = result.shape if hasattr(result, "nrows") else (result.size,)
            return (name, typ, tuple(int(x) for x in shape), has_mask, name_of(pkg, "binop", accum), name_of(pkg, "semiring", semiring), desc_bits(pkg, desc))

        types_ = ["BOOL", "INT8", "INT64", "UINT16", "FP32", "FP64"]
        n = 0
        for ta, tb in itertools.product(types_, types_):
            for variant in range(8):
                res = []
                for pkg, rec in ((ref, rrec), (gb, grec)):
                    A = pkg.Matrix.sparse(getattr(pkg, ta), 3, 5)
                    B = pkg.Matrix.sparse(getattr(pkg, tb), 5, 4)
                    Bt = pkg.Matrix.sparse(getattr(pkg, tb), 4, 5)
                    At = pkg.Matrix.sparse(getattr(pkg, ta), 5, 3)
                    u = pkg.Vector.sparse(getattr(pkg, tb), 5)
                    u3 = pkg.Vector.sparse(getattr(pkg, tb), 3)
                    T = getattr(pkg, ta)
                    d = pkg.descriptor
                    if variant == 0:
                        out = A.mxm(B)
                    elif variant == 1:
                        out = A.mxm(Bt, desc=d.T1, semiring=T.MIN_PLUS if ta != "BOOL" else T.LOR_LAND)
                    elif variant == 2:
                        out = A.mxv(u, cast=pkg.FP64)
                    elif variant == 3:
                        # square operand: with a descriptor that does NOT transpose, the reference sizes the implicit output
                        # by ncols (its Descriptor.__contains__ always answers True, descriptor.py:126-142) and then fails in
                        # GrB_mxv on a non-square A; the mirror sizes it correctly -- the one deliberate deviation
                        S = pkg.Matrix.sparse(getattr(pkg, ta), 5, 5)
                        out = S.mxv(u, accum=pkg.INT64.MIN, mask=pkg.Vector.sparse(pkg.BOOL, 5), desc=d.RC)
                    elif variant == 4:
                        out = u3.vxm(A)
                    elif variant == 5:
                        with (T.PLUS_PLUS if ta != "BOOL" else T.LOR_LOR), pkg.Accum(pkg.FP32.PLUS):
                            out = A @ B
                    elif variant == 6:
                        out = u.vxm(A, desc=d.T1, semiring=pkg.INT64.PLUS_PAIR)
                    else:
                        with d.S:
                            out = A.mxm(B, mask=pkg.Matrix.sparse(pkg.BOOL, 3, 4), cast=pkg.UINT8)
                    res.append(summarize(pkg, rec, out))
                assert res[0] == res[1], (ta, tb, variant, res)
                n += 1
        print("OK", n)
    """)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


FFI_CASES = r'''
MASKS={}
v=lambda p,t="INT64": p.Vector.from_lists([0,2,5],[1,2,3],8,getattr(p,t))
w=lambda p,t="INT64": p.Vector.from_lists([0,1,2,6],[1,2,3,4],8,getattr(p,t))
m=lambda p,t="INT64": p.Matrix.from_lists([0,1,2],[1,2,0],[1,2,3],4,4,getattr(p,t))
m2=lambda p,t="INT64": p.Matrix.from_lists([0,1,2,3,3],[1,2,0,0,3],[1,2,3,4,5],4,4,getattr(p,t))
bmask=lambda p: p.Vector.from_lists([0,2],[True,False],8,p.BOOL)
cases={
 "eadd": lambda p: v(p).eadd(w(p)), "eadd_max": lambda p: v(p).eadd(w(p), p.INT64.MAX), "eadd_mon": lambda p: v(p).eadd(w(p), p.INT64.PLUS_MONOID),
 "or": lambda p: v(p) | w(p), "add": lambda p: v(p) + w(p), "sub": lambda p: v(p) - w(p), "mul": lambda p: v(p) * w(p), "div": lambda p: v(p) / w(p),
 "and": lambda p: v(p) & w(p), "emult": lambda p: v(p).emult(w(p)), "emult_mixed": lambda p: v(p,"FP32").emult(v(p,"INT8")),
 "apply": lambda p: v(p).apply(p.INT64.AINV), "neg": lambda p: -v(p), "abs": lambda p: abs(v(p)), "inv": lambda p: ~v(p,"FP64"),
 "first": lambda p: v(p).apply_first(2, p.INT64.PLUS), "second": lambda p: v(p).apply_second(p.INT64.MINUS, 2),
 "add3": lambda p: v(p) + 3, "radd3": lambda p: 3 + v(p), "rsub": lambda p: 3 - v(p), "sub3": lambda p: v(p) - 3, "mul3": lambda p: v(p) * 3, "rmul": lambda p: 3 * v(p),
 "div3": lambda p: v(p) / 3, "rdiv": lambda p: 15 / v(p), "mulf": lambda p: v(p,"FP64") * 2.5,
 "iadd": lambda p: v(p).__iadd__(3), "iaddv": lambda p: v(p).__iadd__(w(p)), "isub": lambda p: v(p).__isub__(3), "isubv": lambda p: v(p).__isub__(w(p)),
 "imul": lambda p: v(p).__imul__(3), "imulv": lambda p: v(p).__imul__(w(p)), "idiv": lambda p: v(p).__itruediv__(3), "idivv": lambda p: v(p).__itruediv__(w(p)),
 "ior": lambda p: v(p).__ior__(w(p)), "iand": lambda p: v(p).__iand__(w(p)),
 "assign_scalar": lambda p: v(p).assign_scalar(3), "assign_scalar_f": lambda p: v(p).assign_scalar(2.5), "assign_scalar_mask": lambda p: v(p,"UINT8").assign_scalar(2, mask=MASKS.setdefault(p.__name__, bmask(p))),
 "setall": lambda p: v(p).__setitem__(slice(None), 3), "setslice": lambda p: v(p).__setitem__(slice(2,7), 1), "setmask": lambda p: v(p).__setitem__(MASKS.setdefault(p.__name__, bmask(p)), 4),
 "setvec": lambda p: v(p).__setitem__(slice(None), w(p)), "assign": lambda p: v(p).assign(w(p)), "getslice": lambda p: v(p)[1:5], "getstride": lambda p: v(p)[7:1:-2], "getlist": lambda p: v(p)[[2,3,5]],
 "reduce_int": lambda p: v(p).reduce_int(), "reduce_bool": lambda p: p.Vector.from_lists([0,2],[True,False],8,p.BOOL).reduce_bool(), "reduce_float": lambda p: v(p,"FP64").reduce_float(), "reduce_int_mon": lambda p: v(p).reduce_int(p.INT64.MAX_MONOID),
 "pattern": lambda p: v(p).pattern(), "pattern8": lambda p: v(p).pattern(p.INT8),
 "nonzero": lambda p: v(p).nonzero(), "select_gt": lambda p: v(p).select(">", 1),
 "m_tril": lambda p: m(p).tril(), "m_triu1": lambda p: m(p).triu(1), "m_offdiag": lambda p: m(p).offdiag(), "m_nonzero": lambda p: m(p).nonzero(),
 "m_select_gt": lambda p: m(p).select(">", 2), "m_select_op": lambda p: m(p).select(p.lib.GxB_TRIL, -1),
 "m_apply": lambda p: m(p).apply(p.INT64.ABS), "m_neg": lambda p: -m(p), "m_first": lambda p: m(p).apply_first(2, p.INT64.PLUS), "m_second": lambda p: m(p).apply_second(p.INT64.TIMES, 3),
 "m_reduce_int": lambda p: m(p).reduce_int(), "m_reduce_float": lambda p: m(p,"FP64").reduce_float(), "m_reduce_bool": lambda p: p.Matrix.from_lists([0,1],[1,0],[True,False],4,4,p.BOOL).reduce_bool(), "m_reduce_vector": lambda p: m(p).reduce_vector(),
 "m_reduce_vector_T0": lambda p: m(p).reduce_vector(desc=p.descriptor.T0),
 "m_eadd": lambda p: m(p).eadd(m2(p)), "m_add": lambda p: m(p) + m2(p), "m_sub": lambda p: m(p) - m2(p), "m_emult": lambda p: m(p).emult(m2(p)), "m_mul": lambda p: m(p) * m2(p), "m_div": lambda p: m(p) / m2(p),
 "m_or": lambda p: m(p) | m2(p), "m_and": lambda p: m(p) & m2(p), "m_pattern": lambda p: m(p).pattern(), "m_add3": lambda p: m(p) + 3, "m_mul3": lambda p: m(p) * 3,
 "m_iadd": lambda p: m(p).__iadd__(m2(p)), "m_isub": lambda p: m(p).__isub__(m2(p)), "m_imul": lambda p: m(p).__imul__(m2(p)), "m_idiv": lambda p: m(p).__itruediv__(m2(p)), "m_ior": lambda p: m(p).__ior__(m2(p)), "m_iand": lambda p: m(p).__iand__(m2(p)), "m_iadd3": lambda p: m(p).__iadd__(3), "m_isub3": lambda p: m(p).__isub__(3), "m_imul3": lambda p: m(p).__imul__(3), "m_idiv3": lambda p: m(p).__itruediv__(3), "m_radd": lambda p: 3 + m(p), "m_rsub": lambda p: 3 - m(p), "m_rmul": lambda p: 3 * m(p), "m_rdiv": lambda p: 12 / m(p), "m_inv": lambda p: ~m(p,"FP64"), "m_abs": lambda p: abs(m(p)),
 "m_transpose": lambda p: m(p).transpose(), "m_T": lambda p: m(p).T,
}

cases.update({
 # operators taken from context managers, string operators, Scalar operands, masks / accumulators / descriptors
 "ctx_binop": lambda p: _with(p.INT64.MAX, lambda: v(p) | w(p)), "ctx_binop_m": lambda p: _with(p.INT64.MIN, lambda: m(p) + m2(p)),
 "ctx_monoid": lambda p: _with(p.INT64.TIMES_MONOID, lambda: v(p).reduce_int()), "ctx_monoid_m": lambda p: _with(p.FP64.MAX_MONOID, lambda: m(p, "FP64").reduce_float()),
 "ctx_accum": lambda p: _with(p.Accum(p.INT64.MIN), lambda: v(p).eadd(w(p), out=v(p))),
 "str_op": lambda p: v(p).emult(w(p), "+"), "str_op2": lambda p: m(p).emult(m2(p), ">="),
 "scalar_first": lambda p: v(p).apply_first(MASKS.setdefault(p.__name__ + "s", p.Scalar.from_value(2)), p.INT8.PLUS),
 "scalar_second": lambda p: m(p).apply_second(p.INT8.MINUS, MASKS.setdefault(p.__name__ + "s", p.Scalar.from_value(2))),
 "eadd_full": lambda p: v(p).eadd(w(p), p.INT64.MIN, out=w(p), mask=MASKS.setdefault(p.__name__ + "2", bmask(p)), accum=p.INT64.PLUS, desc=p.descriptor.RSC),
 "m_eadd_T": lambda p: m(p).eadd(m2(p), p.INT64.MAX, desc=p.descriptor.T0T1), "m_select_desc": lambda p: m(p).select("<", 3, desc=p.descriptor.T0),
 "cast": lambda p: v(p).eadd(w(p), cast=p.FP32), "m_cast": lambda p: m(p).emult(m2(p, "FP32"), cast=p.FP64),
})
'''


def test_every_operation_makes_the_same_ffi_call_as_the_reference():
    """Rows (f)1 / (f)3 host logic: for ~110 user-level expressions on vectors and matrices (eadd / emult / apply /
    assign / extract / reduce / select / pattern / transpose, every arithmetic operator incl. the reflected and in-place
    forms, operators from `with` contexts and strings, Scalar operands, masks, accumulators, descriptors, casts) the
    mirror hands the C ABI the same function, operator handles, operands (in the same order), scalars, index lists and
    descriptor as the unmodified reference does.  Nothing computes: both run on the recorder of tests/ffi_recorder.py."""
    r = _run("""
        import sys
        sys.path.insert(0, %r)
        from ffi_recorder import *
        def _with(cm, fn):
            with cm:
                return fn()
        exec(%r)
        bad = []
        for k, fn in cases.items():
            a, b = both(fn)
            if a != b:
                bad.append((k, a, b))
            elif a[:1] == ("EXC",) and k != "m_inv":
                bad.append((k, "raises in both", a))
        assert not bad, bad
        print("OK", len(cases))
    """ % (os.path.join(ROOT, "tests"), FFI_CASES))
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
