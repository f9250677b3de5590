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
    assert passed >= 35, out[-3000:]          # 38 of the reference's 121 tests need nothing but handle plumbing
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
