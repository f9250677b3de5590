"""CPU: the dict model of the vector loop-glue operations (oracle/vecmodel.py, the checker of
tests/test_vector_ops_gpu.py) reproduces the known answers of the reference's own unit tests
(/root/reference/tests/test_vector.py, cited per case; inputs / outputs transcribed as data)."""
import numpy as np
from oracle import vecmodel as vm
from oracle.pymodel import DT

I64 = "INT64"


def vec(I, X, typ=I64):
    return {int(i): DT[typ](x) for i, x in zip(I, X)}


def lists(d):
    ks = sorted(d)
    return [ks, [d[k].item() for k in ks]]


def eadd(u, v, op="PLUS", typ=I64):
    T, zt = vm.ewise("add", op, typ, u, typ, v, typ)
    return vm.write({}, zt, None, None, T, zt, {})


def emult(u, v, op="TIMES", typ=I64):
    T, zt = vm.ewise("mult", op, typ, u, typ, v, typ)
    return vm.write({}, zt, None, None, T, zt, {})


def bind(u, op, s, first, typ=I64):
    T, zt = vm.bind(op, typ, s, "INT64", u, first)
    return vm.write({}, zt, None, None, T, zt, {})


def test_eadd_known_answers():
    """tests/test_vector.py:98-163"""
    V = list(range(2, 10))
    v = vec(V, V); v[0] = np.int64(1)
    w = vec(V, V); w[1] = np.int64(1)
    ref = vec(V, range(4, 20, 2)); ref[0] = np.int64(1); ref[1] = np.int64(1)
    assert eadd(v, w) == ref
    assert lists(eadd(v, w, "MINUS")) == [list(range(10)), [1, 1] + [0] * 8]
    idx = [0, 2, 3, 4, 5, 6, 7, 8, 9]
    assert lists(bind(v, "MINUS", 1, True)) == [idx, [0, -1, -2, -3, -4, -5, -6, -7, -8]]
    assert lists(bind(v, "MINUS", 1, False)) == [idx, [0, 1, 2, 3, 4, 5, 6, 7, 8]]
    assert lists(bind(v, "PLUS", 1, True)) == [idx, [2, 3, 4, 5, 6, 7, 8, 9, 10]]
    assert lists(eadd(v, v)) == [idx, [2, 4, 6, 8, 10, 12, 14, 16, 18]]


def test_emult_scalar_apply_known_answers():
    """tests/test_vector.py:166-195, 318-328, 414-420, 439-515, 553-559"""
    V = list(range(1, 11))
    v = vec(range(10), V)
    assert lists(emult(v, v))[1] == [x * x for x in V]
    assert lists(emult(v, v, "PLUS"))[1] == [x + x for x in V]
    assert lists(emult(v, v, "DIV"))[1] == [1] * 10
    f = vec([0, 1, 2], [2.0, 4.0, 8.0], "FP64")
    T, zt = vm.apply("AINV", I64, f)
    assert lists(vm.write({}, "FP64", None, None, T, zt, {})) == [[0, 1, 2], [-2.0, -4.0, -8.0]]
    T, zt = vm.apply("MINV", "FP64", f)
    assert lists(T) == [[0, 1, 2], [0.5, 0.25, 0.125]]
    u, w = vec([1], [5], "UINT64"), vec([1], [9], "UINT64")
    assert eadd(u, w, "BOR", "UINT64")[1] == 5 | 9
    m = vec([0, 1], [4, 2])
    T, zt = vm.bind("PLUS", "INT8", 2, "INT64", m, True)
    assert lists(vm.write({}, I64, None, None, T, zt, {})) == [[0, 1], [6, 4]]
    m = vec([0, 1], [5, 1])
    T, zt = vm.bind("MINUS", "INT8", 2, "INT64", m, False)
    assert lists(vm.write({}, I64, None, None, T, zt, {})) == [[0, 1], [3, -1]]
    assert lists(bind(m, "TIMES", 3, False)) == [[0, 1], [15, 3]]
    assert lists(bind(vec([0, 1], [15, 3]), "DIV", 3, False)) == [[0, 1], [5, 1]]
    assert lists(bind(vec([0, 1], [3, 5]), "DIV", 15, True)) == [[0, 1], [5, 3]]
    T, _ = vm.apply("AINV", I64, vec([0, 1], [0, 2]))
    assert lists(T) == [[0, 1], [0, -2]]
    T, _ = vm.apply("ABS", I64, vec([0, 1], [0, -2]))
    assert lists(T) == [[0, 1], [0, 2]]


def test_pattern_reduce_assign_known_answers():
    """tests/test_vector.py:198-240, 273-297"""
    v = vec([0, 2], [0, 42])
    T, _ = vm.apply("ONE", "BOOL", v)
    assert lists(T) == [[0, 2], [True, True]]
    T, _ = vm.apply("ONE", "INT8", v)
    assert lists(T) == [[0, 2], [1, 1]]
    assert not vm.reduce("LOR", "BOOL", {}, False) and vm.reduce("LOR", "BOOL", vec([3], [True], "BOOL"), False)
    assert vm.reduce("PLUS", I64, {}, 0) == 0 and vm.reduce("PLUS", I64, vec([3, 4], [3, 4]), 0) == 7
    assert vm.reduce("PLUS", "FP64", vec([3, 4], [3.3, 4.4], "FP64"), 0.0) == 7.7
    # v[:] = w ; v[1:] = w[9:1:-1] (inclusive stop) ; v[:] = 3 ; v[1:] = 0
    w = vec(range(10), range(10))
    v = vm.write({}, I64, None, None, w, I64, {})
    assert v == w
    src = {q: w[i] for q, i in enumerate(range(9, 0, -1))}                   # extract 9,8,...,1
    region = set(range(1, 10))
    v = vm.write(v, I64, None, None, {i: src[q] for q, i in enumerate(range(1, 10))}, I64, {}, region=region)
    assert lists(v) == [list(range(10)), [0, 9, 8, 7, 6, 5, 4, 3, 2, 1]]
    v = vm.write(v, I64, None, None, {i: np.int64(3) for i in range(10)}, I64, {})
    assert lists(v)[1] == [3] * 10
    v = vm.write(v, I64, None, None, {i: np.int64(0) for i in range(1, 10)}, I64, {}, region=set(range(1, 10)))
    assert lists(v)[1] == [3] + [0] * 9
    # masked scalar assign with replace: the notebook BFS's v.assign_scalar(level, mask=q)
    lv = vm.write(vec([0], [1], "UINT8"), "UINT8", (vec([2, 3], [True, False], "BOOL"), "BOOL"), None,
                  {i: np.uint8(2) for i in range(5)}, "UINT8", {})
    assert lists(lv) == [[0, 2], [1, 2]]
