"""Tiny dict-based model of C<M> = accum(C, A (+).(x) B) -- TEST INFRASTRUCTURE ONLY.

A second, independent restatement of SURVEY.md section 8c steps 1-6 (pure Python, numpy
scalars for typed wrap-around) used to cross-check grb_oracle.c on small random cases.
Follows the same reference call sites (/root/reference/pygraphblas/matrix.py:2574, :2716,
vector.py:961).  Supports the operators the reference's tests and demos use.
"""
import numpy as np
import warnings

DT = {"BOOL": np.bool_, "INT8": np.int8, "INT16": np.int16, "INT32": np.int32, "INT64": np.int64,
      "UINT8": np.uint8, "UINT16": np.uint16, "UINT32": np.uint32, "UINT64": np.uint64,
      "FP32": np.float32, "FP64": np.float64}
CMP = {"EQ", "NE", "GT", "LT", "GE", "LE"}


def cast(v, to):
    dt = DT[to]
    if to == "BOOL":
        return np.bool_(v != 0)
    v = np.asarray(v)
    if v.dtype.kind == "f" and np.dtype(dt).kind in "iu":
        if np.isnan(v):
            return dt(0)
        info = np.iinfo(dt)
        if v <= info.min:
            return dt(info.min)
        if v >= info.max:
            return dt(info.max)
        return dt(int(v))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return v.astype(dt)[()]


def binop(op, typ, x, y):
    """z = op(x, y), x and y of type typ."""
    dt = DT[typ]
    x, y = dt(x), dt(y)
    if typ == "BOOL":
        table = {"FIRST": x, "ANY": x, "DIV": x, "SECOND": y, "RDIV": y, "PAIR": True,
                 "MIN": x and y, "TIMES": x and y, "LAND": x and y, "MAX": x or y, "PLUS": x or y, "LOR": x or y,
                 "MINUS": x != y, "RMINUS": x != y, "LXOR": x != y, "ISNE": x != y, "NE": x != y,
                 "ISEQ": x == y, "EQ": x == y, "GT": x and not y, "LT": (not x) and y, "GE": x or not y, "LE": (not x) or y}
        return np.bool_(table[op])
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        if op in ("FIRST", "ANY"):
            return x
        if op == "SECOND":
            return y
        if op == "PAIR":
            return dt(1)
        if op == "MIN":
            return np.fmin(x, y) if np.dtype(dt).kind == "f" else min(x, y)
        if op == "MAX":
            return np.fmax(x, y) if np.dtype(dt).kind == "f" else max(x, y)
        if op == "PLUS":
            return dt(x + y)
        if op == "MINUS":
            return dt(x - y)
        if op == "RMINUS":
            return dt(y - x)
        if op == "TIMES":
            return dt(x * y)
        if op in ("DIV", "RDIV"):
            a, b = (x, y) if op == "DIV" else (y, x)
            if np.dtype(dt).kind == "f":
                return dt(a / b)
            info = np.iinfo(dt)                      # GraphBLAS integer division: x/0 and INT_MIN/-1 are defined
            if b == 0:
                return dt(0) if a == 0 else dt(info.min if a < 0 else info.max)
            if np.dtype(dt).kind == "i" and b == -1:
                return dt(0) - a
            q = abs(int(a)) // abs(int(b))
            return dt(q if (int(a) < 0) == (int(b) < 0) else -q)
        if op == "BOR":
            return dt(x | y)
        if op == "BAND":
            return dt(x & y)
        if op == "BXOR":
            return dt(x ^ y)
        if op == "LOR":
            return dt((x != 0) or (y != 0))
        if op == "LAND":
            return dt((x != 0) and (y != 0))
        if op == "LXOR":
            return dt((x != 0) != (y != 0))
        if op in CMP:
            return np.bool_({"EQ": x == y, "NE": x != y, "GT": x > y, "LT": x < y, "GE": x >= y, "LE": x <= y}[op])
        if op.startswith("IS") and op[2:] in CMP:
            return dt({"EQ": x == y, "NE": x != y, "GT": x > y, "LT": x < y, "GE": x >= y, "LE": x <= y}[op[2:]])
    raise NotImplementedError(op)


def mxm(C, ctype, M, mtype, accum, semiring, A, atype, B, btype, desc):
    """All matrices are dicts {(i, j): value}.  desc: dict(replace, mask_comp, mask_struct, tran0, tran1)."""
    add, mul, mt = semiring
    zt = "BOOL" if mul in CMP else mt
    if desc["tran0"]:
        A = {(j, i): v for (i, j), v in A.items()}
    if desc["tran1"]:
        B = {(j, i): v for (i, j), v in B.items()}
    Brows = {}
    for (k, j), v in sorted(B.items()):
        Brows.setdefault(k, []).append((j, v))
    T = {}
    for (i, k), a in sorted(A.items()):
        for j, b in Brows.get(k, ()):
            p = binop(mul, mt, cast(a, mt), cast(b, mt))
            T[(i, j)] = binop(add, zt, T[(i, j)], p) if (i, j) in T else p
    if accum is None:
        Z = {k: cast(v, ctype) for k, v in T.items()}
    else:
        aop, at = accum
        azt = "BOOL" if aop in CMP else at
        Z = {}
        for k in set(C) | set(T):
            if k in C and k in T:
                Z[k] = cast(binop(aop, at, cast(C[k], at), cast(T[k], at)), ctype)
            elif k in C:
                Z[k] = C[k]
            else:
                Z[k] = cast(T[k], ctype)

    def m(k):
        if M is None:
            return not desc["mask_comp"]      # C<!NULL>: the complement of "no mask" lets nothing through (C API 1.3, 4.3)
        r = k in M and (desc["mask_struct"] or bool(cast(M[k], "BOOL")))
        return (not r) if desc["mask_comp"] else r

    out = {}
    for k in set(C) | set(Z):
        if m(k):
            if k in Z:
                out[k] = Z[k]
        elif not desc["replace"] and k in C:
            out[k] = C[k]
    return out
