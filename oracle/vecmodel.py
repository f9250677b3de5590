"""Dict model of the GraphBLAS vector "loop glue" operations -- TEST INFRASTRUCTURE ONLY.

Restates, for vectors held as {index: numpy scalar}, the C API 1.3 semantics of the calls the reference makes
from /root/reference/pygraphblas/vector.py: eWiseAdd (:604-735), eWiseMult (:737-833), apply (:1101-1178),
assign / assign_scalar (:1461-1524), extract (:1526-1560), reduce (:533-593), all finished by the standard
write-back  w<mask> = accum(w, T)  (GrB_assign applies it to the positions of
the index list only).  Operators and typecasts come from oracle/pymodel.py, which the golden vectors pin.
"""
import math
import warnings
import numpy as np
from .pymodel import DT, CMP, binop, cast


def ztype(op, typ):
    return "BOOL" if op in CMP else typ


def unop(op, typ, x):
    dt = DT[typ]
    x = dt(x)
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        if op == "IDENTITY":
            return x
        if op == "ONE":
            return dt(1)
        if op == "LNOT":
            return dt(not (x != 0))
        if typ == "BOOL":
            return {"AINV": x, "MINV": np.bool_(True), "ABS": x}[op]
        kind = np.dtype(dt).kind
        if op == "AINV":
            return dt(0) - x if kind != "f" else -x
        if op == "ABS":
            return x if kind == "u" else (dt(0) - x if x < 0 else x)
        if op == "MINV":
            if kind == "f":
                return dt(1) / x
            if x == 0:
                return dt(np.iinfo(dt).max)
            if kind == "i" and x == -1:
                return dt(-1)
            return dt(int(1 / int(x)))
        if op == "BNOT":
            return ~x
        f = {"SQRT": math.sqrt, "EXP": math.exp, "LOG": math.log, "SIN": math.sin, "COS": math.cos, "TANH": math.tanh,
             "FLOOR": math.floor, "CEIL": math.ceil, "EXPM1": math.expm1, "LOG1P": math.log1p, "ERF": math.erf}[op]
        return dt(f(float(x)))


def write(w, wtype, mask, accum, T, ttype, desc, region=None, n=None):
    """w<mask> = accum(w, T); region (set of indices or None) limits the assignment GrB_assign style."""
    if accum is None:
        Z = {k: cast(v, wtype) for k, v in T.items()}
        if region is not None:
            for k, v in w.items():
                if k not in region:
                    Z[k] = v
    else:
        aop, at = accum
        Z = dict(w)
        for k, t in T.items():
            if k in w:
                Z[k] = cast(binop(aop, at, cast(w[k], at), cast(t, at)), wtype)
            else:
                Z[k] = cast(t, wtype)

    def m(k):
        if mask is None:
            return not desc.get("mask_comp", False)      # w<!NULL>: nothing is let through
        mv, _ = mask
        r = k in mv and (desc.get("mask_struct", False) or bool(cast(mv[k], "BOOL")))
        return (not r) if desc.get("mask_comp", False) else r

    out = {}
    for k in set(w) | set(Z):
        if m(k):
            if k in Z:
                out[k] = Z[k]
        elif not desc.get("replace", False) and k in w:
            out[k] = w[k]
    return out


def ewise(mode, op, optype, u, ut, v, vt):
    zt = ztype(op, optype)
    T = {}
    for k in set(u) | set(v):
        if k in u and k in v:
            T[k] = binop(op, optype, cast(u[k], optype), cast(v[k], optype))
        elif mode == "add":
            T[k] = cast(u[k], zt) if k in u else cast(v[k], zt)
    return T, zt


def apply(op, optype, u):
    zt = "BOOL" if op in ("ISINF", "ISNAN", "ISFINITE") else optype
    return {k: unop(op, optype, cast(x, optype)) for k, x in u.items()}, zt


def bind(op, optype, scalar, stype, u, first):
    s = cast(scalar, optype) if stype != optype else DT[optype](scalar)
    zt = ztype(op, optype)
    if first:
        return {k: binop(op, optype, s, cast(x, optype)) for k, x in u.items()}, zt
    return {k: binop(op, optype, cast(x, optype), s) for k, x in u.items()}, zt


def reduce(op, optype, u, ident):
    acc = DT[optype](ident)
    for k in sorted(u):
        acc = binop(op, optype, acc, cast(u[k], optype))
    return acc
