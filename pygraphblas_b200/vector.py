"""Vector: host-side mirror of /root/reference/pygraphblas/vector.py for the hot path:
`Vector.vxm` (vector.py:835-971), `@` / `@=` (vector.py:973-977), `_get_args`
(vector.py:1078-1099) and the plumbing around it (sparse / from_lists / dup / nvals /
to_lists / element access / iseq / wait), plus dense import/export for bulk transfers."""
from functools import partial
import numpy as np

from .base import lib, ffi, NULL, _check, NoValue
from . import types
from .ops import current_semiring, current_accum
from .descriptor import current_desc, T1 as _T1

GxB_INDEX_MAX = 1 << 60


class Vector:
    __slots__ = ("_vector", "__weakref__")

    def __init__(self, handle):
        self._vector = handle

    def __del__(self):
        try:        # at interpreter shutdown the binding may already be torn down
            if lib is not None and getattr(self, "_vector", None) is not None:
                lib.GrB_Vector_free(self._vector)
        except Exception:
            pass

    # ------------------------------------------------------------------ construction
    @classmethod
    def sparse(cls, typ, size=None):
        """Empty vector; size defaults to GxB_INDEX_MAX (vector.py:250-286)."""
        v = ffi.new("GrB_Vector*")
        _check(lib.GrB_Vector_new(v, typ.gb_type, GxB_INDEX_MAX if size is None else size))
        return cls(v)

    @classmethod
    def from_lists(cls, I, V, size=None, typ=None):
        """(vector.py:330-356); later duplicates win."""
        I = np.ascontiguousarray(I, dtype=np.uint64)
        if typ is None:
            typ = types.from_python(V[0]) if len(V) else types.FP64
        if size is None:
            size = int(I.max()) + 1 if len(I) else 1
        X = np.ascontiguousarray(V, dtype=typ.dtype)
        v = cls.sparse(typ, size)
        _check(typ._Vector_build(v._vector[0], ffi.cast("GrB_Index*", I.ctypes.data), ffi.cast(typ.ptr, X.ctypes.data), len(I), NULL))
        return v

    @classmethod
    def from_list(cls, V, typ=None):
        return cls.from_lists(list(range(len(V))), V, len(V), typ)

    @classmethod
    def from_numpy(cls, x, present=None, typ=None):
        """Dense import straight into HBM (B200 extension `B200_Vector_import_dense`)."""
        x = np.ascontiguousarray(x)
        typ = typ or types._dtype_lookup(x.dtype)
        x = np.ascontiguousarray(x, dtype=typ.dtype)
        p = NULL
        if present is not None:
            present = np.ascontiguousarray(present, dtype=np.uint8)
            p = ffi.cast("uint8_t*", present.ctypes.data)
        v = ffi.new("GrB_Vector*")
        _check(lib.B200_Vector_import_dense(v, typ.gb_type, len(x), ffi.cast("void*", x.ctypes.data), p, 0))
        return cls(v)

    def set_numpy(self, x, present=None):
        """Refill this vector from a dense host array without reallocating in HBM."""
        typ = self.type
        x = np.ascontiguousarray(x, dtype=typ.dtype)
        p = NULL
        if present is not None:
            present = np.ascontiguousarray(present, dtype=np.uint8)
            p = ffi.cast("uint8_t*", present.ctypes.data)
        _check(lib.B200_Vector_set_dense(self._vector[0], ffi.cast("void*", x.ctypes.data), p, 0))

    def to_numpy(self, out=None, present_out=None):
        """(values, present) dense host arrays (B200 extension `B200_Vector_export_dense`)."""
        typ = self.type
        n = self.size
        x = np.empty(n, typ.dtype) if out is None else out
        p = np.empty(n, np.uint8) if present_out is None else present_out
        _check(lib.B200_Vector_export_dense(self._vector[0], ffi.cast("void*", x.ctypes.data), ffi.cast("uint8_t*", p.ctypes.data), 0))
        return x, p

    def device_ptrs(self):
        """(values_ptr, present_ptr or 0): raw CUDA device pointers of the dense HBM form."""
        vals = ffi.new("void**")
        pres = ffi.new("uint8_t**")
        _check(lib.B200_Vector_device_ptrs(self._vector[0], vals, pres))
        return int(ffi.cast("uintptr_t", vals[0])), int(ffi.cast("uintptr_t", pres[0]))

    def dup(self):
        v = ffi.new("GrB_Vector*")
        _check(lib.GrB_Vector_dup(v, self._vector[0]))
        return Vector(v)

    # ------------------------------------------------------------------ properties
    @property
    def gb_type(self):
        t = ffi.new("GrB_Type*")
        _check(lib.GxB_Vector_type(t, self._vector[0]))
        return t[0]

    @property
    def type(self):
        return types.from_handle(self.gb_type)

    @property
    def size(self):
        n = ffi.new("GrB_Index*")
        _check(lib.GrB_Vector_size(n, self._vector[0]))
        return n[0]

    @property
    def shape(self):
        return (self.size,)

    @property
    def nvals(self):
        n = ffi.new("GrB_Index*")
        _check(lib.GrB_Vector_nvals(n, self._vector[0]))
        return n[0]

    def __len__(self):
        return self.nvals

    def clear(self):
        _check(lib.GrB_Vector_clear(self._vector[0]))

    def wait(self):
        _check(lib.GrB_Vector_wait(self._vector))

    # ------------------------------------------------------------------ element access
    def to_arrays(self):
        typ = self.type
        n = self.nvals
        I = np.empty(n, np.uint64)
        X = np.empty(n, typ.dtype)
        nv = ffi.new("GrB_Index*", n)
        _check(typ._Vector_extractTuples(ffi.cast("GrB_Index*", I.ctypes.data), ffi.cast(typ.ptr, X.ctypes.data), nv, self._vector[0]))
        return I, X

    def to_lists(self):
        I, X = self.to_arrays()
        return [I.tolist(), X.tolist()]

    def __iter__(self):
        I, X = self.to_lists()
        return iter(zip(I, X))

    def __getitem__(self, i):
        typ = self.type
        x = ffi.new(typ.ptr)
        res = typ._Vector_extractElement(x, self._vector[0], i)
        if res == lib.GrB_NO_VALUE:
            raise NoValue(f"no value at {i}")
        _check(res)
        return typ.from_value(x[0])

    def get(self, i, default=None):
        try:
            return self[i]
        except NoValue:
            return default

    def __setitem__(self, i, value):
        typ = self.type
        _check(typ._Vector_setElement(self._vector[0], typ.from_value(value), i))

    def __delitem__(self, i):
        _check(lib.GrB_Vector_removeElement(self._vector[0], i))

    def __contains__(self, i):
        return self.get(i) is not None

    def iseq(self, other):
        """(vector.py:213-235)"""
        if not isinstance(other, Vector) or self.type is not other.type or self.size != other.size:
            return False
        a, b = self.to_arrays(), other.to_arrays()
        return all(np.array_equal(x, y) for x, y in zip(a, b))

    def isne(self, other):
        return not self.iseq(other)

    # ------------------------------------------------------------------ the hot path
    def _get_args(self, mask=None, accum=None, desc=None):
        """(vector.py:1078-1099)"""
        if isinstance(mask, Vector):
            mask = mask._vector[0]
        else:
            mask = NULL
        if accum is None:
            accum = current_accum.get(NULL)
        if accum is not NULL:
            accum = accum.get_op()
        if desc is None:
            desc = current_desc.get(NULL)
        if desc is not NULL:
            desc = desc.get_desc()
        return mask, accum, desc

    def vxm(self, other, semiring=None, cast=None, out=None, mask=None, accum=None, desc=None):
        """Vector-matrix multiply w'<mask'> = accum(w', u' (+).(x) A)  (vector.py:835-971)."""
        from .matrix import Matrix
        if not isinstance(other, Matrix):
            raise TypeError("Right argument to vxm must be a Matrix.")
        if semiring is None:
            semiring = current_semiring.get(NULL)
        if out is None:
            d = desc if desc is not None else current_desc.get(None)
            new_dimension = other.nrows if (d is not None and _T1 in d) else other.ncols   # vector.py:946
            if cast is not None:
                typ = cast
            elif semiring is not NULL:
                typ = semiring.ztype
            else:
                typ = types.promote(self.type, other.type)
            out = Vector.sparse(typ, new_dimension)
        if semiring is NULL:
            semiring = out.type._default_semiring()
        mask, accum, desc = self._get_args(mask, accum, desc)
        _check(lib.GrB_vxm(out._vector[0], mask, accum, semiring.get_op(), self._vector[0], other._matrix[0], desc))
        return out

    def __matmul__(self, other):
        from .matrix import Matrix
        if isinstance(other, Matrix):
            return self.vxm(other)
        raise TypeError("Right argument to @ must be a Matrix.")

    def __imatmul__(self, other):
        return self.vxm(other, out=self)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            attr = getattr(self.type, name)
        except AttributeError:
            raise AttributeError(f"Vector has no attribute or type operator {name}")
        return partial(attr, self)

    def __repr__(self):
        return f"<Vector ({self.size} : {self.nvals}:{self.type.name})>"
