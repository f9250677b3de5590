"""GxB_Scalar wrapper (mirror of /root/reference/pygraphblas/scalar.py:11-90): the thunk of `select` and the
bound operand of apply_first / apply_second."""
from .base import lib, ffi, _check
from . import types


class Scalar:
    def __init__(self, handle, typ):
        self._scalar, self.type = handle, typ

    def __del__(self):
        try:
            lib.GxB_Scalar_free(self._scalar)
        except Exception:
            pass

    @classmethod
    def from_type(cls, typ):
        s = ffi.new("GxB_Scalar*")
        _check(lib.GxB_Scalar_new(s, typ.gb_type))
        return cls(s, typ)

    @classmethod
    def from_value(cls, value):
        typ = types.from_python(value)
        self = cls.from_type(typ)
        _check(getattr(lib, f"GxB_Scalar_setElement_{typ.name}")(self._scalar[0], typ.from_value(value)))
        return self

    @property
    def nvals(self):
        n = ffi.new("GrB_Index*")
        _check(lib.GxB_Scalar_nvals(n, self._scalar[0]))
        return n[0]

    def __getitem__(self, _):
        x = ffi.new(self.type.ptr)
        _check(getattr(lib, f"GxB_Scalar_extractElement_{self.type.name}")(x, self._scalar[0]))
        return self.type.from_value(x[0])
