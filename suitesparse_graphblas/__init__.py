"""Drop-in for the third-party ``suitesparse_graphblas`` module the reference imports
(/root/reference/pygraphblas/__init__.py:248, base.py:7):

    from suitesparse_graphblas import lib, ffi, initialize, is_initialized

Put this directory (the repository root) ahead of the real package on PYTHONPATH and the
UNMODIFIED reference package runs on libb200grb.so: ``Matrix.mxm`` / ``Matrix.mxv`` / ``Vector.vxm``
execute on the B200, the rest of the GraphBLAS API is present so that the import succeeds
(non-hot-path entry points either do host-side handle plumbing or refuse with GrB_INVALID_VALUE).
It is the binding stub INTEGRATION.md describes; nothing in the product's own tests or benchmark
depends on it.
"""
from pygraphblas_b200._ffi import ffi, lib, initialize, is_initialized   # noqa: F401

__all__ = ["lib", "ffi", "initialize", "is_initialized"]
