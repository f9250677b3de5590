"""`suitesparse_graphblas.io` of the binding stub: the binary `.grb` reader / writer the reference imports
(/root/reference/pygraphblas/matrix.py:492, 937: `from suitesparse_graphblas.io import binary`)."""
from . import binary   # noqa: F401
