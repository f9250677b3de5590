"""CFFI (ABI mode) binding of libb200grb.so -- the counterpart of the third-party
``suitesparse_graphblas`` module the reference imports
(/root/reference/pygraphblas/__init__.py:248: ``lib, ffi, initialize, is_initialized``).

The cdef text is derived from include/b200grb.h (+ the generated operator and typed
headers) by dropping preprocessor lines, so the header stays the single source of truth
for the C ABI.
"""
import os
import re
from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_INCLUDE = os.path.join(_ROOT, "include")
LIB_PATH = os.path.join(_HERE, "libb200grb.so")


def header_cdef():
    """The C declarations of the ABI as one cffi-parsable string."""
    def read(name):
        with open(os.path.join(_INCLUDE, name)) as f:
            return f.read()
    text = read("b200grb.h")
    text = text.replace('#include "b200grb_ops.h"', read("b200grb_ops.h"))
    text = text.replace('#include "b200grb_typed.h"', read("b200grb_typed.h"))
    text = text.replace('#include "b200grb_compat.h"', read("b200grb_compat.h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("#define ") and len(s.split()) == 3 and re.fullmatch(r"-?\d+", s.split()[2]):
            out.append(line)                    # integer constants (GxB_INDEX_MAX, ...) are part of the ABI
            continue
        if s.startswith("#") or s.startswith('extern "C"') or s == "}":
            continue
        if s.startswith("extern const double "):   # ABI mode cannot read `const double` constants; as variables it can
            line = line.replace("extern const double ", "extern double ")
        out.append(line)
    return "\n".join(out)


def declared_symbols():
    """Every function / global object name the headers declare (used by the ABI test)."""
    cdef = header_cdef()
    names = set(re.findall(r"\b((?:GrB|GxB|B200)_\w+)\s*\(", cdef))
    for decl in re.findall(r"extern\s+(?:const\s+)?\w+\s+([^;]+);", cdef):
        for n in decl.split(","):
            names.add(n.strip().lstrip("*"))
    # enum constants and typedef names are not symbols
    enums = set(re.findall(r"\b((?:GrB|GxB)_[A-Za-z0-9_]+)\s*=\s*-?\d+", cdef))
    enums |= set(re.findall(r"#define\s+(\w+)", cdef))
    return sorted(n for n in names if n not in enums)


ffi = FFI()
ffi.cdef(header_cdef())

if not os.path.exists(LIB_PATH) or os.environ.get("B200GRB_BUILD"):
    # first import of a fresh checkout (or B200GRB_BUILD=1 after editing the sources): compile in-tree with nvcc
    try:
        from .build import build_library
        build_library(verbose=bool(os.environ.get("B200GRB_BUILD")))
    except Exception as e:                                   # no nvcc on this machine
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing and could not be built: {e}")
if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m pygraphblas_b200.build` "
        "(there is no CPU fallback for the CUDA library)")
lib = ffi.dlopen(LIB_PATH)

_initialized = False


def initialize(blocking=False, memory_manager="c"):
    """Same signature as suitesparse_graphblas.initialize (pygraphblas/__init__.py:251-256)."""
    global _initialized
    info = lib.GrB_init(lib.GrB_BLOCKING if blocking else lib.GrB_NONBLOCKING)
    if info != lib.GrB_SUCCESS:
        raise RuntimeError(f"GrB_init failed: {info}")
    _initialized = True


def is_initialized():
    return _initialized
