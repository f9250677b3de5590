#!/usr/bin/env python
"""Transcribe the reference's own golden vectors for the mxm / mxv / vxm path.

The reference's arithmetic lives in SuiteSparse:GraphBLAS, which is not installable in
this container, so the reference cannot be *run* to generate
fixtures.  What it does hold are literal expected values in its unit tests and doctests
(real SuiteSparse outputs recorded by the reference's authors).  This script writes them
verbatim -- inputs and expected outputs, each with the file:line it was read from --
to tests/golden/reference_goldens.json.  tests/test_oracle.py pins the CPU oracle to
them; tests/test_parity_gpu.py pins the CUDA path to them.

    python tests/golden/make_goldens.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def mat(typ, nrows, ncols, I, J, X):
    return {"type": typ, "nrows": nrows, "ncols": ncols, "I": I, "J": J, "X": X}


def vec(typ, size, I, X):
    return {"type": typ, "size": size, "I": I, "X": X}


# operands used throughout the reference's mxm/mxv/vxm tests and doctests
M3 = mat("INT64", 3, 3, [0, 1, 2], [1, 2, 0], [1, 2, 3])     # matrix.py:2421, test_matrix.py:250
N3 = mat("INT64", 3, 3, [0, 1, 2], [1, 2, 0], [2, 3, 4])     # matrix.py:2422, test_matrix.py:251
V3 = vec("INT64", 3, [0, 1, 2], [2, 3, 4])                   # matrix.py:2609, test_matrix.py:295
R3 = mat("INT64", 3, 3, [0, 1, 2], [2, 0, 1], [3, 8, 6])     # test_matrix.py:256 (= m @ n)
M43 = mat("INT64", 4, 3, [0, 1, 2, 3], [1, 2, 0, 1], [1, 2, 3, 4])    # test_matrix.py:294
M34 = mat("INT64", 3, 4, [0, 1, 2, 0], [1, 2, 0, 3], [1, 2, 3, 4])    # test_vector.py:299
MB = mat("BOOL", 3, 3, [0, 1, 2], [1, 2, 0], [True, True, True])      # test_descriptor.py:14
PT = ["PLUS", "TIMES", "INT64"]


def transpose(m):
    return mat(m["type"], m["ncols"], m["nrows"], m["J"], m["I"], m["X"])


cases = []


def mxm(id_, src, A, B, expect, semiring=PT, C=None, mask=None, accum=None, desc="", out_type="INT64"):
    cases.append({"id": id_, "source": src, "op": "mxm", "A": A, "B": B, "C": C or mat(out_type, A["nrows"] if "T0" not in desc else A["ncols"], expect["ncols"], [], [], []),
                  "mask": mask, "accum": accum, "semiring": semiring, "desc": desc, "expect": expect})


def mxv(id_, src, A, u, expect, semiring=PT, w=None, mask=None, accum=None, desc="", out_type="INT64"):
    cases.append({"id": id_, "source": src, "op": "mxv", "A": A, "u": u, "w": w or vec(out_type, expect["size"], [], []),
                  "mask": mask, "accum": accum, "semiring": semiring, "desc": desc, "expect": expect})


def vxm(id_, src, u, A, expect, semiring=PT, w=None, mask=None, accum=None, desc="", out_type="INT64"):
    cases.append({"id": id_, "source": src, "op": "vxm", "A": A, "u": u, "w": w or vec(out_type, expect["size"], [], []),
                  "mask": mask, "accum": accum, "semiring": semiring, "desc": desc, "expect": expect})


# ---------------------------------------------------------------- tests/test_matrix.py
mxm("test_mxm", "tests/test_matrix.py:249-257", M3, N3, R3)
mxm("test_mxm_imatmul_alias", "tests/test_matrix.py:258-259 (m @= n: C aliases A)", M3, N3, R3, C=M3)
mxm("test_mxm_lor_land_typecast", "tests/test_matrix.py:260-261 (INT64 operands cast to BOOL)", R3, N3,
    mat("BOOL", 3, 3, [0, 1, 2], [0, 1, 2], [True, True, True]), semiring=["LOR", "LAND", "BOOL"], out_type="BOOL")
mxm("test_mxm_context_plus_plus", "tests/test_matrix.py:268-271", M3, N3,
    mat("INT64", 3, 3, [0, 1, 2], [2, 0, 1], [4, 6, 5]), semiring=["PLUS", "PLUS", "INT64"])
mxm("test_mxm_context_lor_land", "tests/test_matrix.py:273-275", M3, N3,
    mat("BOOL", 3, 3, [0, 1, 2], [2, 0, 1], [True, True, True]), semiring=["LOR", "LAND", "BOOL"], out_type="BOOL")
mxv("test_mxv", "tests/test_matrix.py:293-297", M43, V3, vec("INT64", 4, [0, 1, 2, 3], [3, 8, 6, 12]))
mxv("test_mxv_transpose_T0", "tests/test_matrix.py:301 (m.transpose().mxv(v, desc=T0))", transpose(M43), V3,
    vec("INT64", 4, [0, 1, 2, 3], [3, 8, 6, 12]), desc="T0")
mxv("test_mxv_plus_plus", "tests/test_matrix.py:303-305", M43, V3, vec("INT64", 4, [0, 1, 2, 3], [4, 6, 5, 7]),
    semiring=["PLUS", "PLUS", "INT64"])
# ---------------------------------------------------------------- tests/test_vector.py
vxm("test_vxm", "tests/test_vector.py:298-303", V3, M34, vec("INT64", 4, [0, 1, 2, 3], [12, 2, 6, 8]))
vxm("test_vxm_mask", "tests/test_vector.py:305-306 (value mask j = {1: True})", V3, M34, vec("INT64", 4, [1], [2]),
    mask=vec("BOOL", 4, [1], [True]))
vxm("test_vxm_T1", "tests/test_vector.py:310 (v.vxm(m.transpose(), desc=T1))", V3, transpose(M34),
    vec("INT64", 4, [0, 1, 2, 3], [12, 2, 6, 8]), desc="T1")
vxm("test_vxm_plus_plus", "tests/test_vector.py:312-314", V3, M34, vec("INT64", 4, [0, 1, 2, 3], [7, 3, 5, 6]),
    semiring=["PLUS", "PLUS", "INT64"])
# ---------------------------------------------------------------- tests/test_descriptor.py (BFS step shape)
W0 = vec("BOOL", 3, [0], [True])
EMPTY_MASK = vec("BOOL", 3, [], [])
mxv("test_RCT0", "tests/test_descriptor.py:13-20 (out aliases input, empty complemented mask, replace, T0)", MB, W0,
    vec("BOOL", 3, [1], [True]), semiring=["LOR", "LAND", "BOOL"], w=W0, mask=EMPTY_MASK, desc="RCT0", out_type="BOOL")
mxv("test_RC", "tests/test_descriptor.py:23-30", MB, W0,
    vec("BOOL", 3, [2], [True]), semiring=["LOR", "LAND", "BOOL"], w=W0, mask=EMPTY_MASK, desc="RC", out_type="BOOL")
# ---------------------------------------------------------------- doctests pygraphblas/matrix.py (mxm)
mxm("doctest_mxm_default", "pygraphblas/matrix.py:2437-2443", M3, N3, R3)
mxm("doctest_mxm_accum_min_out", "pygraphblas/matrix.py:2464-2471 (o = m.dup(); o.mxm(n, accum=INT64.min, out=o))", M3, N3,
    mat("INT64", 3, 3, [0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1], [1, 3, 8, 2, 3, 6]), C=M3, accum=["MIN", "INT64"])
mxm("doctest_mxm_min_plus", "pygraphblas/matrix.py:2486-2492", M3, N3,
    mat("INT64", 3, 3, [0, 1, 2], [2, 0, 1], [4, 6, 5]), semiring=["MIN", "PLUS", "INT64"])
mxm("doctest_mxm_T0", "pygraphblas/matrix.py:2523-2529", M3, N3,
    mat("INT64", 3, 3, [0, 1, 2], [0, 1, 2], [12, 2, 6]), desc="T0")
mxm("doctest_mxm_cast_fp32", "pygraphblas/matrix.py:2544-2550 (cast=types.FP32: FP32 output, PLUS_TIMES_FP32)", M3, N3,
    mat("FP32", 3, 3, [0, 1, 2], [2, 0, 1], [3.0, 8.0, 6.0]), semiring=["PLUS", "TIMES", "FP32"], out_type="FP32")
# ---------------------------------------------------------------- doctests pygraphblas/matrix.py (mxv)
mxv("doctest_mxv_default", "pygraphblas/matrix.py:2610-2614", M3, V3, vec("INT64", 3, [0, 1, 2], [3, 8, 6]))
mxv("doctest_mxv_accum_plus_out", "pygraphblas/matrix.py:2628-2634 (o = v.dup(); m.mxv(v, accum=INT64.plus, out=o))", M3, V3,
    vec("INT64", 3, [0, 1, 2], [5, 11, 10]), w=V3, accum=["PLUS", "INT64"])
mxv("doctest_mxv_min_plus", "pygraphblas/matrix.py:2643-2647", M3, V3, vec("INT64", 3, [0, 1, 2], [4, 6, 5]),
    semiring=["MIN", "PLUS", "INT64"])
mxv("doctest_mxv_T0", "pygraphblas/matrix.py:2665-2669", M3, V3, vec("INT64", 3, [0, 1, 2], [12, 2, 6]), desc="T0")
mxv("doctest_mxv_mask", "pygraphblas/matrix.py:2678-2683 (mask = previous result with o[1] deleted)", M3, V3,
    vec("INT64", 3, [0, 2], [3, 6]), mask=vec("INT64", 3, [0, 2], [12, 6]))
mxv("doctest_mxv_cast_fp32", "pygraphblas/matrix.py:2685-2689", M3, V3, vec("FP32", 3, [0, 1, 2], [3.0, 8.0, 6.0]),
    semiring=["PLUS", "TIMES", "FP32"], out_type="FP32")
# ---------------------------------------------------------------- doctests pygraphblas/vector.py (vxm)
vxm("doctest_vxm_default", "pygraphblas/vector.py:857-861", V3, M3, vec("INT64", 3, [0, 1, 2], [12, 2, 6]))
vxm("doctest_vxm_accum_plus_out", "pygraphblas/vector.py:875-881", V3, M3, vec("INT64", 3, [0, 1, 2], [14, 5, 10]),
    w=V3, accum=["PLUS", "INT64"])
vxm("doctest_vxm_accum_min_imatmul", "pygraphblas/vector.py:882-888 (with Accum(INT64.min): o @= M)", V3, M3,
    vec("INT64", 3, [0, 1, 2], [2, 2, 4]), w=V3, accum=["MIN", "INT64"])
vxm("doctest_vxm_min_plus", "pygraphblas/vector.py:894-898", V3, M3, vec("INT64", 3, [0, 1, 2], [7, 3, 5]),
    semiring=["MIN", "PLUS", "INT64"])
vxm("doctest_vxm_T0_ignored", "pygraphblas/vector.py:922-926 (INP0 does not apply to the vector)", V3, M3,
    vec("INT64", 3, [0, 1, 2], [12, 2, 6]), desc="T0")
vxm("doctest_vxm_mask", "pygraphblas/vector.py:932-937", V3, M3, vec("INT64", 3, [0, 2], [12, 6]),
    mask=vec("INT64", 3, [0, 2], [12, 6]))
# ---------------------------------------------------------------- out_degree (plus_pair mxv with an iso vector)
mxv("doctest_out_degree", "pygraphblas/matrix.py:3548-3556 (self.cast(UINT64).plus_pair(Vector.iso(1, nrows)))",
    mat("UINT64", 3, 3, [0, 1, 0, 2], [1, 2, 2, 0], [42, 0, 3, 149]), vec("INT64", 3, [0, 1, 2], [1, 1, 1]),
    vec("UINT64", 3, [0, 1, 2], [2, 1, 1]), semiring=["PLUS", "PAIR", "UINT64"], out_type="UINT64")

# known answers
known = {
    "karate_triangles": {"source": "demo/Triangle-Counting.ipynb:33,56 (networkx.karate_club_graph)", "value": 45},
    "promotion": {"source": "tests/test_matrix.py:1017-1028", "cases": [["FP32", "FP64", "FP64"], ["FP32", "UINT8", "FP32"], ["INT8", "UINT8", "INT8"]]},
}

with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
    json.dump({"reference": "Graphegon/pygraphblas @ 2d89301", "cases": cases, "known_answers": known}, f, indent=1)
print(f"wrote {len(cases)} cases")


# tests/golden/test_binfile.grb.b64: the reference's binary fixture /root/reference/docs/test_binfile.grb (1021 bytes, the matrix of
# docs/test_mm.mm in SuiteSparse's BITMAP form), base64-encoded as a golden vector for the `.grb` reader:
#   python -c "import base64; print(base64.b64encode(open('/root/reference/docs/test_binfile.grb','rb').read()).decode())" > tests/golden/test_binfile.grb.b64
