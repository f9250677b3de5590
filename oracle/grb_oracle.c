/* grb_oracle.c -- CPU ORACLE for the mxm / mxv / vxm hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may
 * load this file's library; the product (pygraphblas_b200/, libb200grb.so) never does.
 *
 * What it restates.  The reference (Graphegon/pygraphblas) contains no arithmetic: its
 * hot path is three FFI calls into the third-party SuiteSparse:GraphBLAS
 *     lib.GrB_mxm   /root/reference/pygraphblas/matrix.py:2574
 *     lib.GrB_mxv   /root/reference/pygraphblas/matrix.py:2716
 *     lib.GrB_vxm   /root/reference/pygraphblas/vector.py:961
 * (dependency `suitesparse-graphblas`, unpinned in /root/reference/setup.py:21; the build
 * scripts pin SuiteSparse:GraphBLAS v5.1.1, /root/reference/build-wheels.sh:13).  Neither
 * is present in /root/reference or installable offline, so this file restates the
 * PUBLISHED algorithm those calls implement -- the GraphBLAS C API 1.3 definition of
 *     C<M> = accum(C, op(A) (+).(x) op(B))
 *  -- and is pinned against every golden vector the
 * reference's own tests and doctests hold for this path (tests/golden/reference_goldens.json,
 * transcribed from /root/reference/tests/test_matrix.py:249-306, test_vector.py:298-315,
 * test_descriptor.py:13-30 and the doctests at matrix.py:2421-2551, 2607-2689,
 * vector.py:854-938), plus scipy / networkx cross-checks (tests/test_oracle.py).
 *
 * Semantics, in order (each step cites where the reference relies on it):
 *  1. A' = INP0==TRAN ? A^T : A;  B' = INP1==TRAN ? B^T : B   (descriptor.T0/T1,
 *     /root/reference/pygraphblas/descriptor.py:150-182; doctest matrix.py:2521-2535).
 *  2. T(i,j) = (+)_k mul(cast_x A'(i,k), cast_y B'(k,j)) over k where BOTH are present;
 *     T(i,j) exists iff that set is non-empty; T has the monoid's type
 *     (semiring.ztype, /root/reference/pygraphblas/types.py:442-461).
 *  3. Z = accum ? (C union T, accum on the intersection) : T        (doctest matrix.py:2465-2472).
 *  4. mask m(i,j): none -> true; STRUCTURE -> present; else present && (bool)value; COMP negates
 *     (test_descriptor.py:13-30: empty complemented mask).
 *  5. where m: C(i,j) = cast_C Z(i,j) or deleted if Z has none; where !m: C kept, or
 *     deleted under OUTP==REPLACE.
 *  6. integers wrap; x/0 and casts follow GraphBLAS rules; FP MIN/MAX ignore NaN.
 *
 * Representation: matrices are row-major sorted COO (I, J, X) with X a typed array;
 * vectors are n x 1 (mxv) or 1 x n (vxm) matrices -- the Python wrapper does that mapping.
 */
#include <stdint.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

enum { T_BOOL = 0, T_INT8, T_INT16, T_INT32, T_INT64, T_UINT8, T_UINT16, T_UINT32, T_UINT64, T_FP32, T_FP64 };
enum { O_FIRST = 0, O_SECOND, O_PAIR, O_ANY, O_MIN, O_MAX, O_PLUS, O_MINUS, O_RMINUS, O_TIMES, O_DIV, O_RDIV,
       O_POW, O_ISEQ, O_ISNE, O_ISGT, O_ISLT, O_ISGE, O_ISLE, O_LOR, O_LAND, O_LXOR, O_BOR, O_BAND, O_BXOR,
       O_BXNOR, O_EQ, O_NE, O_GT, O_LT, O_GE, O_LE };

typedef union { bool b; int8_t i8; int16_t i16; int32_t i32; int64_t i64;
                uint8_t u8; uint16_t u16; uint32_t u32; uint64_t u64; float f32; double f64; } val_t;

static size_t tsize(int t) {
    switch (t) { case T_BOOL: case T_INT8: case T_UINT8: return 1; case T_INT16: case T_UINT16: return 2;
                 case T_INT32: case T_UINT32: case T_FP32: return 4; default: return 8; }
}
static val_t vload(int t, const void *p, size_t k) {
    val_t v; memset(&v, 0, sizeof v);
    memcpy(&v, (const char *)p + k * tsize(t), tsize(t));
    if (t == T_BOOL) v.b = (*((const uint8_t *)p + k)) != 0;
    return v;
}
static void vstore(int t, void *p, size_t k, val_t v) { memcpy((char *)p + k * tsize(t), &v, tsize(t)); }

/* ---- typecasting (GraphBLAS: C rules, except float->integer saturates, NaN -> 0) */
static long double to_real(val_t v, int t) {
    switch (t) { case T_BOOL: return v.b; case T_INT8: return v.i8; case T_INT16: return v.i16; case T_INT32: return v.i32;
                 case T_INT64: return (long double)v.i64; case T_UINT8: return v.u8; case T_UINT16: return v.u16;
                 case T_UINT32: return v.u32; case T_UINT64: return (long double)v.u64; case T_FP32: return v.f32; default: return v.f64; }
}
static uint64_t to_bits(val_t v, int t) {   /* two's complement image of an integer/bool value */
    switch (t) { case T_BOOL: return v.b; case T_INT8: return (uint64_t)(int64_t)v.i8; case T_INT16: return (uint64_t)(int64_t)v.i16;
                 case T_INT32: return (uint64_t)(int64_t)v.i32; case T_INT64: return (uint64_t)v.i64; case T_UINT8: return v.u8;
                 case T_UINT16: return v.u16; case T_UINT32: return v.u32; default: return v.u64; }
}
static bool is_fp(int t) { return t == T_FP32 || t == T_FP64; }
static int64_t sat_s(double d, int64_t lo, int64_t hi) { if (isnan(d)) return 0; if (d <= (double)lo) return lo; if (d >= (double)hi) return hi; return (int64_t)d; }
static uint64_t sat_u(double d, uint64_t hi) { if (isnan(d)) return 0; if (d <= 0) return 0; if (d >= (double)hi) return hi; return (uint64_t)d; }

static val_t vcast(val_t v, int from, int to) {
    val_t r; memset(&r, 0, sizeof r);
    if (from == to) return v;
    if (to == T_BOOL) { r.b = is_fp(from) ? (to_real(v, from) != 0) : (to_bits(v, from) != 0); return r; }
    if (to == T_FP32) {
        if (from == T_FP64) r.f32 = (float)v.f64; else if (from == T_INT64) r.f32 = (float)v.i64;
        else if (from == T_UINT64) r.f32 = (float)v.u64; else r.f32 = (float)to_real(v, from);
        return r;
    }
    if (to == T_FP64) {
        if (from == T_FP32) r.f64 = v.f32; else if (from == T_INT64) r.f64 = (double)v.i64;
        else if (from == T_UINT64) r.f64 = (double)v.u64; else r.f64 = (double)to_real(v, from);
        return r;
    }
    if (is_fp(from)) {
        double d = from == T_FP32 ? (double)v.f32 : v.f64;
        switch (to) {
            case T_INT8: r.i8 = (int8_t)sat_s(d, INT8_MIN, INT8_MAX); break;
            case T_INT16: r.i16 = (int16_t)sat_s(d, INT16_MIN, INT16_MAX); break;
            case T_INT32: r.i32 = (int32_t)sat_s(d, INT32_MIN, INT32_MAX); break;
            case T_INT64: r.i64 = sat_s(d, INT64_MIN, INT64_MAX); break;
            case T_UINT8: r.u8 = (uint8_t)sat_u(d, UINT8_MAX); break;
            case T_UINT16: r.u16 = (uint16_t)sat_u(d, UINT16_MAX); break;
            case T_UINT32: r.u32 = (uint32_t)sat_u(d, UINT32_MAX); break;
            default: r.u64 = sat_u(d, UINT64_MAX); break;
        }
        return r;
    }
    uint64_t bits = to_bits(v, from);   /* integer -> integer: modular */
    switch (to) {
        case T_INT8: r.i8 = (int8_t)bits; break; case T_INT16: r.i16 = (int16_t)bits; break;
        case T_INT32: r.i32 = (int32_t)bits; break; case T_INT64: r.i64 = (int64_t)bits; break;
        case T_UINT8: r.u8 = (uint8_t)bits; break; case T_UINT16: r.u16 = (uint16_t)bits; break;
        case T_UINT32: r.u32 = (uint32_t)bits; break; default: r.u64 = bits; break;
    }
    return r;
}

/* ---- binary operators.  One macro body per C type; integer arithmetic is done in the
 *      unsigned type of the same width so that overflow wraps. */
#define INT_OPS(NAME, T, U, SIGNED, TMIN, TMAX)                                                    \
static val_t op_##NAME(int op, T x, T y, bool *isbool) {                                           \
    val_t r; memset(&r, 0, sizeof r); T z = 0; *isbool = false;                                     \
    switch (op) {                                                                                   \
        case O_FIRST: case O_ANY: z = x; break; case O_SECOND: z = y; break; case O_PAIR: z = 1; break; \
        case O_MIN: z = x < y ? x : y; break; case O_MAX: z = x > y ? x : y; break;                 \
        case O_PLUS: z = (T)((U)x + (U)y); break; case O_MINUS: z = (T)((U)x - (U)y); break;        \
        case O_RMINUS: z = (T)((U)y - (U)x); break; case O_TIMES: z = (T)((U)x * (U)y); break;      \
        case O_DIV: case O_RDIV: { T n = op == O_DIV ? x : y, d = op == O_DIV ? y : x;             \
            if (SIGNED && d == (T)-1) z = (T)((U)0 - (U)n);                                         \
            else if (d == 0) z = n == 0 ? 0 : ((SIGNED && n < 0) ? TMIN : TMAX);                    \
            else z = (T)(n / d); } break;                                                           \
        case O_POW: { double p = pow((double)x, (double)y);                                         \
            z = SIGNED ? (T)sat_s(p, (int64_t)TMIN, (int64_t)TMAX) : (T)sat_u(p, (uint64_t)TMAX); } break; \
        case O_ISEQ: z = x == y; break; case O_ISNE: z = x != y; break; case O_ISGT: z = x > y; break;  \
        case O_ISLT: z = x < y; break; case O_ISGE: z = x >= y; break; case O_ISLE: z = x <= y; break;  \
        case O_LOR: z = (x != 0) || (y != 0); break; case O_LAND: z = (x != 0) && (y != 0); break;  \
        case O_LXOR: z = (x != 0) != (y != 0); break;                                               \
        case O_BOR: z = (T)((U)x | (U)y); break; case O_BAND: z = (T)((U)x & (U)y); break;          \
        case O_BXOR: z = (T)((U)x ^ (U)y); break; case O_BXNOR: z = (T)~((U)x ^ (U)y); break;       \
        case O_EQ: *isbool = true; r.b = x == y; return r; case O_NE: *isbool = true; r.b = x != y; return r; \
        case O_GT: *isbool = true; r.b = x > y; return r;  case O_LT: *isbool = true; r.b = x < y; return r;  \
        case O_GE: *isbool = true; r.b = x >= y; return r; case O_LE: *isbool = true; r.b = x <= y; return r; \
        default: z = x;                                                                             \
    }                                                                                               \
    memcpy(&r, &z, sizeof z); return r;                                                             \
}
INT_OPS(i8, int8_t, uint8_t, 1, INT8_MIN, INT8_MAX)
INT_OPS(i16, int16_t, uint16_t, 1, INT16_MIN, INT16_MAX)
INT_OPS(i32, int32_t, uint32_t, 1, INT32_MIN, INT32_MAX)
INT_OPS(i64, int64_t, uint64_t, 1, INT64_MIN, INT64_MAX)
INT_OPS(u8, uint8_t, uint8_t, 0, 0, UINT8_MAX)
INT_OPS(u16, uint16_t, uint16_t, 0, 0, UINT16_MAX)
INT_OPS(u32, uint32_t, uint32_t, 0, 0, UINT32_MAX)
INT_OPS(u64, uint64_t, uint64_t, 0, 0, UINT64_MAX)

#define FP_OPS(NAME, T, FMIN, FMAX, POW)                                                            \
static val_t op_##NAME(int op, T x, T y, bool *isbool) {                                            \
    val_t r; memset(&r, 0, sizeof r); T z = 0; *isbool = false;                                      \
    switch (op) {                                                                                    \
        case O_FIRST: case O_ANY: z = x; break; case O_SECOND: z = y; break; case O_PAIR: z = 1; break;  \
        case O_MIN: z = FMIN(x, y); break; case O_MAX: z = FMAX(x, y); break;                        \
        case O_PLUS: z = x + y; break; case O_MINUS: z = x - y; break; case O_RMINUS: z = y - x; break;  \
        case O_TIMES: z = x * y; break; case O_DIV: z = x / y; break; case O_RDIV: z = y / x; break; \
        case O_POW: z = POW(x, y); break;                                                            \
        case O_ISEQ: z = x == y; break; case O_ISNE: z = x != y; break; case O_ISGT: z = x > y; break;   \
        case O_ISLT: z = x < y; break; case O_ISGE: z = x >= y; break; case O_ISLE: z = x <= y; break;   \
        case O_LOR: z = (x != 0) || (y != 0); break; case O_LAND: z = (x != 0) && (y != 0); break;   \
        case O_LXOR: z = (x != 0) != (y != 0); break;                                                \
        case O_EQ: *isbool = true; r.b = x == y; return r; case O_NE: *isbool = true; r.b = x != y; return r; \
        case O_GT: *isbool = true; r.b = x > y; return r;  case O_LT: *isbool = true; r.b = x < y; return r;  \
        case O_GE: *isbool = true; r.b = x >= y; return r; case O_LE: *isbool = true; r.b = x <= y; return r; \
        default: z = x;                                                                              \
    }                                                                                                \
    memcpy(&r, &z, sizeof z); return r;                                                              \
}
FP_OPS(f32, float, fminf, fmaxf, powf)
FP_OPS(f64, double, fmin, fmax, pow)

static val_t op_bool(int op, bool x, bool y) {
    val_t r; memset(&r, 0, sizeof r);
    switch (op) {   /* arithmetic names on BOOL alias logical operators */
        case O_FIRST: case O_ANY: case O_DIV: r.b = x; break;
        case O_SECOND: case O_RDIV: r.b = y; break;
        case O_PAIR: r.b = true; break;
        case O_MIN: case O_TIMES: case O_LAND: r.b = x && y; break;
        case O_MAX: case O_PLUS: case O_LOR: r.b = x || y; break;
        case O_MINUS: case O_RMINUS: case O_LXOR: case O_ISNE: case O_NE: r.b = x != y; break;
        case O_POW: r.b = x || !y; break;
        case O_ISEQ: case O_EQ: r.b = x == y; break;
        case O_ISGT: case O_GT: r.b = x && !y; break;
        case O_ISLT: case O_LT: r.b = !x && y; break;
        case O_ISGE: case O_GE: r.b = x || !y; break;
        case O_ISLE: case O_LE: r.b = !x || y; break;
        default: r.b = x;
    }
    return r;
}

/* z = op(x, y); x, y of type t; *zt receives the result type (t, or BOOL for comparisons) */
static val_t binop(int op, int t, val_t x, val_t y, int *zt) {
    bool isb = false; val_t r;
    switch (t) {
        case T_BOOL: r = op_bool(op, x.b, y.b); *zt = T_BOOL; return r;
        case T_INT8: r = op_i8(op, x.i8, y.i8, &isb); break;    case T_INT16: r = op_i16(op, x.i16, y.i16, &isb); break;
        case T_INT32: r = op_i32(op, x.i32, y.i32, &isb); break; case T_INT64: r = op_i64(op, x.i64, y.i64, &isb); break;
        case T_UINT8: r = op_u8(op, x.u8, y.u8, &isb); break;    case T_UINT16: r = op_u16(op, x.u16, y.u16, &isb); break;
        case T_UINT32: r = op_u32(op, x.u32, y.u32, &isb); break; case T_UINT64: r = op_u64(op, x.u64, y.u64, &isb); break;
        case T_FP32: r = op_f32(op, x.f32, y.f32, &isb); break;  default: r = op_f64(op, x.f64, y.f64, &isb); break;
    }
    *zt = isb ? T_BOOL : t;
    return r;
}

/* ---- sparse matrix in CSR built from row-major COO */
typedef struct { int64_t nrows, ncols, nvals; int type; int64_t *ptr; int64_t *col; val_t *val; } csr_t;

static void csr_free(csr_t *m) { free(m->ptr); free(m->col); free(m->val); memset(m, 0, sizeof *m); }

static int csr_from_coo(csr_t *m, int type, int64_t nrows, int64_t ncols, int64_t nvals,
                        const uint64_t *I, const uint64_t *J, const void *X, bool transpose) {
    m->type = type; m->nvals = nvals;
    m->nrows = transpose ? ncols : nrows; m->ncols = transpose ? nrows : ncols;
    m->ptr = calloc((size_t)m->nrows + 1, sizeof(int64_t));
    m->col = malloc((size_t)(nvals ? nvals : 1) * sizeof(int64_t));
    m->val = malloc((size_t)(nvals ? nvals : 1) * sizeof(val_t));
    if (!m->ptr || !m->col || !m->val) return -1;
    for (int64_t k = 0; k < nvals; ++k) m->ptr[(transpose ? J[k] : I[k]) + 1]++;
    for (int64_t r = 0; r < m->nrows; ++r) m->ptr[r + 1] += m->ptr[r];
    int64_t *cur = malloc((size_t)(m->nrows ? m->nrows : 1) * sizeof(int64_t));
    if (!cur) return -1;
    memcpy(cur, m->ptr, (size_t)m->nrows * sizeof(int64_t));
    /* input is row-major sorted, so a counting pass keeps columns ascending in either orientation */
    for (int64_t k = 0; k < nvals; ++k) {
        int64_t r = transpose ? (int64_t)J[k] : (int64_t)I[k], c = transpose ? (int64_t)I[k] : (int64_t)J[k];
        int64_t p = cur[r]++;
        m->col[p] = c; m->val[p] = vload(type, X, (size_t)k);
    }
    free(cur);
    return 0;
}

typedef struct {
    int replace, mask_comp, mask_struct, tran0, tran1;
} odesc_t;

static int cmp_i64(const void *p, const void *q) { int64_t a = *(const int64_t *)p, b = *(const int64_t *)q; return a < b ? -1 : a > b; }

/* result handle */
typedef struct { int64_t nvals; uint64_t *I; uint64_t *J; void *X; } ores_t;

void oracle_free_result(ores_t *r) { free(r->I); free(r->J); free(r->X); memset(r, 0, sizeof *r); }

/* C<M> = accum(C, A' (+).(x) B').  All matrices row-major sorted COO.  Returns 0 on success.
 *   has_mask / has_accum: flags; accum_op on type accum_t; semiring: add_op on ztype, mul_op on mul_t
 *   flip_mul: evaluate mul(B'(k,j), A'(i,k)) instead of mul(A'(i,k), B'(k,j))   (never set by the
 *   wrapper -- vxm is expressed as a 1 x n matrix times A, which already has u on the left). */
int oracle_mxm(ores_t *out,
               int ctype, int64_t cnrows, int64_t cncols, int64_t cnvals, const uint64_t *CI, const uint64_t *CJ, const void *CX,
               int has_mask, int mtype, int64_t mnvals, const uint64_t *MI, const uint64_t *MJ, const void *MX,
               int has_accum, int accum_op, int accum_t,
               int add_op, int mul_op, int mul_t,
               int atype, int64_t anrows, int64_t ancols, int64_t anvals, const uint64_t *AI, const uint64_t *AJ, const void *AX,
               int btype, int64_t bnrows, int64_t bncols, int64_t bnvals, const uint64_t *BI, const uint64_t *BJ, const void *BX,
               const odesc_t *d)
{
    csr_t A, B, C, M; memset(&A, 0, sizeof A); memset(&B, 0, sizeof B); memset(&C, 0, sizeof C); memset(&M, 0, sizeof M);
    memset(out, 0, sizeof *out);
    if (csr_from_coo(&A, atype, anrows, ancols, anvals, AI, AJ, AX, d->tran0)) return -1;
    if (csr_from_coo(&B, btype, bnrows, bncols, bnvals, BI, BJ, BX, d->tran1)) return -1;
    if (csr_from_coo(&C, ctype, cnrows, cncols, cnvals, CI, CJ, CX, false)) return -1;
    if (has_mask && csr_from_coo(&M, mtype, cnrows, cncols, mnvals, MI, MJ, MX, false)) return -1;
    if (A.ncols != B.nrows || A.nrows != C.nrows || B.ncols != C.ncols) return -2;

    int ztype = 0; { val_t z; memset(&z, 0, sizeof z); (void)binop(mul_op, mul_t, z, z, &ztype); }
    const int64_t n = C.ncols;
    /* dense sparse-accumulator for one row of T */
    val_t *spa = malloc((size_t)n * sizeof(val_t));
    uint8_t *mark = calloc((size_t)n, 1);
    int64_t *list = malloc((size_t)n * sizeof(int64_t));
    /* worst case per row: |C(i,:)| + |T(i,:)| <= 2n, grow the output dynamically */
    int64_t cap = cnvals + 1024, cnt = 0;
    uint64_t *OI = malloc((size_t)cap * 8), *OJ = malloc((size_t)cap * 8);
    val_t *OV = malloc((size_t)cap * sizeof(val_t));
    if (!spa || !mark || !list || !OI || !OJ || !OV) return -1;

    for (int64_t i = 0; i < C.nrows; ++i) {
        /* step 2: T(i,:) */
        int64_t nl = 0;
        for (int64_t pa = A.ptr[i]; pa < A.ptr[i + 1]; ++pa) {
            const int64_t k = A.col[pa];
            const val_t a = vcast(A.val[pa], A.type, mul_t);
            for (int64_t pb = B.ptr[k]; pb < B.ptr[k + 1]; ++pb) {
                const int64_t j = B.col[pb];
                const val_t b = vcast(B.val[pb], B.type, mul_t);
                int zt; const val_t prod = binop(mul_op, mul_t, a, b, &zt);
                if (!mark[j]) { mark[j] = 1; spa[j] = prod; list[nl++] = j; }
                else { int at; spa[j] = binop(add_op, ztype, spa[j], prod, &at); }
            }
        }
        /* ascending column order: merge the three sorted streams C(i,:), T(i,:), M(i,:) by scanning
         * the union of their columns */
        /* sort list (insertion for short, qsort otherwise) */
        if (nl > 1) {
            if (nl < 32) { for (int64_t x = 1; x < nl; ++x) { int64_t v = list[x], y = x - 1; while (y >= 0 && list[y] > v) { list[y + 1] = list[y]; --y; } list[y + 1] = v; } }
            else {
                qsort(list, (size_t)nl, sizeof(int64_t), cmp_i64);
            }
        }
        int64_t pc = C.ptr[i], pce = C.ptr[i + 1], pt = 0, pm = has_mask ? M.ptr[i] : 0, pme = has_mask ? M.ptr[i + 1] : 0;
        while (pc < pce || pt < nl) {
            const int64_t jc = pc < pce ? C.col[pc] : INT64_MAX, jt = pt < nl ? list[pt] : INT64_MAX;
            const int64_t j = jc < jt ? jc : jt;
            const bool cp = jc == j, tp = jt == j;
            /* step 4: mask */
            bool m = !(d->mask_comp && !has_mask);   /* C<!NULL>: the complement of "no mask" lets nothing through (C API 1.3, 4.3) */
            if (has_mask) {
                while (pm < pme && M.col[pm] < j) ++pm;
                m = pm < pme && M.col[pm] == j;
                if (m && !d->mask_struct) m = vcast(M.val[pm], M.type, T_BOOL).b;
                if (d->mask_comp) m = !m;
            }
            val_t res; bool rp = false; memset(&res, 0, sizeof res);
            if (m) {   /* steps 3 and 5 */
                if (has_accum) {
                    if (cp && tp) {
                        int zt; val_t z = binop(accum_op, accum_t, vcast(C.val[pc], C.type, accum_t), vcast(spa[j], ztype, accum_t), &zt);
                        res = vcast(z, zt, C.type); rp = true;
                    } else if (cp) { res = C.val[pc]; rp = true; }
                    else { res = vcast(spa[j], ztype, C.type); rp = true; }
                } else if (tp) { res = vcast(spa[j], ztype, C.type); rp = true; }
            } else if (!d->replace && cp) { res = C.val[pc]; rp = true; }
            if (rp) {
                if (cnt == cap) { cap *= 2; OI = realloc(OI, (size_t)cap * 8); OJ = realloc(OJ, (size_t)cap * 8); OV = realloc(OV, (size_t)cap * sizeof(val_t)); if (!OI || !OJ || !OV) return -1; }
                OI[cnt] = (uint64_t)i; OJ[cnt] = (uint64_t)j; OV[cnt] = res; ++cnt;
            }
            if (cp) ++pc;
            if (tp) ++pt;
        }
        for (int64_t x = 0; x < nl; ++x) mark[list[x]] = 0;
    }
    out->nvals = cnt; out->I = OI; out->J = OJ;
    out->X = malloc((size_t)(cnt ? cnt : 1) * tsize(ctype));
    for (int64_t k = 0; k < cnt; ++k) vstore(ctype, out->X, (size_t)k, OV[k]);
    free(OV); free(spa); free(mark); free(list);
    csr_free(&A); csr_free(&B); csr_free(&C); if (has_mask) csr_free(&M);
    return 0;
}
