// matrix_assign.cu -- sub-matrix assignment and extraction on the device:
//
//   GrB_Matrix_assign_<T>   C<Mask>(I,J) = accum(C(I,J), x)     /root/reference/pygraphblas/matrix.py:3106-3238 (assign_scalar;
//                                                               Matrix.dense :183 -> :179, used by tests/test_matrix.py:858-864 test_pow)
//   GrB_Matrix_extract      C<Mask> = accum(C, A(I,J))          matrix.py:2807-2860 (extract_matrix, M[0:1, :] slices)
//   GrB_Matrix_assign       C<Mask>(I,J) = accum(C(I,J), A)     matrix.py:3057-3104 (assign_matrix)
//   GrB_Col_extract / GrB_Row_assign / GrB_Col_assign           matrix.py:2862-2897, 3005-3031
//   GxB_Matrix_diag / GxB_Vector_diag                           matrix.py:2202-2236 / vector.py diag
//   GrB_Matrix_kronecker_BinaryOp                               matrix.py:2739-2805
//
// Every operation forms T (a CSR in HBM) with one or two streaming kernels and hands it to the common write-back
// C<Mask> = accum(C, T) of spgemm.cu.  GrB_assign semantics (C API 1.3 section 4.3.7): the result Z is C with the
// region I x J replaced by T (no accumulator: entries of C inside the region that T lacks are deleted) or merged with
// it (accumulator), and only then the mask -- which spans all of C -- and GrB_REPLACE apply.
// Index lists (GrB_ALL, explicit, GxB_RANGE / STRIDE / BACKWARDS as /root/reference/pygraphblas/base.py:216-252 builds
// them) are expanded on the host: they are call arguments, not data.  Nothing here computes values on the host.
#include "common.cuh"
#include <algorithm>
#include <vector>
#include <cub/device/device_segmented_radix_sort.cuh>
#include "../../include/b200grb_compat.h"

static inline int agrid(int64_t n, int threads = 256) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, threads), (int64_t)G.num_sms * 16)); }

GrB_Info index_list(const GrB_Index *I, GrB_Index ni, uint64_t dim, bool *all, std::vector<uint64_t> &out, std::string *err, const char *fn);   // vector_ops.cu
extern "C" GrB_Info GrB_Matrix_dup(GrB_Matrix *C, const GrB_Matrix A);
extern "C" GrB_Info GrB_Matrix_free(GrB_Matrix *A);
extern "C" GrB_Info GrB_Matrix_new(GrB_Matrix *A, GrB_Type type, GrB_Index nrows, GrB_Index ncols);

static GrB_BinaryOp second_op(int tc) {
    switch (tc) {
        case TC_BOOL: return GrB_SECOND_BOOL;     case TC_INT8: return GrB_SECOND_INT8;     case TC_INT16: return GrB_SECOND_INT16;
        case TC_INT32: return GrB_SECOND_INT32;   case TC_INT64: return GrB_SECOND_INT64;   case TC_UINT8: return GrB_SECOND_UINT8;
        case TC_UINT16: return GrB_SECOND_UINT16; case TC_UINT32: return GrB_SECOND_UINT32; case TC_UINT64: return GrB_SECOND_UINT64;
        case TC_FP32: return GrB_SECOND_FP32;     default: return GrB_SECOND_FP64;
    }
}

template <typename T> static GrB_Info upload(const std::vector<T> &h, T **d, std::string *err) {
    GB_TRY(dalloc(d, h.size(), err));
    if (!h.empty()) CU_TRY(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, G.stream), err);
    CU_TRY(cudaStreamSynchronize(G.stream), err);          // h is a caller-owned temporary
    return GrB_SUCCESS;
}
static GrB_Info read_i64(const int64_t *d, int64_t *h, std::string *err) {
    CU_TRY(cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, G.stream), err);
    CU_TRY(cudaStreamSynchronize(G.stream), err);
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ the region I x J of a matrix
struct Region {
    bool all_rows = true, all_cols = true;
    std::vector<uint64_t> I, J;          // as given (after expansion), empty when all
    uint8_t *rowflag = nullptr;          // [nrows] 1 where the row is in I   (NULL: every row)
    uint8_t *colflag = nullptr;          // [ncols] 1 where the column is in J (NULL: every column)
    uint32_t *jsorted = nullptr;         // sorted distinct columns of J        (NULL: 0..ncols-1)
    int64_t nj_distinct = 0, ni_distinct = 0;
    void release() { dfree(rowflag); dfree(colflag); dfree(jsorted); rowflag = colflag = nullptr; jsorted = nullptr; }
};
__global__ void flag_kernel(const uint64_t *idx, int64_t k, uint8_t *flag) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < k; q += (int64_t)gridDim.x * blockDim.x) flag[idx[q]] = 1;
}
static GrB_Info region_build(Region &R, const GrB_Index *I, GrB_Index ni, const GrB_Index *J, GrB_Index nj, uint64_t nrows, uint64_t ncols,
                             std::string *err, const char *fn) {
    GB_TRY(index_list(I, ni, nrows, &R.all_rows, R.I, err, fn));
    GB_TRY(index_list(J, nj, ncols, &R.all_cols, R.J, err, fn));
    R.ni_distinct = (int64_t)nrows; R.nj_distinct = (int64_t)ncols;
    if (!R.all_rows) {
        std::vector<uint64_t> s(R.I); std::sort(s.begin(), s.end()); s.erase(std::unique(s.begin(), s.end()), s.end());
        R.ni_distinct = (int64_t)s.size();
        uint64_t *d = nullptr; GB_TRY(upload(s, &d, err));
        GB_TRY(dalloc(&R.rowflag, (size_t)nrows, err));
        CU_TRY(cudaMemsetAsync(R.rowflag, 0, (size_t)nrows, G.stream), err);
        if (!s.empty()) { flag_kernel<<<agrid((int64_t)s.size()), 256, 0, G.stream>>>(d, (int64_t)s.size(), R.rowflag); GB_LAUNCHED(); }
        dfree(d);
    }
    if (!R.all_cols) {
        std::vector<uint64_t> s(R.J); std::sort(s.begin(), s.end()); s.erase(std::unique(s.begin(), s.end()), s.end());
        R.nj_distinct = (int64_t)s.size();
        std::vector<uint32_t> s32(s.begin(), s.end());
        GB_TRY(upload(s32, &R.jsorted, err));
        uint64_t *d = nullptr; GB_TRY(upload(s, &d, err));
        GB_TRY(dalloc(&R.colflag, (size_t)ncols, err));
        CU_TRY(cudaMemsetAsync(R.colflag, 0, (size_t)ncols, G.stream), err);
        if (!s.empty()) { flag_kernel<<<agrid((int64_t)s.size()), 256, 0, G.stream>>>(d, (int64_t)s.size(), R.colflag); GB_LAUNCHED(); }
        dfree(d);
    }
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ T = the region filled with one value
__global__ void fill_count_kernel(const uint8_t *rowflag, int64_t nrows, int64_t per_row, int64_t *rowptr) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nrows; r += (int64_t)gridDim.x * blockDim.x)
        rowptr[r] = (r < nrows && (!rowflag || rowflag[r])) ? per_row : 0;
}
__global__ void fill_entries_kernel(const int64_t *rowptr, int64_t nrows, const uint32_t *jsorted, uint32_t *col, void *val, int tc, Sc x) {
    // one warp per row
    const int lane = threadIdx.x & 31;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        for (int64_t k = b + lane; k < e; k += 32) { col[k] = jsorted ? jsorted[k - b] : (uint32_t)(k - b); sc_store(tc, val, (size_t)k, x); }
    }
}
static GrB_Info region_filled(const Region &R, int64_t nrows, int64_t ncols, int tc, Sc x, Csr &T, std::string *err) {
    T = Csr(); T.nrows = nrows; T.ncols = ncols;
    GB_TRY(dalloc(&T.rowptr, (size_t)nrows + 1, err));
    fill_count_kernel<<<agrid(nrows + 1), 256, 0, G.stream>>>(R.rowflag, nrows, R.nj_distinct, T.rowptr); GB_LAUNCHED();
    GB_TRY(dev_exclusive_scan(T.rowptr, nrows + 1, err));
    T.nnz = R.ni_distinct * R.nj_distinct;
    GB_TRY(dalloc(&T.col, (size_t)T.nnz, err));
    GB_TRY(dmalloc(&T.val, (size_t)T.nnz * tc_size(tc) + 16, err));
    if (T.nnz > 0) { fill_entries_kernel<<<agrid(nrows * 32), 256, 0, G.stream>>>(T.rowptr, nrows, R.jsorted, T.col, T.val, tc, x); GB_LAUNCHED(); }
    GB_TRY(dev_build_rowptr32(T, err));
    CU_TRY(cudaGetLastError(), err);
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ C minus the region (entries of C outside I x J)
__global__ void outside_count_kernel(const int64_t *ptr, const uint32_t *col, int64_t nrows, const uint8_t *rowflag, const uint8_t *colflag, int64_t *cnt) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nrows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = 0;
        if (r < nrows) {
            const bool rin = !rowflag || rowflag[r];
            for (int64_t k = ptr[r]; k < ptr[r + 1]; ++k) c += !(rin && (!colflag || colflag[col[k]]));
        }
        cnt[r] = c;
    }
}
__global__ void outside_fill_kernel(const int64_t *ptr, const uint32_t *col, const uint8_t *val, int vsize, int64_t nrows, const uint8_t *rowflag,
                                    const uint8_t *colflag, const int64_t *optr, uint32_t *ocol, uint8_t *oval) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        const bool rin = !rowflag || rowflag[r];
        int64_t o = optr[r];
        for (int64_t k = ptr[r]; k < ptr[r + 1]; ++k) if (!(rin && (!colflag || colflag[col[k]]))) {
            ocol[o] = col[k];
            for (int b = 0; b < vsize; ++b) oval[o * vsize + b] = val[k * vsize + b];
            ++o;
        }
    }
}
static GrB_Info csr_outside_region(const Csr &c, size_t vsize, const Region &R, Csr &out, std::string *err) {
    out = Csr(); out.nrows = c.nrows; out.ncols = c.ncols;
    GB_TRY(dalloc(&out.rowptr, (size_t)c.nrows + 1, err));
    outside_count_kernel<<<agrid(c.nrows + 1), 256, 0, G.stream>>>(c.rowptr, c.col, c.nrows, R.rowflag, R.colflag, out.rowptr); GB_LAUNCHED();
    GB_TRY(dev_exclusive_scan(out.rowptr, c.nrows + 1, err));
    GB_TRY(read_i64(out.rowptr + c.nrows, &out.nnz, err));
    GB_TRY(dalloc(&out.col, (size_t)out.nnz, err));
    GB_TRY(dmalloc(&out.val, (size_t)out.nnz * vsize + 16, err));
    if (out.nnz > 0) {
        outside_fill_kernel<<<agrid(c.nrows), 256, 0, G.stream>>>(c.rowptr, c.col, (const uint8_t *)c.val, (int)vsize, c.nrows, R.rowflag, R.colflag,
                                                                  out.rowptr, out.col, (uint8_t *)out.val); GB_LAUNCHED();
    }
    GB_TRY(dev_build_rowptr32(out, err));
    CU_TRY(cudaGetLastError(), err);
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ GrB_assign write-back:  C<Mask>(region) = accum(C(region), T)
// T is a CSR of C's dimensions holding the new content of the region (type ttc); consumed.
// t_covers_region: T has an entry at every position of the region (scalar fill), so nothing needs deleting first.
static GrB_Info assign_writeback(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const DescFlags &f, const Region &R,
                                 Csr &T, int ttc, bool t_covers_region, std::string *err) {
    const int ctc = C->type->code;
    if (!Mask && f.mask_comp) return matrix_writeback(C, nullptr, accum, f, T, ttc, false, err);     // nothing is let through
    DescFlags plain{}; plain.replace = false; plain.mask_comp = false; plain.mask_struct = false; plain.tran0 = plain.tran1 = false; plain.axb = f.axb;
    const GrB_BinaryOp merge = accum ? accum : second_op(ctc);
    const bool whole = R.all_rows && R.all_cols;
    // Z = C with the region replaced by / merged with T, built in a scratch matrix unless it can go straight into C
    GrB_Matrix Zm = C;
    if (Mask) { GrB_Info r = GrB_Matrix_dup(&Zm, C); if (r != GrB_SUCCESS) { csr_free(T); return r; } }
    GrB_Info r = GrB_SUCCESS;
    if (!accum && !t_covers_region && !whole) {
        // entries of C inside the region that T lacks are deleted: drop the region from C first
        r = matrix_ensure_device(Zm);
        if (r == GrB_SUCCESS && Zm->dev.nnz > 0) {
            Csr keep;
            r = csr_outside_region(Zm->dev, Zm->type->size, R, keep, err);
            if (r == GrB_SUCCESS) matrix_adopt_device(Zm, keep); else csr_free(keep);
        }
    }
    if (r == GrB_SUCCESS) {
        if (!accum && whole) r = matrix_writeback(Zm, nullptr, nullptr, plain, T, ttc, false, err);        // Z = T
        else r = matrix_writeback(Zm, nullptr, merge, plain, T, ttc, false, err);                          // Z = C (+) T
    } else csr_free(T);
    if (!Mask || r != GrB_SUCCESS) { if (Zm != C) GrB_Matrix_free(&Zm); return r; }
    // C<Mask> = Z  (mask and GrB_REPLACE span all of C)
    r = matrix_ensure_device(Zm);
    if (r == GrB_SUCCESS) {
        Csr z = Zm->dev; Zm->dev = Csr(); Zm->host_valid = true;      // steal Z's CSR (values already of C's type)
        r = matrix_writeback(C, Mask, nullptr, f, z, ctc, false, err);
    }
    GrB_Matrix_free(&Zm);
    return r;
}

static GrB_Info matrix_assign_scalar(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, Sc x, int xtc, const GrB_Index *I, GrB_Index ni,
                                     const GrB_Index *J, GrB_Index nj, const GrB_Descriptor desc, const char *fn) {
    GB_LOCK; GB_CHECK_INIT;
    if (!C) return gb_fail(GrB_NULL_POINTER, nullptr, "%s: NULL argument", fn);
    if (!gb_valid_matrix(C) || (Mask && !gb_valid_matrix(Mask))) return gb_fail(GrB_UNINITIALIZED_OBJECT, nullptr, "%s: invalid object", fn);
    std::string *err = &C->err;
    if (accum && accum->opcode == OP_USER) return gb_fail(GrB_INVALID_VALUE, err, "%s: user-defined accumulators are host function pointers and cannot run on the GPU", fn);
    if (Mask && (Mask->nrows != C->nrows || Mask->ncols != C->ncols)) return gb_fail(GrB_DIMENSION_MISMATCH, err, "%s: the mask must have C's dimensions", fn);
    if (!G.have_device) return gb_fail(GrB_PANIC, err, "%s: no CUDA device: libb200grb computes only on the GPU (no CPU fallback)", fn);
    if (C->nrows >= ((uint64_t)1 << 31) || C->ncols >= ((uint64_t)1 << 31)) return gb_fail(GrB_INVALID_VALUE, err, "%s: dimensions beyond 2^31 cannot be filled", fn);
    GbBurble burble(fn);
    const DescFlags f = desc_flags(desc);
    Region R;
    GrB_Info r = region_build(R, I, ni, J, nj, C->nrows, C->ncols, err, fn);
    if (r != GrB_SUCCESS) { R.release(); return r; }
    if ((double)R.ni_distinct * (double)R.nj_distinct >= 4.0e9) { R.release(); return gb_fail(GrB_OUT_OF_MEMORY, err, "%s: the filled region would hold >= 4e9 entries", fn); }
    Csr T;
    r = region_filled(R, (int64_t)C->nrows, (int64_t)C->ncols, xtc, x, T, err);
    if (r == GrB_SUCCESS) { burble.note("region fill + write-back", (double)T.nnz * (4 + tc_size(xtc))); r = assign_writeback(C, Mask, accum, f, R, T, xtc, /*t_covers_region=*/true, err); }
    else csr_free(T);
    R.release();
    return r;
}

#define GB_MASSIGN(TN, CT, TC, FIELD) \
    extern "C" GrB_Info GrB_Matrix_assign_##TN(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, CT x, const GrB_Index *I, GrB_Index ni, \
                                               const GrB_Index *J, GrB_Index nj, const GrB_Descriptor desc) { \
        Sc s; s.u = 0; s.FIELD = x; return matrix_assign_scalar(C, Mask, accum, s, TC, I, ni, J, nj, desc, "GrB_Matrix_assign_" #TN); }
GB_MASSIGN(BOOL, bool, TC_BOOL, u) GB_MASSIGN(INT8, int8_t, TC_INT8, i) GB_MASSIGN(INT16, int16_t, TC_INT16, i) GB_MASSIGN(INT32, int32_t, TC_INT32, i)
GB_MASSIGN(INT64, int64_t, TC_INT64, i) GB_MASSIGN(UINT8, uint8_t, TC_UINT8, u) GB_MASSIGN(UINT16, uint16_t, TC_UINT16, u) GB_MASSIGN(UINT32, uint32_t, TC_UINT32, u)
GB_MASSIGN(UINT64, uint64_t, TC_UINT64, u) GB_MASSIGN(FP32, float, TC_FP32, d) GB_MASSIGN(FP64, double, TC_FP64, d)
#undef GB_MASSIGN
