// Run plan, hot-column plan, and the FP32 / FP64 instantiations of the run kernels.
#include "spmv_run.cuh"
#include <cub/device/device_radix_sort.cuh>

// u_perm[i] = u[perm[i]]  (element size 1/2/4/8)
__global__ void permute_u_kernel(const uint32_t *perm, const uint8_t *u, uint8_t *out, int vsize, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t c = perm[i];
        switch (vsize) {
            case 1: out[i] = u[c]; break;
            case 2: ((uint16_t *)out)[i] = ((const uint16_t *)u)[c]; break;
            case 4: ((uint32_t *)out)[i] = ((const uint32_t *)u)[c]; break;
            default: ((uint64_t *)out)[i] = ((const uint64_t *)u)[c]; break;
        }
    }
}

void spmv_permute_u(const uint32_t *perm, const void *u, void *out, int vsize, int64_t n) {
    if (n > 0) { permute_u_kernel<<<hgrid(n), 256, 0, G.stream>>>(perm, (const uint8_t *)u, (uint8_t *)out, vsize, n); GB_LAUNCHED(); }
}

// ---- run plan (cached per CSR)
__global__ void plan_nonempty_kernel(const uint32_t *rowptr, int64_t nrows, int64_t *flag, uint8_t *pres) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        const int ne = rowptr[r + 1] > rowptr[r];
        flag[r] = ne; pres[r] = (uint8_t)ne;
    }
}
__global__ void plan_rows_kernel(const uint32_t *rowptr, const int64_t *rank, int64_t nrows, uint32_t *nzrow, uint32_t *headw) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t rs = rowptr[r];
        if (rowptr[r + 1] > rs) { nzrow[rank[r]] = (uint32_t)r; atomicOr(&headw[rs >> 5], 1u << (rs & 31)); }
    }
}
__global__ void plan_runs_kernel(const uint32_t *headw, int64_t nruns, int64_t nwords, uint16_t *lane_rank, int64_t *run_cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t run = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (run >= nruns) return;
    const int64_t w = run * 8 + (lane >> 2);
    const uint32_t hw = w < nwords ? headw[w] : 0u;
    const int pc = __popc((hw >> ((lane & 3) * 8)) & 0xffu);
    int inc = pc;
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    lane_rank[run * 32 + lane] = (uint16_t)(inc - pc);
    if (lane == 31) run_cnt[run] = inc;
}
__global__ void plan_base_kernel(const int64_t *scan, int64_t nruns, uint32_t *run_base) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= nruns; k += (int64_t)gridDim.x * blockDim.x) run_base[k] = (uint32_t)scan[k];
}
// the row that starts last inside a run always holds the run's last entry: it is the run's "open" row
// (possibly ending exactly at the run's end), completed by the fix-up kernel
__global__ void plan_tails_kernel(const uint32_t *run_base, const uint32_t *nzrow, const uint32_t *rowptr, int64_t nruns,
                                  int32_t *tail_row, uint32_t *tail_last) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nruns; k += (int64_t)gridDim.x * blockDim.x) {
        int32_t tr = -1; uint32_t tl = 0;
        if (run_base[k + 1] > run_base[k]) {
            const uint32_t r = nzrow[run_base[k + 1] - 1];
            const uint32_t re = rowptr[r + 1];
            tr = (int32_t)r; tl = (re - 1) / RUN;
        }
        tail_row[k] = tr; tail_last[k] = tl;
    }
}
static inline int rgrid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (int64_t)G.num_sms * 16)); }

GrB_Info spmv_run_plan(Csr &c, std::string *err) {
    if (c.run_headw) return GrB_SUCCESS;
    if (!c.rowptr32) return gb_fail(GrB_INVALID_VALUE, err, "mxv: matrices with >= 2^32 entries are not supported");
    const int64_t nwords = ceil_div(c.nnz, 32);
    c.nruns = ceil_div(c.nnz, RUN);
    int64_t *flag = nullptr, *cnt = nullptr;
    GB_TRY(dalloc(&flag, (size_t)c.nrows + 1, err));
    GB_TRY(dalloc(&cnt, (size_t)c.nruns + 1, err));
    GB_TRY(dalloc(&c.pres_tmpl, (size_t)c.nrows, err));
    GB_TRY(dalloc(&c.run_headw, (size_t)nwords + 8, err));
    GB_TRY(dalloc(&c.run_lane, (size_t)c.nruns * 32, err));
    GB_TRY(dalloc(&c.run_base, (size_t)c.nruns + 1, err));
    GB_TRY(dalloc(&c.run_tail_row, (size_t)c.nruns, err));
    GB_TRY(dalloc(&c.run_tail_last, (size_t)c.nruns, err));
    CU_TRY(cudaMemsetAsync(c.run_headw, 0, ((size_t)nwords + 8) * 4, G.stream), err);
    CU_TRY(cudaMemsetAsync(flag + c.nrows, 0, 8, G.stream), err);
    plan_nonempty_kernel<<<rgrid(c.nrows), 256, 0, G.stream>>>(c.rowptr32, c.nrows, flag, c.pres_tmpl); GB_LAUNCHED();
    GB_TRY(dev_exclusive_scan(flag, c.nrows + 1, err));
    int64_t nz = 0;
    CU_TRY(cudaMemcpyAsync(&nz, flag + c.nrows, 8, cudaMemcpyDeviceToHost, G.stream), err);
    CU_TRY(cudaStreamSynchronize(G.stream), err);
    c.nnzrows = nz;
    GB_TRY(dalloc(&c.nzrow, (size_t)nz, err));
    plan_rows_kernel<<<rgrid(c.nrows), 256, 0, G.stream>>>(c.rowptr32, flag, c.nrows, c.nzrow, c.run_headw); GB_LAUNCHED();
    CU_TRY(cudaMemsetAsync(cnt + c.nruns, 0, 8, G.stream), err);
    plan_runs_kernel<<<(unsigned)ceil_div(c.nruns * 32, 256), 256, 0, G.stream>>>(c.run_headw, c.nruns, nwords, c.run_lane, cnt); GB_LAUNCHED();
    GB_TRY(dev_exclusive_scan(cnt, c.nruns + 1, err));
    plan_base_kernel<<<rgrid(c.nruns + 1), 256, 0, G.stream>>>(cnt, c.nruns, c.run_base); GB_LAUNCHED();
    plan_tails_kernel<<<rgrid(c.nruns), 256, 0, G.stream>>>(c.run_base, c.nzrow, c.rowptr32, c.nruns, c.run_tail_row, c.run_tail_last); GB_LAUNCHED();
    dfree(flag); dfree(cnt);
    CU_TRY(cudaGetLastError(), err);
    return GrB_SUCCESS;
}

// ---- hot-column plan: relabel the columns by descending in-degree (cached per CSR)
__global__ void hot_count_kernel(const uint32_t *col, int64_t nnz, uint32_t *deg) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) atomicAdd(&deg[col[k]], 1u);
}
__global__ void hot_iota_kernel(uint32_t *a, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) a[k] = (uint32_t)k;
}
__global__ void hot_invert_kernel(const uint32_t *perm, const uint32_t *deg_sorted, int64_t n, uint32_t *inv, unsigned long long *used) {
    unsigned long long c = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        inv[perm[k]] = (uint32_t)k; c += deg_sorted[k] != 0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(used, c);
}
__global__ void hot_cover_kernel(const uint32_t *deg_sorted, int64_t k, unsigned long long *sum) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) c += deg_sorted[i];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(sum, c);
}
__global__ void hot_relabel_kernel(const uint32_t *col, const uint32_t *inv, int64_t nnz, uint32_t *out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) out[k] = inv[col[k]];
}

GrB_Info spmv_hot_plan(Csr &c, std::string *err) {
    if (c.hcol) return GrB_SUCCESS;
    const int64_t n = c.ncols;
    uint32_t *deg = nullptr, *deg_sorted = nullptr, *ids = nullptr, *inv = nullptr; unsigned long long *used = nullptr;
    GB_TRY(dalloc(&deg, (size_t)n, err)); GB_TRY(dalloc(&deg_sorted, (size_t)n, err)); GB_TRY(dalloc(&ids, (size_t)n, err));
    GB_TRY(dalloc(&inv, (size_t)n, err)); GB_TRY(dalloc(&used, 2, err));
    GB_TRY(dalloc(&c.hperm, (size_t)n, err));
    GB_TRY(dalloc(&c.hcol, (size_t)c.nnz, err));
    CU_TRY(cudaMemsetAsync(deg, 0, (size_t)n * 4, G.stream), err);
    CU_TRY(cudaMemsetAsync(used, 0, 16, G.stream), err);
    hot_count_kernel<<<hgrid(c.nnz), 256, 0, G.stream>>>(c.col, c.nnz, deg); GB_LAUNCHED();
    hot_iota_kernel<<<hgrid(n), 256, 0, G.stream>>>(ids, n); GB_LAUNCHED();
    size_t tmp_bytes = 0;     // stable sort: equal degrees keep ascending column order (deterministic plan)
    CU_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, deg, deg_sorted, ids, c.hperm, n, 0, 32, G.stream), err);
    void *tmp = nullptr; GB_TRY(dmalloc(&tmp, tmp_bytes, err));
    CU_TRY(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, deg, deg_sorted, ids, c.hperm, n, 0, 32, G.stream), err);
    G.launches += 8;
    hot_invert_kernel<<<hgrid(n), 256, 0, G.stream>>>(c.hperm, deg_sorted, n, inv, used); GB_LAUNCHED();
    hot_relabel_kernel<<<hgrid(c.nnz), 256, 0, G.stream>>>(c.col, inv, c.nnz, c.hcol); GB_LAUNCHED();
    const int64_t topk = std::min<int64_t>(n, 40960);
    hot_cover_kernel<<<hgrid(topk), 256, 0, G.stream>>>(deg_sorted, topk, used + 1); GB_LAUNCHED();
    unsigned long long h[2] = {0, 0};
    CU_TRY(cudaMemcpyAsync(h, used, 16, cudaMemcpyDeviceToHost, G.stream), err);
    CU_TRY(cudaStreamSynchronize(G.stream), err);
    c.hused = (int64_t)h[0];
    c.hot_cover = c.nnz ? (double)h[1] / (double)c.nnz : 0.0;
    dfree(tmp); dfree(deg); dfree(deg_sorted); dfree(ids); dfree(inv); dfree(used);
    return GrB_SUCCESS;
}


template <typename T> static bool spmv_run_fast(int add, int mul, const RunArgs &a, size_t hot_bytes, int64_t hused) {
#define GB_FAST(A, M) if (add == A && mul == M) { spmv_run_launch<T, T, A, M>(a, hot_bytes, hused); return true; }
    GB_FAST(OP_PLUS, OP_TIMES) GB_FAST(OP_MIN, OP_PLUS) GB_FAST(OP_PLUS, OP_SECOND) GB_FAST(OP_PLUS, OP_FIRST)
    GB_FAST(OP_PLUS, OP_PAIR) GB_FAST(OP_MIN, OP_FIRST) GB_FAST(OP_MIN, OP_SECOND)
#undef GB_FAST
    return false;
}
bool spmv_run_fast_int(int xt, int add, int mul, const RunArgs &a, size_t hot_bytes, int64_t hused);
bool spmv_run_dispatch(int xt, int add, int mul, const RunArgs &a, size_t hot_bytes, int64_t hused) {
    switch (xt) {
        case TC_FP32: return spmv_run_fast<float>(add, mul, a, hot_bytes, hused);
        case TC_FP64: return spmv_run_fast<double>(add, mul, a, hot_bytes, hused);
        case TC_INT32:
        case TC_INT64:
        case TC_UINT32:
        case TC_UINT64:
        case TC_BOOL: return spmv_run_fast_int(xt, add, mul, a, hot_bytes, hused);
        default: return false;
    }
}
