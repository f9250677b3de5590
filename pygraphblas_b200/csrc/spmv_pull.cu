// Masked pull kernels for the saturating monoids (LOR, LAND, ANY): the BFS step.
#include "spmv_args.cuh"

template <typename ZT> __device__ __forceinline__ bool monoid_saturated(int add, ZT v) {
    switch (add) {
        case OP_LOR: return v != (ZT)0;
        case OP_LAND: return v == (ZT)0;
        case OP_ANY: return true;
        default: return false;
    }
}
// The three monoids of this kernel (LOR, LAND, ANY) need no running value: the fold of a row's products is
// decided by how many there are (0, 1, more), whether one of them saturates, and the first one --
//   0 products: no entry;  1: that product, as is;  more: ANY -> any of them, LOR -> "one was non-zero",
//   LAND -> "none was zero" (1 or 0 in the monoid's type).
// A row may stop early once its result can no longer change.
template <typename ZT> __device__ __forceinline__ ZT pull_result(int add, int n, bool sat, ZT first) {
    if (n <= 1 || add == OP_ANY) return first;
    return add == OP_LOR ? (ZT)(sat ? 1 : 0) : (ZT)(sat ? 0 : 1);
}
template <typename ZT> __device__ __forceinline__ bool pull_settled(int add, ZT v) {     // saturating AND already the final value
    switch (add) {
        case OP_LOR: return v == (ZT)1;
        case OP_LAND: return v == (ZT)0;
        case OP_ANY: return true;
        default: return false;
    }
}
template <typename XT, typename ZT>
__device__ __forceinline__ ZT pull_product(const PullArgs &p, const XT *aval, const XT *uval, uint32_t k, uint32_t c) {
    const XT a = gload<XT>(aval + k), u = gload<XT>(uval + c);
    return p.flip ? MulApply<XT, ZT>::f(p.mul_op, u, a) : MulApply<XT, ZT>::f(p.mul_op, a, u);
}

// Warp batches of 32 rows, their entries flattened: lane i owns row base+i (mask, accumulators, result) while
// the entries of all 32 rows are walked 32 at a time, so short rows cost one slot per entry instead of one
// warp iteration per row.  Rows advance in rounds of at most `cap` entries each; a row whose result is settled
// leaves the batch at the end of the round.
template <typename XT, typename ZT>
__global__ void __launch_bounds__(256) spmv_masked_pull_kernel(const PullArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const XT *aval = static_cast<const XT *>(p.aval), *uval = static_cast<const XT *>(p.uval);
    ZT *tval = static_cast<ZT *>(p.tval);
    const int add = p.add_op;
    for (int64_t base = warp * 32; base < p.nrows; base += nwarps * 32) {
        const int64_t mr = base + lane;
        bool m = false;
        if (mr < p.nrows) {
            m = p.mpres ? p.mpres[mr] != 0 : true;
            if (m && !p.mask_struct) m = sc_cast(sc_load(p.mtc, p.mval, (size_t)mr), p.mtc, TC_BOOL).u != 0;
            if (p.mask_comp) m = !m;
            if (!m) p.tpres[mr] = 0;
        }
        uint32_t pos = 0, rem = 0;
        if (m) {
            pos = p.rowptr[mr]; rem = p.rowptr[mr + 1] - pos;
            if (rem > PULL_LONG) { p.long_rows[atomicAdd(p.long_count, 1)] = (uint32_t)mr; rem = 0; m = false; }
        }
        int n_it = 0; bool sat = false; ZT first = (ZT)0;
        unsigned active;
        while ((active = __ballot_sync(0xffffffffu, rem > 0)) != 0) {
            const uint32_t cap = __popc(active) > 8 ? 32u : 128u;
            const uint32_t take = rem < cap ? rem : cap;
            uint32_t incl = take;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const uint32_t off = incl - take, total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t delta = pos - off;                  // entry index = delta(owner) + flat position
            for (uint32_t f0 = 0; f0 < total; f0 += 32) {
                const uint32_t f = f0 + lane;
                int own = 0;                                   // first lane whose inclusive end exceeds f
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) { const uint32_t t = __shfl_sync(0xffffffffu, incl, own + step - 1); if (t <= f) own += step; }
                const uint32_t k = __shfl_sync(0xffffffffu, delta, own) + f;
                bool has = false; ZT v = (ZT)0;
                if (f < total) {
                    const uint32_t c = __ldg(p.col + k);
                    if (!p.upres || __ldg(p.upres + c)) { v = pull_product<XT, ZT>(p, aval, uval, k, c); has = true; }
                }
                const unsigned hasmask = __ballot_sync(0xffffffffu, has);
                const unsigned satmask = __ballot_sync(0xffffffffu, has && monoid_saturated<ZT>(add, v));
                // the slots of this chunk that belong to this lane's row
                const uint32_t s0 = off > f0 ? off : f0, s1 = (off + take) < (f0 + 32) ? (off + take) : (f0 + 32);
                unsigned seg = 0;
                if (s1 > s0) { const uint32_t len = s1 - s0; seg = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << (s0 - f0); }
                const unsigned mine = hasmask & seg;
                const ZT fv = shfl_idx_t<ZT>(v, mine ? __ffs(mine) - 1 : lane);
                if (mine) {
                    if (n_it == 0) first = fv;
                    n_it = min(2, n_it + __popc(mine));
                    sat |= (satmask & seg) != 0;
                }
            }
            if (sat && (n_it >= 2 || pull_settled<ZT>(add, first))) rem = 0;
            else { rem -= take; pos += take; }
        }
        if (m) { tval[mr] = pull_result<ZT>(add, n_it, sat, first); p.tpres[mr] = (uint8_t)(n_it > 0); }
    }
}
// Long rows (hubs): one 1024-thread CTA per row, 1024 entries per iteration, early exit CTA-wide.
template <typename XT, typename ZT>
__global__ void __launch_bounds__(1024) spmv_pull_long_kernel(const PullArgs p) {
    __shared__ int s_n[32]; __shared__ int s_sat[32]; __shared__ ZT s_first[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const XT *aval = static_cast<const XT *>(p.aval), *uval = static_cast<const XT *>(p.uval);
    ZT *tval = static_cast<ZT *>(p.tval);
    const int add = p.add_op;
    const int nlong = *p.long_count;
    for (int w = blockIdx.x; w < nlong; w += gridDim.x) {
        const uint32_t r = p.long_rows[w];
        const uint32_t rs = p.rowptr[r], re = p.rowptr[r + 1];
        int n_it = 0; bool sat = false; ZT first = (ZT)0;
        for (uint32_t b0 = rs; b0 < re; b0 += 1024) {
            const uint32_t k = b0 + threadIdx.x;
            int stop = 0;
            if (k < re) {
                const uint32_t c = __ldg(p.col + k);
                if (!p.upres || __ldg(p.upres + c)) {
                    const ZT v = pull_product<XT, ZT>(p, aval, uval, k, c);
                    if (n_it == 0) first = v;
                    n_it = min(2, n_it + 1);
                    if (monoid_saturated<ZT>(add, v)) { sat = true; stop = pull_settled<ZT>(add, v); }
                }
            }
            if (__syncthreads_or(stop)) break;
        }
        const unsigned hasmask = __ballot_sync(0xffffffffu, n_it > 0);
        const int wn = min(2, __reduce_add_sync(0xffffffffu, n_it));
        const int wsat = __any_sync(0xffffffffu, sat);
        const ZT wfirst = shfl_idx_t<ZT>(first, hasmask ? __ffs(hasmask) - 1 : 0);
        if (lane == 0) { s_n[wid] = wn; s_sat[wid] = wsat; s_first[wid] = wfirst; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int N = 0, S = 0; ZT F = (ZT)0;
            for (int q = 0; q < 32; ++q) { if (s_n[q] && !N) F = s_first[q]; N = min(2, N + s_n[q]); S |= s_sat[q]; }
            tval[r] = pull_result<ZT>(add, N, S != 0, F); p.tpres[r] = (uint8_t)(N > 0);
        }
        __syncthreads();
    }
}
GrB_Info spmv_masked_pull_dispatch(int xt, int zt, const PullArgs &a, std::string *err) {
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(a.nrows, 256), (int64_t)G.num_sms * 8));
    const int lgrid = (int)std::max<int64_t>(1, std::min<int64_t>(a.long_cap, (int64_t)G.num_sms * 2));
#define GB_PULL(XT_, ZT_) do { spmv_masked_pull_kernel<XT_, ZT_><<<grid, 256, 0, G.stream>>>(a); GB_LAUNCHED(); \
        if (a.has_long) { spmv_pull_long_kernel<XT_, ZT_><<<lgrid, 1024, 0, G.stream>>>(a); GB_LAUNCHED(); } return GrB_SUCCESS; } while (0)
    if (xt == zt) {
        switch (xt) {
#define GB_GEN(TC, T) case TC: GB_PULL(T, T);
            GB_GEN(TC_BOOL, bool) GB_GEN(TC_INT8, int8_t) GB_GEN(TC_INT16, int16_t) GB_GEN(TC_INT32, int32_t) GB_GEN(TC_INT64, int64_t)
            GB_GEN(TC_UINT8, uint8_t) GB_GEN(TC_UINT16, uint16_t) GB_GEN(TC_UINT32, uint32_t) GB_GEN(TC_UINT64, uint64_t)
            GB_GEN(TC_FP32, float) GB_GEN(TC_FP64, double)
#undef GB_GEN
        }
    } else if (zt == TC_BOOL) {
        switch (xt) {
#define GB_GEN(TC, T) case TC: GB_PULL(T, bool);
            GB_GEN(TC_INT8, int8_t) GB_GEN(TC_INT16, int16_t) GB_GEN(TC_INT32, int32_t) GB_GEN(TC_INT64, int64_t)
            GB_GEN(TC_UINT8, uint8_t) GB_GEN(TC_UINT16, uint16_t) GB_GEN(TC_UINT32, uint32_t) GB_GEN(TC_UINT64, uint64_t)
            GB_GEN(TC_FP32, float) GB_GEN(TC_FP64, double)
#undef GB_GEN
        }
    }
#undef GB_PULL
    return gb_fail(GrB_DOMAIN_MISMATCH, err, "mxv: unsupported semiring domains (x=%d, z=%d)", xt, zt);
}

