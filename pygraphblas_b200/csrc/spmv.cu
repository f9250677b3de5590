// spmv.cu -- GrB_mxv / GrB_vxm on sm_100a:  w<mask> = accum(w, op(A) (+).(x) u)
//
// Replaces the SuiteSparse call behind /root/reference/pygraphblas/matrix.py:2716
// (Matrix.mxv) and /root/reference/pygraphblas/vector.py:961 (Vector.vxm).
//
// Layout in HBM: A is CSR (32-bit row offsets shadow, 32-bit column ids, values of
// the matrix type), vectors are dense value arrays + one presence byte per position
// (no presence array at all when every position is present).
//
// Kernel (spmv_tile_kernel): nnz-split, row-segmented.  The nnz range is cut into
// fixed tiles of SPMV_TILE entries, one CTA per tile, so R-MAT hub rows cannot
// serialise a warp.  Each thread streams its slice of colidx/values with 128-bit
// coalesced loads, gathers u[col] (L2-resident), and parks the products in shared
// memory; the CTA then reduces the row segments that fall inside the tile
// (thread-per-row for short segments, warp-per-row for long ones).  Rows that
// straddle tile boundaries leave per-tile head/tail partials that a small fix-up
// kernel combines in a fixed order -- the result is deterministic for a given matrix.
//
// Algorithmic bytes per call: nnz*(4 + sizeof(a)) + (nrows+1)*4
//   + ncols*sizeof(u) + nrows*(sizeof(t) + 1).
#include "common.cuh"
#include <algorithm>
#include <type_traits>
#include <cub/device/device_radix_sort.cuh>

GrB_Info dev_exclusive_scan(int64_t *data, int64_t n, std::string *err);

static constexpr int SPMV_THREADS = 256;
static constexpr int SPMV_VEC = 4;                       // entries per 128-bit column load
static constexpr int SPMV_GROUPS = 2;                    // 128-bit loads per thread per array
static constexpr int SPMV_TILE = SPMV_THREADS * SPMV_VEC * SPMV_GROUPS;   // 2048 nnz per CTA

struct SpmvArgs {
    const uint32_t *rowptr; const uint32_t *col; const void *aval;
    const uint32_t *tile_row; int64_t ntiles; int64_t nrows; int64_t nnz;
    const void *uval; const uint8_t *upres;
    void *tval; uint8_t *tpres;
    void *head_val; uint8_t *head_has; void *tail_val; uint8_t *tail_has; int32_t *tail_row;
    int add_op, mul_op;
    int flip;      // 0: z = mul(a, u) (mxv)   1: z = mul(u, a) (vxm)
    int need_a;    // the multiply reads A's value
    int need_u;    // the multiply reads u's value
};

// ---- plan: tile_row[t] = row holding nnz t*TILE (tile_row[0] = 0, tile_row[ntiles] = nrows)
__global__ void spmv_plan_kernel(const uint32_t *rowptr, int64_t nrows, int64_t ntiles, uint32_t *tile_row) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == 0) { tile_row[0] = 0; return; }
    if (t == ntiles) { tile_row[t] = (uint32_t)nrows; return; }
    const uint32_t x = (uint32_t)(t * SPMV_TILE);
    int64_t lo = 0, hi = nrows;           // first r in [0, nrows] with rowptr[r] > x
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowptr[mid] > x) hi = mid; else lo = mid + 1; }
    tile_row[t] = (uint32_t)(lo - 1);
}

static GrB_Info spmv_plan(Csr &c, std::string *err) {
    if (c.tile_row) return GrB_SUCCESS;
    if (!c.rowptr32) return gb_fail(GrB_INVALID_VALUE, err, "mxv: matrices with >= 2^32 entries are not supported");
    c.ntiles = ceil_div(c.nnz, SPMV_TILE);
    GB_TRY(dalloc(&c.tile_row, (size_t)c.ntiles + 1, err));
    const int64_t n = c.ntiles + 1;
    spmv_plan_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, G.stream>>>(c.rowptr32, c.nrows, c.ntiles, c.tile_row); GB_LAUNCHED();
    CU_TRY(cudaGetLastError(), err);
    return GrB_SUCCESS;
}

// ---- 128-bit-granular loads of four consecutive entries
__device__ __forceinline__ uint4 ldg_stream128(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <typename T> __device__ __forceinline__ void load4(const T *p, T (&out)[4]) {
    if constexpr (sizeof(T) == 4) {
        const uint4 v = ldg_stream128(p);
        out[0] = reinterpret_cast<const T &>(v.x); out[1] = reinterpret_cast<const T &>(v.y);
        out[2] = reinterpret_cast<const T &>(v.z); out[3] = reinterpret_cast<const T &>(v.w);
    } else if constexpr (sizeof(T) == 8) {
        const uint4 v0 = ldg_stream128(p);
        const uint4 v1 = ldg_stream128(reinterpret_cast<const uint4 *>(p) + 1);
        uint64_t q[4] = {((uint64_t)v0.y << 32) | v0.x, ((uint64_t)v0.w << 32) | v0.z,
                         ((uint64_t)v1.y << 32) | v1.x, ((uint64_t)v1.w << 32) | v1.z};
        for (int k = 0; k < 4; ++k) out[k] = reinterpret_cast<const T &>(q[k]);
    } else if constexpr (sizeof(T) == 2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        uint16_t q[4] = {(uint16_t)(v.x & 0xffff), (uint16_t)(v.x >> 16), (uint16_t)(v.y & 0xffff), (uint16_t)(v.y >> 16)};
        for (int k = 0; k < 4; ++k) out[k] = reinterpret_cast<const T &>(q[k]);
    } else {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(p));
        uint8_t q[4] = {(uint8_t)(v & 0xff), (uint8_t)((v >> 8) & 0xff), (uint8_t)((v >> 16) & 0xff), (uint8_t)(v >> 24)};
        for (int k = 0; k < 4; ++k) out[k] = reinterpret_cast<const T &>(q[k]);
    }
}

template <typename T> __device__ __forceinline__ T gload(const T *p) {
    if constexpr (sizeof(T) == 1) { const unsigned char v = __ldg(reinterpret_cast<const unsigned char *>(p)); return reinterpret_cast<const T &>(v); }
    else return __ldg(p);
}

template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int o) {
    if constexpr (sizeof(T) == 8) {
        long long x = reinterpret_cast<long long &>(v);
        x = __shfl_xor_sync(0xffffffffu, x, o);
        return reinterpret_cast<T &>(x);
    } else if constexpr (sizeof(T) == 4) {
        int x = reinterpret_cast<int &>(v);
        x = __shfl_xor_sync(0xffffffffu, x, o);
        return reinterpret_cast<T &>(x);
    } else {
        int x = (int)v;
        x = __shfl_xor_sync(0xffffffffu, x, o);
        return (T)x;
    }
}

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 5); }   // breaks power-of-two strides

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// A partial monoid value: `has` says whether anything was folded in yet (identity-free, so that
// ANY and user-visible "no entry" semantics need no special cases).
template <typename ZT> struct Part { ZT v; int has; };
template <typename ZT> __device__ __forceinline__ Part<ZT> part_join(int add, Part<ZT> a, Part<ZT> b) {
    Part<ZT> r;
    r.has = a.has | b.has;
    r.v = a.has ? (b.has ? MulApply<ZT, ZT>::f(add, a.v, b.v) : a.v) : b.v;
    return r;
}
template <typename T> __device__ __forceinline__ T shfl_down_t(T v, int d) {
    if constexpr (sizeof(T) == 8) { long long x = reinterpret_cast<long long &>(v); x = __shfl_down_sync(0xffffffffu, x, d); return reinterpret_cast<T &>(x); }
    else if constexpr (sizeof(T) == 4) { int x = reinterpret_cast<int &>(v); x = __shfl_down_sync(0xffffffffu, x, d); return reinterpret_cast<T &>(x); }
    else { int x = (int)v; x = __shfl_down_sync(0xffffffffu, x, d); return (T)x; }
}

// Shared-memory working set of one 256-thread group processing one tile.
template <typename ZT> struct TileSmem {
    int32_t *headrow;      // [SPMV_TILE] row (relative to the tile's first row) starting at this entry, or -1
    ZT *wv; int *wflag;    // [8] per-warp scan aggregates: value, (has | stop << 1)
};
static constexpr int SPMV_WARPS = SPMV_THREADS / 32;
static constexpr int SPMV_ITEMS = SPMV_VEC * SPMV_GROUPS;       // consecutive entries per thread

// One tile of SPMV_TILE entries handled by a group of SPMV_THREADS threads (gtid = 0..255).
//
//   1. every thread streams its 8 consecutive entries (two 128-bit loads per array), gathers u and
//      keeps the 8 products in registers;
//   2. the rows of the tile are walked thread-per-row (coalesced rowptr reads): empty rows are
//      written out as "no entry", the others mark their first entry in shared memory;
//   3. item-centric segmented reduction: each thread folds its 8 items between row marks (rows
//      that begin and end inside a thread are final), then a segmented suffix scan over the
//      threads (shuffles inside a warp, 8 aggregates across warps) completes the rows that span
//      threads.  Work per thread is constant, whatever the row lengths: hub rows and runs of
//      short rows cost the same.
//   4. what sticks out of the tile goes to the per-tile head / tail slots for the fix-up kernel.
//
// HOT: columns are the relabelled ids of the hot-column plan; ids below hot_n are read from the
// shared-memory table s_hot instead of going through L1/L2.
template <typename XT, typename ZT, int ADD, int MUL, bool HOT>
__device__ __forceinline__ void spmv_tile_body(const SpmvArgs &p, const int64_t tile, const TileSmem<ZT> sm, const XT *s_hot,
                                               const uint32_t hot_n, const int gtid, const int bar_id) {
    const int add = ADD >= 0 ? ADD : p.add_op;
    const int mul = MUL >= 0 ? MUL : p.mul_op;
    const int64_t tstart = tile * SPMV_TILE;
    const int64_t tend = min(tstart + (int64_t)SPMV_TILE, p.nnz);
    const bool sparse_u = !HOT && p.upres != nullptr;
    const XT *aval = static_cast<const XT *>(p.aval);
    const XT *uval = static_cast<const XT *>(p.uval);
    ZT *tval = static_cast<ZT *>(p.tval);
    const int lane = gtid & 31, warp = gtid >> 5;
    const int loc0 = gtid * SPMV_ITEMS;
    const int64_t k0 = tstart + loc0;

    // ---- (1) stream: all loads of a phase are issued back to back (memory-level parallelism)
    uint32_t c[SPMV_ITEMS]; XT a[SPMV_ITEMS]; bool ok[SPMV_ITEMS];
    if (k0 + SPMV_ITEMS <= tend) {
#pragma unroll
        for (int g = 0; g < SPMV_GROUPS; ++g) {
            load4<uint32_t>(p.col + k0 + g * 4, *reinterpret_cast<uint32_t(*)[4]>(&c[g * 4]));
            if (p.need_a) load4<XT>(aval + k0 + g * 4, *reinterpret_cast<XT(*)[4]>(&a[g * 4]));
        }
#pragma unroll
        for (int j = 0; j < SPMV_ITEMS; ++j) ok[j] = true;
    } else {
#pragma unroll
        for (int j = 0; j < SPMV_ITEMS; ++j) {
            ok[j] = k0 + j < tend;
            c[j] = ok[j] ? p.col[k0 + j] : 0u;                    // 0 is always a valid column to gather
            if (p.need_a) a[j] = ok[j] ? aval[k0 + j] : (XT)1;
        }
    }
    const int64_t r0 = p.tile_row[tile];
    const int64_t r1 = min((int64_t)p.tile_row[tile + 1], p.nrows - 1);
    int64_t pre_rs = 0, pre_re = 0;                               // row pointers of this thread's first row
    if (r0 + gtid <= r1) { pre_rs = p.rowptr[r0 + gtid]; pre_re = p.rowptr[r0 + gtid + 1]; }
#pragma unroll
    for (int j = 0; j < SPMV_ITEMS; j += 4) *reinterpret_cast<int4 *>(&sm.headrow[loc0 + j]) = make_int4(-1, -1, -1, -1);
    if (gtid == 0) { p.tail_row[tile] = -1; p.head_has[tile] = 0; p.tail_has[tile] = 0; }
    uint8_t hs[SPMV_ITEMS]; XT uv[SPMV_ITEMS];
    if (sparse_u) {
#pragma unroll
        for (int j = 0; j < SPMV_ITEMS; ++j) hs[j] = __ldg(p.upres + c[j]);
    }
    if (p.need_u) {
#pragma unroll
        for (int j = 0; j < SPMV_ITEMS; ++j) {
            if (HOT && c[j] < hot_n) uv[j] = s_hot[c[j]];
            else uv[j] = gload<XT>(uval + c[j]);
        }
    }
    group_barrier(bar_id, SPMV_THREADS);                          // head marks are clear

    // ---- (2) rows of the tile: empty ones are final, the others mark their first entry
    for (int64_t r = r0 + gtid; r <= r1; r += SPMV_THREADS) {
        const bool first = r == r0 + gtid;
        const int64_t rs = first ? pre_rs : (int64_t)p.rowptr[r], re = first ? pre_re : (int64_t)p.rowptr[r + 1];
        if (rs == re) { p.tpres[r] = 0; tval[r] = (ZT)0; }
        else if (rs >= tstart && rs < tend) sm.headrow[rs - tstart] = (int32_t)(r - r0);
    }
    Part<ZT> prod[SPMV_ITEMS];
#pragma unroll
    for (int j = 0; j < SPMV_ITEMS; ++j) {
        const XT av = p.need_a ? a[j] : (XT)1;
        const XT uu = p.need_u ? uv[j] : (XT)1;
        prod[j].v = p.flip ? MulApply<XT, ZT>::f(mul, uu, av) : MulApply<XT, ZT>::f(mul, av, uu);
        prod[j].has = ok[j] && (!sparse_u || hs[j]);
    }
    group_barrier(bar_id, SPMV_THREADS);                          // head marks are complete

    // ---- (3a) fold this thread's items between row marks
    int32_t h[SPMV_ITEMS];
#pragma unroll
    for (int j = 0; j < SPMV_ITEMS; j += 4) {
        const int4 t4 = *reinterpret_cast<const int4 *>(&sm.headrow[loc0 + j]);
        h[j] = t4.x; h[j + 1] = t4.y; h[j + 2] = t4.z; h[j + 3] = t4.w;
    }
    Part<ZT> acc{(ZT)0, 0}, lead{(ZT)0, 0};
    bool seen = false; int32_t cur = -1;
#pragma unroll
    for (int j = 0; j < SPMV_ITEMS; ++j) {
        if (h[j] >= 0) {
            if (!seen) lead = acc;
            else { tval[r0 + cur] = acc.v; p.tpres[r0 + cur] = (uint8_t)acc.has; }   // row began and ended in this thread
            seen = true; cur = h[j]; acc.has = 0;
        }
        acc = part_join<ZT>(add, acc, prod[j]);
    }
    if (!seen) { lead = acc; acc.has = 0; }                        // no mark: everything continues an earlier row

    // ---- (3b) segmented suffix scan of the leads: X_t = lead_t (+) (stop_t ? nothing : X_{t+1}), stop = thread has a mark
    Part<ZT> x = lead; int stop = seen ? 1 : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Part<ZT> y; y.v = shfl_down_t<ZT>(x.v, d); y.has = __shfl_down_sync(0xffffffffu, x.has, d);
        const int ystop = __shfl_down_sync(0xffffffffu, stop, d);
        if (lane + d < 32) { if (!stop) x = part_join<ZT>(add, x, y); stop |= ystop; }
    }
    if (lane == 0) { sm.wv[warp] = x.v; sm.wflag[warp] = x.has | (stop << 1); }
    group_barrier(bar_id, SPMV_THREADS);
    // carry into this warp = what the following warps contribute to a segment still open at its end
    Part<ZT> carry{(ZT)0, 0}; int carry_stop = 0;
    for (int w = SPMV_WARPS - 1; w > warp; --w) {
        Part<ZT> y; y.v = sm.wv[w]; const int f = sm.wflag[w]; y.has = f & 1;
        const int ystop = f >> 1;
        if (ystop) { carry = y; carry_stop = 1; } else carry = part_join<ZT>(add, y, carry);
    }
    if (!stop) { x = part_join<ZT>(add, x, carry); stop |= carry_stop; }
    // S_t = X_{t+1}: what follows this thread's open row
    Part<ZT> nxt; nxt.v = shfl_down_t<ZT>(x.v, 1); nxt.has = __shfl_down_sync(0xffffffffu, x.has, 1);
    int nxt_stop = __shfl_down_sync(0xffffffffu, stop, 1);
    if (lane == 31) { nxt = carry; nxt_stop = carry_stop; }

    // ---- (4) rows still open at the end of a thread, and what sticks out of the tile
    if (seen) {
        const Part<ZT> total = part_join<ZT>(add, acc, nxt);
        if (nxt_stop) { tval[r0 + cur] = total.v; p.tpres[r0 + cur] = (uint8_t)total.has; }
        else { static_cast<ZT *>(p.tail_val)[tile] = total.v; p.tail_has[tile] = (uint8_t)total.has; p.tail_row[tile] = (int32_t)(r0 + cur); }
    }
    if (gtid == 0 && h[0] < 0) {                                   // the tile starts inside a row of an earlier tile
        static_cast<ZT *>(p.head_val)[tile] = x.v; p.head_has[tile] = (uint8_t)x.has;
    }
}

// One CTA per tile (general path: any u, any semiring).
template <typename XT, typename ZT, int ADD, int MUL>
__global__ void __launch_bounds__(SPMV_THREADS, 5) spmv_tile_kernel(const SpmvArgs p) {
    __shared__ __align__(16) int32_t s_headrow[SPMV_TILE];
    __shared__ ZT s_wv[SPMV_WARPS];
    __shared__ int s_wflag[SPMV_WARPS];
    const TileSmem<ZT> sm{s_headrow, s_wv, s_wflag};
    spmv_tile_body<XT, ZT, ADD, MUL, false>(p, blockIdx.x, sm, nullptr, 0u, threadIdx.x, 0);
}

// Persistent variant for dense u on large matrices: HOT_GROUPS independent 256-thread groups per CTA
// stride over the tiles, all sharing a shared-memory table with the u values of the hot_n most
// frequently referenced columns (the matrix's columns were relabelled by descending in-degree, so
// "hot" is simply "id < hot_n").  On R-MAT graphs half of all gathers hit the table.
template <typename ZT> __host__ __device__ constexpr size_t hot_group_bytes() {
    return ((SPMV_TILE * sizeof(int32_t) + SPMV_WARPS * (sizeof(ZT) + sizeof(int))) + 15) & ~(size_t)15;
}
template <typename XT, typename ZT, int ADD, int MUL, int HOT_GROUPS>
__global__ void __launch_bounds__(SPMV_THREADS * HOT_GROUPS) spmv_hot_kernel(const SpmvArgs p, const uint32_t hot_n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int group = threadIdx.x / SPMV_THREADS, gtid = threadIdx.x % SPMV_THREADS;
    unsigned char *gbase = smem_raw + (size_t)group * hot_group_bytes<ZT>();
    TileSmem<ZT> sm;
    sm.headrow = reinterpret_cast<int32_t *>(gbase);
    sm.wv = reinterpret_cast<ZT *>(gbase + SPMV_TILE * sizeof(int32_t));
    sm.wflag = reinterpret_cast<int *>(gbase + SPMV_TILE * sizeof(int32_t) + SPMV_WARPS * sizeof(ZT));
    XT *s_hot = reinterpret_cast<XT *>(smem_raw + HOT_GROUPS * hot_group_bytes<ZT>());
    const XT *uval = static_cast<const XT *>(p.uval);
    for (uint32_t i = threadIdx.x; i < hot_n; i += blockDim.x) s_hot[i] = uval[i];
    __syncthreads();
    for (int64_t tile = (int64_t)blockIdx.x * HOT_GROUPS + group; tile < p.ntiles; tile += (int64_t)gridDim.x * HOT_GROUPS) {
        spmv_tile_body<XT, ZT, ADD, MUL, true>(p, tile, sm, s_hot, hot_n, gtid, group + 1);
        group_barrier(group + 1, SPMV_THREADS);     // the group's smem is reused by its next tile
    }
}

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

// ---- fix-up: rows that straddle tiles = tail partial of the tile they start in
//      (+) head partials of the following tiles, combined by one warp in a fixed order
template <typename ZT, int ADD>
__global__ void __launch_bounds__(256) spmv_fixup_kernel(const SpmvArgs p) {
    const int add = ADD >= 0 ? ADD : p.add_op;
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= p.ntiles) return;
    const int32_t r = p.tail_row[t];
    if (r < 0) return;
    const ZT ident = monoid_identity<ZT>(add);
    const int64_t re = p.rowptr[r + 1];
    const int64_t last_tile = (re - 1) / SPMV_TILE;
    ZT acc = ident; int has = 0;
    if (lane == 0 && p.tail_has[t]) { acc = static_cast<const ZT *>(p.tail_val)[t]; has = 1; }
    for (int64_t tt = t + 1 + lane; tt <= last_tile; tt += 32) if (p.head_has[tt]) {
        const ZT v = static_cast<const ZT *>(p.head_val)[tt];
        acc = has ? MulApply<ZT, ZT>::f(add, acc, v) : v; has = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const ZT ov = shfl_xor_t<ZT>(acc, o);
        const int oh = __shfl_xor_sync(0xffffffffu, has, o);
        if (oh) { acc = has ? MulApply<ZT, ZT>::f(add, acc, ov) : ov; has = 1; }
    }
    if (lane == 0) { static_cast<ZT *>(p.tval)[r] = acc; p.tpres[r] = (uint8_t)(has != 0); }
}

__global__ void clear_presence_kernel(uint8_t *p, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) p[k] = 0;
}

struct HotLaunch { bool on; int64_t hused; int groups; size_t table_bytes; };

template <typename XT, typename ZT, int ADD, int MUL, int GROUPS>
static void spmv_hot_launch(const SpmvArgs &a, const HotLaunch &h) {
    auto kernel = spmv_hot_kernel<XT, ZT, ADD, MUL, GROUPS>;
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, G.device);
    const size_t fixed = GROUPS * hot_group_bytes<ZT>();
    size_t avail = (size_t)max_optin > fixed + 1024 ? (size_t)max_optin - fixed - 1024 : 0;
    avail = std::min(avail, h.table_bytes);
    const uint32_t hot_n = (uint32_t)std::min<int64_t>(h.hused, (int64_t)(avail / sizeof(XT)));
    const size_t smem = fixed + (size_t)hot_n * sizeof(XT);
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SPMV_THREADS * GROUPS, smem);
    per_sm = std::max(per_sm, 1);
    const int ctas = (int)std::min<int64_t>((int64_t)G.num_sms * per_sm, ceil_div(a.ntiles, GROUPS));
    kernel<<<ctas, SPMV_THREADS * GROUPS, smem, G.stream>>>(a, hot_n); GB_LAUNCHED();
}

template <typename XT, typename ZT, int ADD, int MUL>
static void spmv_launch(const SpmvArgs &a, const HotLaunch &h) {
    bool launched = false;
    if constexpr (std::is_same<XT, ZT>::value) {
        if (h.on) { launched = true; if (h.groups == 2) spmv_hot_launch<XT, ZT, ADD, MUL, 2>(a, h); else spmv_hot_launch<XT, ZT, ADD, MUL, 4>(a, h); }
    }
    if (!launched) { spmv_tile_kernel<XT, ZT, ADD, MUL><<<(unsigned)a.ntiles, SPMV_THREADS, 0, G.stream>>>(a); GB_LAUNCHED(); }
    spmv_fixup_kernel<ZT, ADD><<<(unsigned)ceil_div(a.ntiles * 32, 256), 256, 0, G.stream>>>(a); GB_LAUNCHED();
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
__global__ void hot_relabel_kernel(const uint32_t *col, const uint32_t *inv, int64_t nnz, uint32_t *out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) out[k] = inv[col[k]];
}
static inline int hgrid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (int64_t)G.num_sms * 16)); }

static GrB_Info spmv_hot_plan(Csr &c, std::string *err) {
    if (c.hcol) return GrB_SUCCESS;
    const int64_t n = c.ncols;
    uint32_t *deg = nullptr, *deg_sorted = nullptr, *ids = nullptr, *inv = nullptr; unsigned long long *used = nullptr;
    GB_TRY(dalloc(&deg, (size_t)n, err)); GB_TRY(dalloc(&deg_sorted, (size_t)n, err)); GB_TRY(dalloc(&ids, (size_t)n, err));
    GB_TRY(dalloc(&inv, (size_t)n, err)); GB_TRY(dalloc(&used, 1, err));
    GB_TRY(dalloc(&c.hperm, (size_t)n, err));
    GB_TRY(dalloc(&c.hcol, (size_t)c.nnz, err));
    CU_TRY(cudaMemsetAsync(deg, 0, (size_t)n * 4, G.stream), err);
    CU_TRY(cudaMemsetAsync(used, 0, 8, G.stream), err);
    hot_count_kernel<<<hgrid(c.nnz), 256, 0, G.stream>>>(c.col, c.nnz, deg); GB_LAUNCHED();
    hot_iota_kernel<<<hgrid(n), 256, 0, G.stream>>>(ids, n); GB_LAUNCHED();
    size_t tmp_bytes = 0;     // stable sort: equal degrees keep ascending column order (deterministic plan)
    CU_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, deg, deg_sorted, ids, c.hperm, n, 0, 32, G.stream), err);
    void *tmp = nullptr; GB_TRY(dmalloc(&tmp, tmp_bytes, err));
    CU_TRY(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, deg, deg_sorted, ids, c.hperm, n, 0, 32, G.stream), err);
    G.launches += 8;
    hot_invert_kernel<<<hgrid(n), 256, 0, G.stream>>>(c.hperm, deg_sorted, n, inv, used); GB_LAUNCHED();
    hot_relabel_kernel<<<hgrid(c.nnz), 256, 0, G.stream>>>(c.col, inv, c.nnz, c.hcol); GB_LAUNCHED();
    unsigned long long h = 0;
    CU_TRY(cudaMemcpyAsync(&h, used, 8, cudaMemcpyDeviceToHost, G.stream), err);
    CU_TRY(cudaStreamSynchronize(G.stream), err);
    c.hused = (int64_t)h;
    dfree(tmp); dfree(deg); dfree(deg_sorted); dfree(ids); dfree(inv); dfree(used);
    return GrB_SUCCESS;
}

// compile-time specialised semirings (BASELINE.json north_star: PLUS_TIMES, LOR_LAND, MIN_PLUS,
// PLUS_SECOND; plus PLUS_PAIR / ANY_PAIR / PLUS_FIRST / MIN_FIRST / MIN_SECOND which the reference's
// demos use); everything else runs the same kernel with run-time operator codes.
template <typename T> static bool spmv_fast(int add, int mul, const SpmvArgs &a, const HotLaunch &h) {
#define GB_FAST(A, M) if (add == A && mul == M) { spmv_launch<T, T, A, M>(a, h); return true; }
    GB_FAST(OP_PLUS, OP_TIMES) GB_FAST(OP_MIN, OP_PLUS) GB_FAST(OP_PLUS, OP_SECOND) GB_FAST(OP_PLUS, OP_FIRST)
    GB_FAST(OP_PLUS, OP_PAIR) GB_FAST(OP_MIN, OP_FIRST) GB_FAST(OP_MIN, OP_SECOND)
#undef GB_FAST
    return false;
}
static bool spmv_fast_bool(int add, int mul, const SpmvArgs &a, const HotLaunch &h) {
#define GB_FAST(A, M) if (add == A && mul == M) { spmv_launch<bool, bool, A, M>(a, h); return true; }
    GB_FAST(OP_LOR, OP_LAND) GB_FAST(OP_ANY, OP_PAIR) GB_FAST(OP_LOR, OP_PAIR) GB_FAST(OP_LOR, OP_SECOND) GB_FAST(OP_LOR, OP_FIRST)
#undef GB_FAST
    return false;
}

static GrB_Info spmv_dispatch(int xt, int zt, int add, int mul, const SpmvArgs &a, const HotLaunch &h, std::string *err) {
    if (xt == zt) {
        switch (xt) {
            case TC_FP32:  if (spmv_fast<float>(add, mul, a, h)) return GrB_SUCCESS; break;
            case TC_FP64:  if (spmv_fast<double>(add, mul, a, h)) return GrB_SUCCESS; break;
            case TC_INT32: if (spmv_fast<int32_t>(add, mul, a, h)) return GrB_SUCCESS; break;
            case TC_INT64: if (spmv_fast<int64_t>(add, mul, a, h)) return GrB_SUCCESS; break;
            case TC_UINT32: if (spmv_fast<uint32_t>(add, mul, a, h)) return GrB_SUCCESS; break;
            case TC_UINT64: if (spmv_fast<uint64_t>(add, mul, a, h)) return GrB_SUCCESS; break;
            case TC_BOOL:  if (spmv_fast_bool(add, mul, a, h)) return GrB_SUCCESS; break;
            default: break;
        }
        switch (xt) {
#define GB_GEN(TC, T) case TC: spmv_launch<T, T, -1, -1>(a, h); return GrB_SUCCESS;
            GB_GEN(TC_BOOL, bool) GB_GEN(TC_INT8, int8_t) GB_GEN(TC_INT16, int16_t) GB_GEN(TC_INT32, int32_t) GB_GEN(TC_INT64, int64_t)
            GB_GEN(TC_UINT8, uint8_t) GB_GEN(TC_UINT16, uint16_t) GB_GEN(TC_UINT32, uint32_t) GB_GEN(TC_UINT64, uint64_t)
            GB_GEN(TC_FP32, float) GB_GEN(TC_FP64, double)
#undef GB_GEN
        }
    } else if (zt == TC_BOOL) {
        switch (xt) {
#define GB_GEN(TC, T) case TC: spmv_launch<T, bool, -1, -1>(a, h); return GrB_SUCCESS;
            GB_GEN(TC_INT8, int8_t) GB_GEN(TC_INT16, int16_t) GB_GEN(TC_INT32, int32_t) GB_GEN(TC_INT64, int64_t)
            GB_GEN(TC_UINT8, uint8_t) GB_GEN(TC_UINT16, uint16_t) GB_GEN(TC_UINT32, uint32_t) GB_GEN(TC_UINT64, uint64_t)
            GB_GEN(TC_FP32, float) GB_GEN(TC_FP64, double)
#undef GB_GEN
        }
    }
    return gb_fail(GrB_DOMAIN_MISMATCH, err, "mxv: unsupported semiring domains (x=%d, z=%d)", xt, zt);
}

// ------------------------------------------------------------------ finalize:  w<mask> = accum(w, t)
struct VecFinalizeArgs {
    int64_t n;
    const void *wval; const uint8_t *wpres; int wtc; int w_exists;
    const void *tval; const uint8_t *tpres; int ttc;
    const void *mval; const uint8_t *mpres; int mtc; int has_mask, mask_comp, mask_struct, replace;
    int accum_op, accum_tc, accum_ztc;   // accum_op < 0: none
    void *oval; uint8_t *opres;
};
__global__ void vec_finalize_kernel(const VecFinalizeArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool tp = a.tpres ? a.tpres[i] != 0 : true;
        const bool wp = a.w_exists ? (a.wpres ? a.wpres[i] != 0 : true) : false;
        bool m = true;
        if (a.has_mask) {
            m = a.mpres ? a.mpres[i] != 0 : true;
            if (m && !a.mask_struct) { const Sc mv = sc_cast(sc_load(a.mtc, a.mval, i), a.mtc, TC_BOOL); m = mv.u != 0; }
            if (a.mask_comp) m = !m;
        }
        Sc out; out.u = 0; bool op = false;
        if (m) {
            if (a.accum_op >= 0) {
                if (wp && tp) {
                    const Sc x = sc_cast(sc_load(a.wtc, a.wval, i), a.wtc, a.accum_tc);
                    const Sc y = sc_cast(sc_load(a.ttc, a.tval, i), a.ttc, a.accum_tc);
                    out = sc_cast(sc_binop(a.accum_op, a.accum_tc, x, y), a.accum_ztc, a.wtc); op = true;
                } else if (wp) { out = sc_load(a.wtc, a.wval, i); op = true; }
                else if (tp) { out = sc_cast(sc_load(a.ttc, a.tval, i), a.ttc, a.wtc); op = true; }
            } else if (tp) { out = sc_cast(sc_load(a.ttc, a.tval, i), a.ttc, a.wtc); op = true; }
        } else if (!a.replace && wp) { out = sc_load(a.wtc, a.wval, i); op = true; }
        if (op) sc_store(a.wtc, a.oval, i, out);
        a.opres[i] = op;
    }
}

static inline int grid_for(int64_t n, int threads = 256) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, threads), (int64_t)G.num_sms * 16));
}

static bool op_uses_x(int op) { return !(op == OP_SECOND || op == OP_PAIR); }
static bool op_uses_y(int op) { return !(op == OP_FIRST || op == OP_PAIR || op == OP_ANY); }

// w<mask> = accum(w, A' (+).(x) u) with `flip` selecting mul(u,a) (vxm) and `use_transpose`
// selecting the cached CSR of A' (so that the kernel always pulls along CSR rows).
static GrB_Info mxv_core(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Semiring s,
                         const GrB_Matrix A, const GrB_Vector u, const DescFlags &f, bool use_transpose, bool flip,
                         const char *fn) {
    std::string *err = &w->err;
    // ---- domain / dimension checks (host, synchronous)
    const GrB_BinaryOp mulop = s->mul; const GrB_BinaryOp addop = s->add->op;
    if (mulop->opcode == OP_USER || addop->opcode == OP_USER || (accum && accum->opcode == OP_USER))
        return gb_fail(GrB_INVALID_VALUE, err, "%s: user-defined operators are host function pointers and cannot run on the GPU (no CPU fallback)", fn);
    const uint64_t out_n = use_transpose ? A->ncols : A->nrows, in_n = use_transpose ? A->nrows : A->ncols;
    if (u->n != in_n || w->n != out_n || (mask && mask->n != out_n))
        return gb_fail(GrB_DIMENSION_MISMATCH, err, "%s: dimensions do not match (A is %llux%llu%s, u %llu, w %llu)", fn,
                       (unsigned long long)A->nrows, (unsigned long long)A->ncols, use_transpose ? " transposed" : "",
                       (unsigned long long)u->n, (unsigned long long)w->n);
    if (!G.have_device) return gb_fail(GrB_PANIC, err, "%s: no CUDA device: libb200grb computes only on the GPU (no CPU fallback)", fn);
    const int xt = mulop->xtype->code, zt = addop->ztype->code;
    const int add = addop->opcode, mul = mulop->opcode;

    // ---- operands in HBM
    if (use_transpose) GB_TRY(matrix_ensure_transpose(A)); else GB_TRY(matrix_ensure_device(A));
    Csr &c = use_transpose ? A->devT : A->dev;
    GB_TRY(vector_ensure_device(u));
    if (mask) GB_TRY(vector_ensure_device(mask));
    const bool need_final = mask != nullptr || accum != nullptr;
    const bool w_empty = w->host_valid && w->hi.empty() && w->pi.empty();   // nothing to merge with
    if (need_final && !w_empty) GB_TRY(vector_ensure_device(w));
    GB_TRY(spmv_plan(c, err));

    const bool a_is_x = !flip;
    const bool need_a = a_is_x ? op_uses_x(mul) : op_uses_y(mul);
    const bool need_u = a_is_x ? op_uses_y(mul) : op_uses_x(mul);
    void *a_cast = nullptr, *u_cast = nullptr;
    const void *aval = c.val, *uval = u->dval;
    if (need_a && A->type->code != xt) { GB_TRY(dev_cast_values(&a_cast, xt, c.val, A->type->code, c.nnz, err)); aval = a_cast; }
    if (need_u && u->type->code != xt) { GB_TRY(dev_cast_values(&u_cast, xt, u->dval, u->type->code, (int64_t)u->n, err)); uval = u_cast; }

    const int64_t n = (int64_t)out_n;
    const size_t zsz = (size_t)tc_size(zt);
    void *tval = nullptr; uint8_t *tpres = nullptr;
    GB_TRY(dmalloc(&tval, (size_t)n * zsz + 16, err));
    GB_TRY(dmalloc((void **)&tpres, (size_t)n + 16, err));

    if (c.nnz == 0) {
        clear_presence_kernel<<<grid_for(n), 256, 0, G.stream>>>(tpres, n); GB_LAUNCHED();
    } else {
        SpmvArgs a{};
        a.rowptr = c.rowptr32; a.col = c.col; a.aval = aval; a.tile_row = c.tile_row; a.ntiles = c.ntiles;
        a.nrows = c.nrows; a.nnz = c.nnz; a.uval = uval; a.upres = u->dpres; a.tval = tval; a.tpres = tpres;
        a.add_op = add; a.mul_op = mul; a.flip = flip; a.need_a = need_a; a.need_u = need_u;
        GB_TRY(dmalloc(&a.head_val, (size_t)c.ntiles * zsz + 16, err));
        GB_TRY(dmalloc(&a.tail_val, (size_t)c.ntiles * zsz + 16, err));
        GB_TRY(dmalloc((void **)&a.head_has, (size_t)c.ntiles + 16, err));
        GB_TRY(dmalloc((void **)&a.tail_has, (size_t)c.ntiles + 16, err));
        GB_TRY(dalloc(&a.tail_row, (size_t)c.ntiles, err));
        // dense u on a large matrix: hot-column plan + shared-memory table (see spmv_hot_kernel)
        HotLaunch hot{false, 0, 4, (size_t)64 << 10};
        if (const char *e = getenv("B200GRB_HOT_GROUPS")) hot.groups = atoi(e);
        if (const char *e = getenv("B200GRB_HOT_KB")) hot.table_bytes = (size_t)atoi(e) << 10;
        void *u_perm = nullptr;
        const bool no_hot = getenv("B200GRB_NO_HOT") != nullptr;
        if (!no_hot && need_u && !u->dpres && xt == zt && c.nnz >= ((int64_t)1 << 20) && c.ncols >= (1 << 16)) {
            GB_TRY(spmv_hot_plan(c, err));
            const size_t xsz = (size_t)tc_size(xt);
            GB_TRY(dmalloc(&u_perm, (size_t)c.hused * xsz + 16, err));
            if (c.hused > 0) { permute_u_kernel<<<hgrid(c.hused), 256, 0, G.stream>>>(c.hperm, (const uint8_t *)uval, (uint8_t *)u_perm, (int)xsz, c.hused); GB_LAUNCHED(); }
            a.col = c.hcol; a.uval = u_perm;
            hot.on = getenv("B200GRB_RELABEL_ONLY") == nullptr; hot.hused = c.hused;
        }
        GrB_Info r = spmv_dispatch(xt, zt, add, mul, a, hot, err);
        dfree(u_perm);
        dfree(a.head_val); dfree(a.tail_val); dfree(a.head_has); dfree(a.tail_has); dfree(a.tail_row);
        if (r != GrB_SUCCESS) { dfree(tval); dfree(tpres); dfree(a_cast); dfree(u_cast); return r; }
    }
    dfree(a_cast); dfree(u_cast);

    // ---- w<mask> = accum(w, t)
    const int wtc = w->type->code;
    if (!need_final) {
        if (wtc == zt) vector_adopt_device(w, tval, tpres);
        else {
            void *cv = nullptr;
            GB_TRY(dev_cast_values(&cv, wtc, tval, zt, n, err));
            dfree(tval);
            vector_adopt_device(w, cv, tpres);
        }
    } else {
        VecFinalizeArgs fa{};
        fa.n = n; fa.wval = w->dval; fa.wpres = w->dpres; fa.wtc = wtc; fa.w_exists = w_empty ? 0 : 1;
        fa.tval = tval; fa.tpres = tpres; fa.ttc = zt;
        if (mask) { fa.mval = mask->dval; fa.mpres = mask->dpres; fa.mtc = mask->type->code; fa.has_mask = 1; }
        fa.mask_comp = f.mask_comp; fa.mask_struct = f.mask_struct; fa.replace = f.replace;
        fa.accum_op = accum ? accum->opcode : -1;
        fa.accum_tc = accum ? accum->xtype->code : 0; fa.accum_ztc = accum ? accum->ztype->code : 0;
        GB_TRY(dmalloc(&fa.oval, (size_t)n * tc_size(wtc) + 16, err));
        GB_TRY(dmalloc((void **)&fa.opres, (size_t)n + 16, err));
        vec_finalize_kernel<<<grid_for(n), 256, 0, G.stream>>>(fa); GB_LAUNCHED();
        dfree(tval); dfree(tpres);
        vector_adopt_device(w, fa.oval, fa.opres);
    }
    CU_TRY(cudaGetLastError(), err);
    return GrB_SUCCESS;
}

static GrB_Info mxv_check(GrB_Vector w, const GrB_Vector mask, const GrB_Semiring s, const GrB_Matrix A, const GrB_Vector u, const char *fn) {
    if (!w || !s || !A || !u) return gb_fail(GrB_NULL_POINTER, nullptr, "%s: NULL argument", fn);
    if (!gb_valid_vector(w) || !gb_valid_vector(u) || !gb_valid_matrix(A) || (mask && !gb_valid_vector(mask)) || s->magic != GB_MAGIC)
        return gb_fail(GrB_UNINITIALIZED_OBJECT, nullptr, "%s: invalid object", fn);
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_mxv(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Semiring semiring,
                            const GrB_Matrix A, const GrB_Vector u, const GrB_Descriptor desc) {
    GB_LOCK; GB_CHECK_INIT;
    GB_TRY(mxv_check(w, mask, semiring, A, u, "GrB_mxv"));
    const DescFlags f = desc_flags(desc);
    return mxv_core(w, mask, accum, semiring, A, u, f, /*use_transpose=*/f.tran0, /*flip=*/false, "GrB_mxv");
}

extern "C" GrB_Info GrB_vxm(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Semiring semiring,
                            const GrB_Vector u, const GrB_Matrix A, const GrB_Descriptor desc) {
    GB_LOCK; GB_CHECK_INIT;
    GB_TRY(mxv_check(w, mask, semiring, A, u, "GrB_vxm"));
    const DescFlags f = desc_flags(desc);
    // w' = u'A  <=>  w = A'u: pull along the rows of A' (INP1 = TRAN cancels the transpose)
    return mxv_core(w, mask, accum, semiring, A, u, f, /*use_transpose=*/!f.tran1, /*flip=*/true, "GrB_vxm");
}
