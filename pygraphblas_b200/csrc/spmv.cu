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

// build with -DB200GRB_PHASE_TIMERS=1 and run with B200GRB_SPMV_DEBUG=1 to get per-phase cycle counts of the tile kernel
#ifndef B200GRB_PHASE_TIMERS
#define B200GRB_PHASE_TIMERS 0
#endif
static constexpr bool SPMV_PHASE_TIMERS = B200GRB_PHASE_TIMERS != 0;
static constexpr int SPMV_THREADS = 256;
static constexpr int SPMV_WARPS = SPMV_THREADS / 32;

struct SpmvArgs {
    const uint32_t *rowptr; const uint32_t *col; const void *aval;
    const uint32_t *tile_row; int64_t ntiles; int64_t nrows; int64_t nnz;
    const void *uval; const uint8_t *upres;
    void *tval; uint8_t *tpres;
    void *head_val; uint8_t *head_has; void *tail_val; uint8_t *tail_has; int32_t *tail_row;
    int add_op, mul_op;
    int flip;      // 0: z = mul(a, u) (mxv)   1: z = mul(u, a) (vxm)   (run-time-operator kernels only)
    int tile;      // entries per tile = SPMV_THREADS * items per thread
    unsigned long long *dbg;   // optional per-phase cycle counters (B200GRB_SPMV_DEBUG), nullptr in production
};

// ---- plan: tile_row[t] = row holding entry t*tile (tile_row[0] = 0, tile_row[ntiles] = nrows)
__global__ void spmv_plan_kernel(const uint32_t *rowptr, int64_t nrows, int64_t ntiles, int tile, uint32_t *tile_row) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == 0) { tile_row[0] = 0; return; }
    if (t == ntiles) { tile_row[t] = (uint32_t)nrows; return; }
    const uint32_t x = (uint32_t)(t * tile);
    int64_t lo = 0, hi = nrows;           // first r in [0, nrows] with rowptr[r] > x
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowptr[mid] > x) hi = mid; else lo = mid + 1; }
    tile_row[t] = (uint32_t)(lo - 1);
}

static GrB_Info spmv_plan(Csr &c, int tile, std::string *err) {
    if (c.tile_row && c.tile_size == tile) return GrB_SUCCESS;
    if (!c.rowptr32) return gb_fail(GrB_INVALID_VALUE, err, "mxv: matrices with >= 2^32 entries are not supported");
    dfree(c.tile_row); c.tile_row = nullptr;
    c.ntiles = ceil_div(c.nnz, tile); c.tile_size = tile;
    GB_TRY(dalloc(&c.tile_row, (size_t)c.ntiles + 1, err));
    const int64_t n = c.ntiles + 1;
    spmv_plan_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, G.stream>>>(c.rowptr32, c.nrows, c.ntiles, tile, c.tile_row); GB_LAUNCHED();
    CU_TRY(cudaGetLastError(), err);
    return GrB_SUCCESS;
}

// ---- 128-bit streaming loads of four consecutive entries (no L1 allocation: L1 is kept for u)
__device__ __forceinline__ uint4 ldg_stream128(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <typename T> __device__ __forceinline__ void load4(const T *p, T *out) {
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
    if constexpr (sizeof(T) == 8) { long long x = reinterpret_cast<long long &>(v); x = __shfl_xor_sync(0xffffffffu, x, o); return reinterpret_cast<T &>(x); }
    else if constexpr (sizeof(T) == 4) { int x = reinterpret_cast<int &>(v); x = __shfl_xor_sync(0xffffffffu, x, o); return reinterpret_cast<T &>(x); }
    else { int x = (int)v; x = __shfl_xor_sync(0xffffffffu, x, o); return (T)x; }
}
template <typename T> __device__ __forceinline__ T shfl_down_t(T v, int d) {
    if constexpr (sizeof(T) == 8) { long long x = reinterpret_cast<long long &>(v); x = __shfl_down_sync(0xffffffffu, x, d); return reinterpret_cast<T &>(x); }
    else if constexpr (sizeof(T) == 4) { int x = reinterpret_cast<int &>(v); x = __shfl_down_sync(0xffffffffu, x, d); return reinterpret_cast<T &>(x); }
    else { int x = (int)v; x = __shfl_down_sync(0xffffffffu, x, d); return (T)x; }
}

// A partial monoid value: `has` says whether anything was folded in yet (identity-free, so that
// ANY and "no entry" need no special cases).
template <typename ZT> struct Part { ZT v; int has; };
template <typename ZT> __device__ __forceinline__ Part<ZT> part_join(int add, Part<ZT> a, Part<ZT> b) {
    Part<ZT> r;
    r.has = a.has | b.has;
    r.v = a.has ? (b.has ? MulApply<ZT, ZT>::f(add, a.v, b.v) : a.v) : b.v;
    return r;
}

// which operands a multiply reads (compile-time for the specialised semirings)
__host__ __device__ constexpr bool mul_reads_x(int op) { return !(op == OP_SECOND || op == OP_PAIR); }
__host__ __device__ constexpr bool mul_reads_y(int op) { return !(op == OP_FIRST || op == OP_PAIR || op == OP_ANY); }

// One tile of SPMV_THREADS * IT consecutive entries per CTA:
//
//   1. every thread streams its IT consecutive entries (128-bit loads), issues all its gathers of u
//      back to back and keeps the products in registers;
//   2. the rows of the tile are walked thread-per-row (coalesced rowptr reads): empty rows are
//      written out as "no entry", the others mark their first entry in shared memory;
//   3. item-centric segmented reduction: each thread folds its items between row marks (rows that
//      begin and end inside a thread are final), then a segmented suffix scan over the threads
//      (shuffles inside a warp, 8 aggregates across warps) completes the rows that span threads.
//      Work per thread is constant whatever the row lengths: hub rows and runs of short rows cost
//      the same;
//   4. what sticks out of the tile goes to the per-tile head / tail slots for the fix-up kernel.
//
// ADD/MUL >= 0: compile-time semiring (flip already folded into MUL by the host); -1: run-time codes.
// SPARSE: u has a presence array (entries with absent u(k) do not contribute).
// Everything after the column / value words of a tile are in registers: gather u, mark rows, fold,
// scan, write.  `sync` is the barrier of the 256 threads that share s_head / s_wv / s_wflag.
template <typename XT, typename ZT, int ADD, int MUL, bool SPARSE, int IT, bool HOT, typename Sync>
__device__ __forceinline__ void spmv_tile_finish(const SpmvArgs &p, const uint32_t tile, const int tid, const int tlen, const int nvalid,
                                                 uint32_t (&c)[IT], XT (&a)[IT], int32_t *s_head, ZT *s_wv, int *s_wflag,
                                                 const XT *s_hot, const uint32_t hot_n, Sync &&sync) {
    constexpr int TILE = SPMV_THREADS * IT;
    constexpr bool NEED_A = MUL < 0 || mul_reads_x(MUL);
    constexpr bool NEED_U = MUL < 0 || mul_reads_y(MUL);
    const int add = ADD >= 0 ? ADD : p.add_op;
    const int mul = MUL >= 0 ? MUL : p.mul_op;
    const int lane = tid & 31, warp = tid >> 5;
    const int loc0 = tid * IT;
    const int64_t tstart = (int64_t)tile * TILE;
    const XT *uval = static_cast<const XT *>(p.uval);
    ZT *tval = static_cast<ZT *>(p.tval);

    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (SPMV_PHASE_TIMERS && p.dbg && tid == 0) t0 = clock64();
    const uint32_t r0 = p.tile_row[tile];
    const uint32_t r1 = (uint32_t)min((int64_t)p.tile_row[tile + 1], p.nrows - 1);
    uint32_t pre_rs = 0, pre_re = 0;                                      // row pointers of this thread's first row
    if (r0 + tid <= r1) { pre_rs = p.rowptr[r0 + tid]; pre_re = p.rowptr[r0 + tid + 1]; }
#pragma unroll
    for (int j = 0; j < IT; j += 4) *reinterpret_cast<int4 *>(&s_head[loc0 + j]) = make_int4(-1, -1, -1, -1);
    if (tid == 0) { p.tail_row[tile] = -1; p.head_has[tile] = 0; p.tail_has[tile] = 0; }
    uint8_t hs[SPARSE ? IT : 1]; XT uv[IT];
    if (SPARSE) {
#pragma unroll
        for (int j = 0; j < IT; ++j) hs[j] = __ldg(p.upres + c[j]);
    }
    if (NEED_U) {
#pragma unroll
        for (int j = 0; j < IT; ++j) {
            if (HOT && c[j] < hot_n) uv[j] = s_hot[c[j]];                     // hot column: shared-memory table, no L1 wavefront
            else uv[j] = gload<XT>(uval + c[j]);
        }
    }
    sync();                                                               // head marks are clear
    if (SPMV_PHASE_TIMERS && p.dbg && tid == 0) t1 = clock64();

    // ---- (2) rows of the tile: empty ones are final, the others mark their first entry
    {
        const uint32_t ts32 = (uint32_t)tstart;                           // nnz < 2^32
        uint32_t rs = pre_rs, re = pre_re;
        for (uint32_t r = r0 + tid; r <= r1; r += SPMV_THREADS) {
            if (r != r0 + tid) { rs = p.rowptr[r]; re = p.rowptr[r + 1]; }
            if (rs == re) { p.tpres[r] = 0; tval[r] = (ZT)0; }
            else if (rs >= ts32 && rs - ts32 < (uint32_t)tlen) s_head[rs - ts32] = (int32_t)(r - r0);
        }
    }
    ZT prod[IT];
#pragma unroll
    for (int j = 0; j < IT; ++j) {
        const XT av = NEED_A ? a[j] : (XT)1;
        const XT uu = NEED_U ? uv[j] : (XT)1;
        if (MUL >= 0) prod[j] = MulApply<XT, ZT>::f(mul, av, uu);
        else prod[j] = p.flip ? MulApply<XT, ZT>::f(mul, uu, av) : MulApply<XT, ZT>::f(mul, av, uu);
    }
    sync();                                                               // head marks are complete
    if (SPMV_PHASE_TIMERS && p.dbg && tid == 0) t2 = clock64();

    // ---- (3a) fold this thread's items between row marks
    int32_t h[IT];
#pragma unroll
    for (int j = 0; j < IT; j += 4) {
        const int4 t4 = *reinterpret_cast<const int4 *>(&s_head[loc0 + j]);
        h[j] = t4.x; h[j + 1] = t4.y; h[j + 2] = t4.z; h[j + 3] = t4.w;
    }
    Part<ZT> acc{(ZT)0, 0}, lead{(ZT)0, 0};
    bool seen = false; int32_t cur = -1;
    if (!SPARSE && nvalid == IT) {
        // every item contributes: presence is structural, no per-item flags
#pragma unroll
        for (int j = 0; j < IT; ++j) {
            if (h[j] >= 0) {
                if (!seen) lead = acc;
                else { tval[r0 + cur] = acc.v; p.tpres[r0 + cur] = 1; }      // row began and ended in this thread
                seen = true; cur = h[j]; acc.v = prod[j]; acc.has = 1;
            } else if (j == 0) { acc.v = prod[0]; acc.has = 1; }
            else acc.v = MulApply<ZT, ZT>::f(add, acc.v, prod[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < IT; ++j) {
            if (h[j] >= 0) {
                if (!seen) lead = acc;
                else { tval[r0 + cur] = acc.v; p.tpres[r0 + cur] = (uint8_t)acc.has; }
                seen = true; cur = h[j]; acc.has = 0;
            }
            const Part<ZT> it{prod[j], (j < nvalid && (!SPARSE || hs[SPARSE ? j : 0])) ? 1 : 0};
            acc = part_join<ZT>(add, acc, it);
        }
    }
    if (!seen) { lead = acc; acc.has = 0; }                               // no mark: everything continues an earlier row

    // ---- (3b) segmented suffix scan of the leads: X_t = lead_t (+) (stop_t ? nothing : X_{t+1}); stop = thread has a mark
    Part<ZT> x = lead; int stop = seen ? 1 : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Part<ZT> y; y.v = shfl_down_t<ZT>(x.v, d); y.has = __shfl_down_sync(0xffffffffu, x.has, d);
        const int ystop = __shfl_down_sync(0xffffffffu, stop, d);
        if (lane + d < 32) { if (!stop) x = part_join<ZT>(add, x, y); stop |= ystop; }
    }
    if (lane == 0) { s_wv[warp] = x.v; s_wflag[warp] = x.has | (stop << 1); }
    sync();
    if (SPMV_PHASE_TIMERS && p.dbg && tid == 0) t3 = clock64();
    Part<ZT> carry{(ZT)0, 0}; int carry_stop = 0;                         // what the following warps add to a row open at this warp's end
    for (int w = SPMV_WARPS - 1; w > warp; --w) {
        Part<ZT> y; y.v = s_wv[w]; const int f = s_wflag[w]; y.has = f & 1;
        if (f >> 1) { carry = y; carry_stop = 1; } else carry = part_join<ZT>(add, y, carry);
    }
    if (!stop) { x = part_join<ZT>(add, x, carry); stop |= carry_stop; }
    Part<ZT> nxt; nxt.v = shfl_down_t<ZT>(x.v, 1); nxt.has = __shfl_down_sync(0xffffffffu, x.has, 1);   // S_t = X_{t+1}
    int nxt_stop = __shfl_down_sync(0xffffffffu, stop, 1);
    if (lane == 31) { nxt = carry; nxt_stop = carry_stop; }

    // ---- (4) rows still open at the end of a thread, and what sticks out of the tile
    if (seen) {
        const Part<ZT> total = part_join<ZT>(add, acc, nxt);
        if (nxt_stop) { tval[r0 + cur] = total.v; p.tpres[r0 + cur] = (uint8_t)total.has; }
        else { static_cast<ZT *>(p.tail_val)[tile] = total.v; p.tail_has[tile] = (uint8_t)total.has; p.tail_row[tile] = (int32_t)(r0 + cur); }
    }
    if (tid == 0 && h[0] < 0) {                                           // the tile starts inside a row of an earlier tile
        static_cast<ZT *>(p.head_val)[tile] = x.v; p.head_has[tile] = (uint8_t)x.has;
    }
    if (SPMV_PHASE_TIMERS && p.dbg && tid == 0) {
        const long long t4 = clock64();
        atomicAdd(&p.dbg[0], (unsigned long long)(t1 - t0)); atomicAdd(&p.dbg[1], (unsigned long long)(t2 - t1));
        atomicAdd(&p.dbg[2], (unsigned long long)(t3 - t2)); atomicAdd(&p.dbg[3], (unsigned long long)(t4 - t3));
        atomicAdd(&p.dbg[4], 1ull);
    }
}

// General path: one CTA per tile, loads straight from global memory into registers.
template <typename XT, typename ZT, int ADD, int MUL, bool SPARSE, int IT>
__global__ void __launch_bounds__(SPMV_THREADS) spmv_tile_kernel(const SpmvArgs p) {
    constexpr int TILE = SPMV_THREADS * IT;
    constexpr bool NEED_A = MUL < 0 || mul_reads_x(MUL);
    __shared__ __align__(16) int32_t s_head[TILE];     // row (relative to the tile's first row) starting at this entry, or -1
    __shared__ ZT s_wv[SPMV_WARPS];
    __shared__ int s_wflag[SPMV_WARPS];
    const int tid = threadIdx.x;
    const uint32_t tile = blockIdx.x;
    const int64_t tstart = (int64_t)tile * TILE;
    const int tlen = (int)min((int64_t)TILE, p.nnz - tstart);            // valid entries in this tile
    const int loc0 = tid * IT;
    const int nvalid = min(max(tlen - loc0, 0), IT);                     // valid entries of this thread
    const uint32_t *colp = p.col + tstart + loc0;
    const XT *avalp = static_cast<const XT *>(p.aval) + tstart + loc0;
    uint32_t c[IT]; XT a[IT];
    if (nvalid == IT) {
#pragma unroll
        for (int g = 0; g < IT / 4; ++g) {
            load4<uint32_t>(colp + g * 4, &c[g * 4]);
            if (NEED_A) load4<XT>(avalp + g * 4, &a[g * 4]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < IT; ++j) {
            c[j] = j < nvalid ? colp[j] : 0u;                            // 0 is always a valid column to gather
            if (NEED_A) a[j] = j < nvalid ? avalp[j] : (XT)1;
        }
    }
    spmv_tile_finish<XT, ZT, ADD, MUL, SPARSE, IT, false>(p, tile, tid, tlen, nvalid, c, a, s_head, s_wv, s_wflag, nullptr, 0u, [] { __syncthreads(); });
}

// Hot-column variant (dense u, specialised semirings, large matrices).  A scattered 4-byte gather
// costs one L1 wavefront per lane, which bounds the general kernel near nnz / (SMs x clock); gathers
// served from shared memory cost a few bank-conflict cycles per warp instead.  The matrix's columns are
// relabelled once by descending in-degree (cached plan), so the hot_n most referenced entries of the
// permuted u form a dense table; persistent CTAs of GROUPS x 256 threads load it once and stride over
// the tiles (on R-MAT graphs the top 32 K of 4 M columns take 53 % of all gathers).
template <typename ZT> __host__ __device__ constexpr size_t hot_group_bytes() { return SPMV_THREADS * 8 * 4 + SPMV_WARPS * 16; }
template <typename XT, typename ZT, int ADD, int MUL, int GROUPS>
__global__ void __launch_bounds__(SPMV_THREADS * GROUPS, (GROUPS == 4 ? 1 : (GROUPS == 3 ? 2 : (GROUPS == 2 ? 3 : 6)))) spmv_hot_kernel(const SpmvArgs p, const uint32_t hot_n) {
    constexpr int IT = 8;
    constexpr int TILE = SPMV_THREADS * IT;
    constexpr bool NEED_A = mul_reads_x(MUL);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int group = threadIdx.x / SPMV_THREADS, tid = threadIdx.x % SPMV_THREADS;
    unsigned char *gbase = smem_raw + (size_t)group * hot_group_bytes<ZT>();
    int32_t *s_head = reinterpret_cast<int32_t *>(gbase);
    ZT *s_wv = reinterpret_cast<ZT *>(gbase + TILE * 4);
    int *s_wflag = reinterpret_cast<int *>(gbase + TILE * 4 + SPMV_WARPS * 8);
    XT *s_hot = reinterpret_cast<XT *>(smem_raw + GROUPS * hot_group_bytes<ZT>());
    const XT *uval = static_cast<const XT *>(p.uval);
    for (uint32_t i = threadIdx.x; i < hot_n; i += blockDim.x) s_hot[i] = uval[i];
    __syncthreads();
    auto sync = [group] { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(SPMV_THREADS) : "memory"); };
    for (int64_t tile = (int64_t)blockIdx.x * GROUPS + group; tile < p.ntiles; tile += (int64_t)gridDim.x * GROUPS) {
        const int64_t tstart = tile * TILE;
        const int tlen = (int)min((int64_t)TILE, p.nnz - tstart);
        const int loc0 = tid * IT;
        const int nvalid = min(max(tlen - loc0, 0), IT);
        const uint32_t *colp = p.col + tstart + loc0;
        const XT *avalp = static_cast<const XT *>(p.aval) + tstart + loc0;
        uint32_t c[IT]; XT a[IT];
        if (nvalid == IT) {
#pragma unroll
            for (int g = 0; g < IT / 4; ++g) {
                load4<uint32_t>(colp + g * 4, &c[g * 4]);
                if (NEED_A) load4<XT>(avalp + g * 4, &a[g * 4]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < IT; ++j) {
                c[j] = j < nvalid ? colp[j] : 0u;
                if (NEED_A) a[j] = j < nvalid ? avalp[j] : (XT)1;
            }
        }
        spmv_tile_finish<XT, ZT, ADD, MUL, false, IT, true>(p, (uint32_t)tile, tid, tlen, nvalid, c, a, s_head, s_wv, s_wflag, s_hot, hot_n, sync);
        sync();                                                               // the group's shared memory is reused by its next tile
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


// ==================================================================================================
// Dense-u kernel for the specialised semirings: warp-independent RUNS.
//
// The entries are cut into runs of 256 (one warp, 8 consecutive entries per lane).  A cached plan gives
// every lane what the tile kernel has to discover with shared-memory marks, a row pass and barriers:
//   run_headw   one bit per entry: "this entry starts a row"
//   run_base    number of row starts before the run;  run_lane: row starts inside the run before the lane
//   nzrow       ids of the non-empty rows (row start rank -> row id);  pres_tmpl: the output's presence
// so a warp needs no shared memory and no barrier: stream 8 entries, gather, fold between the row-start
// bits (rows inside a lane are final), one segmented suffix scan over the 32 lanes, and the two partial
// rows sticking out of the run go to per-run slots that spmv_run_fixup_kernel combines in a fixed order
// (deterministic).  ~3.5x fewer instructions per entry than the tile kernel.
// HOT: persistent CTAs, hot_n most referenced entries of the (relabelled) u in a shared-memory table.
static constexpr int RUN = 256;

struct RunArgs {
    const uint32_t *col; const void *aval; const void *uval;
    const uint32_t *headw; const uint16_t *lane_rank; const uint32_t *run_base; const uint32_t *nzrow; const uint32_t *rowptr;
    const int32_t *tail_row; const uint32_t *tail_last;       // structural: which row is open at a run's end, how far it reaches
    int64_t nruns; int64_t nnz;
    void *tval;
    void *head_val; void *tail_val;                            // per run: partial of the row it starts inside / of the row open at its end
    int add_op, mul_op, flip;                                  // run-time operator codes (kernels instantiated with ADD = MUL = -1)
    const uint8_t *upres;                                      // SPARSE kernels: presence bytes of u ...
    uint8_t *tpres; uint8_t *head_has; uint8_t *tail_has;      // ... and of everything they produce
};

template <typename XT, typename ZT, int ADD_C, int MUL_C, bool HOT, bool SPARSE = false>
__device__ __forceinline__ void spmv_run_body(const RunArgs &p, const int64_t run, const int lane, const XT *s_hot, const uint32_t hot_n) {
    // ADD_C / MUL_C >= 0: compile-time semiring; -1: run-time operator codes (both operands are read)
    constexpr bool NEED_A = MUL_C < 0 || mul_reads_x(MUL_C);
    constexpr bool NEED_U = MUL_C < 0 || mul_reads_y(MUL_C);
    const int ADD = ADD_C >= 0 ? ADD_C : p.add_op;
    const int MUL = MUL_C >= 0 ? MUL_C : p.mul_op;
    const int64_t q = run * RUN + lane * 8;
    const int nvalid = (int)min((int64_t)8, max((int64_t)0, p.nnz - q));
    const XT *uval = static_cast<const XT *>(p.uval);
    ZT *tval = static_cast<ZT *>(p.tval);
    uint32_t c[8]; XT a[8];
    if (nvalid == 8) {
        load4<uint32_t>(p.col + q, &c[0]); load4<uint32_t>(p.col + q + 4, &c[4]);
        if (NEED_A) { load4<XT>(static_cast<const XT *>(p.aval) + q, &a[0]); load4<XT>(static_cast<const XT *>(p.aval) + q + 4, &a[4]); }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = j < nvalid ? p.col[q + j] : 0u;
            if (NEED_A) a[j] = j < nvalid ? static_cast<const XT *>(p.aval)[q + j] : (XT)1;
        }
    }
    const uint32_t hw = nvalid > 0 ? __ldg(p.headw + (q >> 5)) : 0u;
    const uint32_t hb = (hw >> ((lane & 3) * 8)) & 0xffu;                 // this lane's 8 row-start bits
    uint32_t rank = __ldg(p.run_base + run) + __ldg(p.lane_rank + run * 32 + lane);   // row starts before this lane's first entry
    XT uv[8]; uint8_t up[8];
    if (SPARSE) {
        // u has holes: a product exists only where u(col) does; the values are fetched only for those
#pragma unroll
        for (int j = 0; j < 8; ++j) up[j] = j < nvalid ? __ldg(p.upres + c[j]) : (uint8_t)0;
#pragma unroll
        for (int j = 0; j < 8; ++j) uv[j] = (NEED_U && up[j]) ? gload<XT>(uval + c[j]) : (XT)0;
    } else if (NEED_U) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (HOT && c[j] < hot_n) uv[j] = s_hot[c[j]];
            else uv[j] = gload<XT>(uval + c[j]);
        }
    }
    ZT prod[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const XT av = NEED_A ? a[j] : (XT)1, uu = NEED_U ? uv[j] : (XT)1;
        prod[j] = (MUL_C < 0 && p.flip) ? MulApply<XT, ZT>::f(MUL, uu, av) : MulApply<XT, ZT>::f(MUL, av, uu);
    }

    // ---- fold between row starts
    Part<ZT> acc{(ZT)0, 0}, lead{(ZT)0, 0};
    bool seen = false; uint32_t cur = 0;
    if (nvalid == 8 && !SPARSE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if ((hb >> j) & 1u) {
                if (!seen) lead = acc;
                else tval[__ldg(p.nzrow + cur)] = acc.v;                    // row began and ended inside this lane
                seen = true; cur = rank++; acc.v = prod[j]; acc.has = 1;
            } else if (j == 0) { acc.v = prod[0]; acc.has = 1; }
            else acc.v = MulApply<ZT, ZT>::f(ADD, acc.v, prod[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nvalid) {
                if ((hb >> j) & 1u) {
                    if (!seen) lead = acc;
                    else { const uint32_t row = __ldg(p.nzrow + cur); tval[row] = acc.v; if (SPARSE) p.tpres[row] = (uint8_t)acc.has; }
                    seen = true; cur = rank++; acc.has = 0;
                }
                const Part<ZT> it{prod[j], SPARSE ? (int)up[j] : 1};
                acc = part_join<ZT>(ADD, acc, it);
            }
        }
    }
    if (!seen) { lead = acc; acc.has = 0; }

    // ---- segmented suffix scan of the leads over the 32 lanes
    Part<ZT> x = lead; int stop = seen ? 1 : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Part<ZT> y; y.v = shfl_down_t<ZT>(x.v, d);
        const int yf = __shfl_down_sync(0xffffffffu, x.has | (stop << 1), d);
        y.has = yf & 1;
        if (lane + d < 32) { if (!stop) x = part_join<ZT>(ADD, x, y); stop |= yf >> 1; }
    }
    Part<ZT> nxt; nxt.v = shfl_down_t<ZT>(x.v, 1);
    const int nf = __shfl_down_sync(0xffffffffu, x.has | (stop << 1), 1);
    nxt.has = nf & 1; int nxt_stop = nf >> 1;
    if (lane == 31) { nxt.has = 0; nxt_stop = 0; }

    // the lane holding the last row start of the run owns the row that is still open at the run's end
    // (which row that is, and where it ends, is structural: run_tail_row / run_tail_last of the plan)
    if (seen) {
        const Part<ZT> total = part_join<ZT>(ADD, acc, nxt);
        if (nxt_stop) { const uint32_t row = __ldg(p.nzrow + cur); tval[row] = total.v; if (SPARSE) p.tpres[row] = (uint8_t)total.has; }
        else { static_cast<ZT *>(p.tail_val)[run] = total.v; if (SPARSE) p.tail_has[run] = (uint8_t)total.has; }
    }
    if (lane == 0 && !(hb & 1u) && nvalid > 0) {                                            // the run starts inside a row of an earlier run
        static_cast<ZT *>(p.head_val)[run] = x.v; if (SPARSE) p.head_has[run] = (uint8_t)x.has;
    }
}

template <typename XT, typename ZT, int ADD, int MUL, bool SPARSE>
__global__ void __launch_bounds__(256) spmv_run_kernel(const RunArgs p) {
    const int64_t run = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (run >= p.nruns) return;
    spmv_run_body<XT, ZT, ADD, MUL, false, SPARSE>(p, run, threadIdx.x & 31, nullptr, 0u);
}

template <typename XT, typename ZT, int ADD, int MUL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) spmv_run_hot_kernel(const RunArgs p, const uint32_t hot_n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XT *s_hot = reinterpret_cast<XT *>(smem_raw);
    const XT *uval = static_cast<const XT *>(p.uval);
    for (uint32_t i = threadIdx.x; i < hot_n; i += blockDim.x) s_hot[i] = uval[i];
    __syncthreads();
    const int warps = blockDim.x >> 5;
    for (int64_t run = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5); run < p.nruns; run += (int64_t)gridDim.x * warps)
        spmv_run_body<XT, ZT, ADD, MUL, true>(p, run, threadIdx.x & 31, s_hot, hot_n);
}

// rows that continue past their run: tail partial (+) head partials of the following runs, 8 lanes per open row.
// Every run after `run` up to tail_last starts inside that row, so its head partial exists.
template <typename ZT, int ADD_C, bool SPARSE>
__global__ void __launch_bounds__(256) spmv_run_fixup_kernel(const RunArgs p) {
    const int ADD = ADD_C >= 0 ? ADD_C : p.add_op;
    const int sub = threadIdx.x & 7;
    const int64_t run = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int32_t r = run < p.nruns ? __ldg(p.tail_row + run) : -1;
    Part<ZT> acc{(ZT)0, 0};
    if (r >= 0) {
        const int64_t last_run = __ldg(p.tail_last + run);
        if (sub == 0) { acc.v = static_cast<const ZT *>(p.tail_val)[run]; acc.has = SPARSE ? (int)p.tail_has[run] : 1; }
        for (int64_t t = run + 1 + sub; t <= last_run; t += 8) {
            const Part<ZT> y{static_cast<const ZT *>(p.head_val)[t], SPARSE ? (int)p.head_has[t] : 1};
            acc = part_join<ZT>(ADD, acc, y);
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        Part<ZT> y; y.v = shfl_xor_t<ZT>(acc.v, o); y.has = __shfl_xor_sync(0xffffffffu, acc.has, o);
        acc = part_join<ZT>(ADD, acc, y);
    }
    if (r >= 0 && sub == 0) { static_cast<ZT *>(p.tval)[r] = acc.v; if (SPARSE) p.tpres[r] = (uint8_t)acc.has; }
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

static GrB_Info spmv_run_plan(Csr &c, std::string *err) {
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

template <typename XT, typename ZT, int ADD, int MUL>
static void spmv_run_launch(const RunArgs &a, size_t hot_bytes, int64_t hused) {
    if (hot_bytes && ADD >= 0 && !a.upres) {
        // two shapes: one 1024-thread CTA per SM with a table of up to ~200 KB, or two 768-thread CTAs
        // per SM (<= 42 registers) with a table of up to ~100 KB each
        const bool two = hot_bytes <= ((size_t)104 << 10) && getenv("B200GRB_HOT_ONE") == nullptr;
        const uint32_t hot_n = (uint32_t)std::min<int64_t>(hused, (int64_t)(hot_bytes / sizeof(XT)));
        const size_t smem = (size_t)hot_n * sizeof(XT);
        if (two) {
            auto kernel = spmv_run_hot_kernel<XT, ZT, ADD, MUL, 768, 2>;
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kernel<<<G.num_sms * 2, 768, smem, G.stream>>>(a, hot_n); GB_LAUNCHED();
        } else {
            auto kernel = spmv_run_hot_kernel<XT, ZT, ADD, MUL, 1024, 1>;
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kernel<<<G.num_sms, 1024, smem, G.stream>>>(a, hot_n); GB_LAUNCHED();
        }
    } else if (a.upres) {
        spmv_run_kernel<XT, ZT, ADD, MUL, true><<<(unsigned)ceil_div(a.nruns, 8), 256, 0, G.stream>>>(a); GB_LAUNCHED();
        spmv_run_fixup_kernel<ZT, ADD, true><<<(unsigned)ceil_div(a.nruns * 8, 256), 256, 0, G.stream>>>(a); GB_LAUNCHED();
        return;
    } else {
        spmv_run_kernel<XT, ZT, ADD, MUL, false><<<(unsigned)ceil_div(a.nruns, 8), 256, 0, G.stream>>>(a); GB_LAUNCHED();
    }
    spmv_run_fixup_kernel<ZT, ADD, false><<<(unsigned)ceil_div(a.nruns * 8, 256), 256, 0, G.stream>>>(a); GB_LAUNCHED();
}

template <typename T> static bool spmv_run_fast(int add, int mul, const RunArgs &a, size_t hot_bytes, int64_t hused) {
#define GB_FAST(A, M) if (add == A && mul == M) { spmv_run_launch<T, T, A, M>(a, hot_bytes, hused); return true; }
    GB_FAST(OP_PLUS, OP_TIMES) GB_FAST(OP_MIN, OP_PLUS) GB_FAST(OP_PLUS, OP_SECOND) GB_FAST(OP_PLUS, OP_FIRST)
    GB_FAST(OP_PLUS, OP_PAIR) GB_FAST(OP_MIN, OP_FIRST) GB_FAST(OP_MIN, OP_SECOND)
#undef GB_FAST
    return false;
}
static bool spmv_run_fast_bool(int add, int mul, const RunArgs &a, size_t hot_bytes, int64_t hused) {
#define GB_FAST(A, M) if (add == A && mul == M) { spmv_run_launch<bool, bool, A, M>(a, hot_bytes, hused); return true; }
    GB_FAST(OP_LOR, OP_LAND) GB_FAST(OP_ANY, OP_PAIR) GB_FAST(OP_LOR, OP_PAIR) GB_FAST(OP_LOR, OP_SECOND) GB_FAST(OP_LOR, OP_FIRST)
#undef GB_FAST
    return false;
}
static bool spmv_run_generic(int xt, int zt, const RunArgs &a) {
#define GB_RUNGEN(XT_, ZT_) do { spmv_run_launch<XT_, ZT_, -1, -1>(a, 0, 0); return true; } while (0)
    if (xt == zt) {
        switch (xt) {
#define GB_GEN(TC, T) case TC: GB_RUNGEN(T, T);
            GB_GEN(TC_BOOL, bool) GB_GEN(TC_INT8, int8_t) GB_GEN(TC_INT16, int16_t) GB_GEN(TC_INT32, int32_t) GB_GEN(TC_INT64, int64_t)
            GB_GEN(TC_UINT8, uint8_t) GB_GEN(TC_UINT16, uint16_t) GB_GEN(TC_UINT32, uint32_t) GB_GEN(TC_UINT64, uint64_t)
            GB_GEN(TC_FP32, float) GB_GEN(TC_FP64, double)
#undef GB_GEN
        }
    } else if (zt == TC_BOOL) {
        switch (xt) {
#define GB_GEN(TC, T) case TC: GB_RUNGEN(T, bool);
            GB_GEN(TC_INT8, int8_t) GB_GEN(TC_INT16, int16_t) GB_GEN(TC_INT32, int32_t) GB_GEN(TC_INT64, int64_t)
            GB_GEN(TC_UINT8, uint8_t) GB_GEN(TC_UINT16, uint16_t) GB_GEN(TC_UINT32, uint32_t) GB_GEN(TC_UINT64, uint64_t)
            GB_GEN(TC_FP32, float) GB_GEN(TC_FP64, double)
#undef GB_GEN
        }
    }
#undef GB_RUNGEN
    return false;
}
static bool spmv_run_dispatch(int xt, int add, int mul, const RunArgs &a, size_t hot_bytes, int64_t hused) {
    switch (xt) {
        case TC_FP32: return spmv_run_fast<float>(add, mul, a, hot_bytes, hused);
        case TC_FP64: return spmv_run_fast<double>(add, mul, a, hot_bytes, hused);
        case TC_INT32: return spmv_run_fast<int32_t>(add, mul, a, hot_bytes, hused);
        case TC_INT64: return spmv_run_fast<int64_t>(add, mul, a, hot_bytes, hused);
        case TC_UINT32: return spmv_run_fast<uint32_t>(add, mul, a, hot_bytes, hused);
        case TC_UINT64: return spmv_run_fast<uint64_t>(add, mul, a, hot_bytes, hused);
        case TC_BOOL: return spmv_run_fast_bool(add, mul, a, hot_bytes, hused);
        default: return false;
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
    const int64_t re = p.rowptr[r + 1];
    const int64_t last_tile = (re - 1) / p.tile;
    Part<ZT> acc{(ZT)0, 0};
    if (lane == 0 && p.tail_has[t]) { acc.v = static_cast<const ZT *>(p.tail_val)[t]; acc.has = 1; }
    for (int64_t tt = t + 1 + lane; tt <= last_tile; tt += 32) if (p.head_has[tt]) {
        const Part<ZT> y{static_cast<const ZT *>(p.head_val)[tt], 1};
        acc = part_join<ZT>(add, acc, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Part<ZT> y; y.v = shfl_xor_t<ZT>(acc.v, o); y.has = __shfl_xor_sync(0xffffffffu, acc.has, o);
        acc = part_join<ZT>(add, acc, y);
    }
    if (lane == 0) { static_cast<ZT *>(p.tval)[r] = acc.v; p.tpres[r] = (uint8_t)(acc.has != 0); }
}

__global__ void clear_presence_kernel(uint8_t *p, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) p[k] = 0;
}

static int g_items_fast = 8, g_items_generic = 8;     // entries per thread (tunable: B200GRB_SPMV_ITEMS)

struct HotLaunch { bool on; int64_t hused; int groups; size_t table_bytes; };
static HotLaunch g_hot{false, 0, 4, (size_t)128 << 10};

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
static inline int hgrid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (int64_t)G.num_sms * 16)); }

static GrB_Info spmv_hot_plan(Csr &c, std::string *err) {
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

template <typename XT, typename ZT, int ADD, int MUL, bool SPARSE, int IT>
static void spmv_launch(const SpmvArgs &a) {
    if constexpr (ADD >= 0 && !SPARSE && IT == 8) {
        if (g_hot.on) {
            if (g_hot.groups == 1) spmv_hot_launch<XT, ZT, ADD, MUL, 1>(a, g_hot);
            else if (g_hot.groups == 2) spmv_hot_launch<XT, ZT, ADD, MUL, 2>(a, g_hot);
            else if (g_hot.groups == 3) spmv_hot_launch<XT, ZT, ADD, MUL, 3>(a, g_hot);
            else spmv_hot_launch<XT, ZT, ADD, MUL, 4>(a, g_hot);
            spmv_fixup_kernel<ZT, ADD><<<(unsigned)ceil_div(a.ntiles * 32, 256), 256, 0, G.stream>>>(a); GB_LAUNCHED();
            return;
        }
    }
    spmv_tile_kernel<XT, ZT, ADD, MUL, SPARSE, IT><<<(unsigned)a.ntiles, SPMV_THREADS, 0, G.stream>>>(a); GB_LAUNCHED();
    spmv_fixup_kernel<ZT, ADD><<<(unsigned)ceil_div(a.ntiles * 32, 256), 256, 0, G.stream>>>(a); GB_LAUNCHED();
}

// compile-time specialised semirings (BASELINE.json north_star: PLUS_TIMES, LOR_LAND, MIN_PLUS,
// PLUS_SECOND; plus PLUS_PAIR / ANY_PAIR / PLUS_FIRST / MIN_FIRST / MIN_SECOND which the reference's
// demos use) for dense u; everything else runs the same kernel with run-time operator codes.
template <typename T> static bool spmv_fast(int add, int mul, int items, const SpmvArgs &a) {
#define GB_FAST(A, M) if (add == A && mul == M) { if (items == 16) spmv_launch<T, T, A, M, false, 16>(a); else if (items == 4) spmv_launch<T, T, A, M, false, 4>(a); else spmv_launch<T, T, A, M, false, 8>(a); return true; }
    GB_FAST(OP_PLUS, OP_TIMES) GB_FAST(OP_MIN, OP_PLUS) GB_FAST(OP_PLUS, OP_SECOND) GB_FAST(OP_PLUS, OP_FIRST)
    GB_FAST(OP_PLUS, OP_PAIR) GB_FAST(OP_MIN, OP_FIRST) GB_FAST(OP_MIN, OP_SECOND)
#undef GB_FAST
    return false;
}
static bool spmv_fast_bool(int add, int mul, int items, const SpmvArgs &a) {
#define GB_FAST(A, M) if (add == A && mul == M) { if (items == 16) spmv_launch<bool, bool, A, M, false, 16>(a); else spmv_launch<bool, bool, A, M, false, 8>(a); return true; }
    GB_FAST(OP_LOR, OP_LAND) GB_FAST(OP_ANY, OP_PAIR) GB_FAST(OP_LOR, OP_PAIR) GB_FAST(OP_LOR, OP_SECOND) GB_FAST(OP_LOR, OP_FIRST)
#undef GB_FAST
    return false;
}

// the tile size a call will use (so that the plan can be built first)
static bool spmv_is_fast(int xt, int zt, int add, int mul, bool sparse_u) {
    if (sparse_u || xt != zt) return false;
    const bool num = (add == OP_PLUS && (mul == OP_TIMES || mul == OP_SECOND || mul == OP_FIRST || mul == OP_PAIR)) ||
                     (add == OP_MIN && (mul == OP_PLUS || mul == OP_FIRST || mul == OP_SECOND));
    const bool boo = (add == OP_LOR && (mul == OP_LAND || mul == OP_PAIR || mul == OP_SECOND || mul == OP_FIRST)) || (add == OP_ANY && mul == OP_PAIR);
    switch (xt) {
        case TC_FP32: case TC_FP64: case TC_INT32: case TC_INT64: case TC_UINT32: case TC_UINT64: return num;
        case TC_BOOL: return boo;
        default: return false;
    }
}

static GrB_Info spmv_dispatch(int xt, int zt, int add, int mul, bool sparse_u, const SpmvArgs &a, std::string *err) {
    if (spmv_is_fast(xt, zt, add, mul, sparse_u)) {
        const int items = a.tile / SPMV_THREADS;
        switch (xt) {
            case TC_FP32:  if (spmv_fast<float>(add, mul, items, a)) return GrB_SUCCESS; break;
            case TC_FP64:  if (spmv_fast<double>(add, mul, items, a)) return GrB_SUCCESS; break;
            case TC_INT32: if (spmv_fast<int32_t>(add, mul, items, a)) return GrB_SUCCESS; break;
            case TC_INT64: if (spmv_fast<int64_t>(add, mul, items, a)) return GrB_SUCCESS; break;
            case TC_UINT32: if (spmv_fast<uint32_t>(add, mul, items, a)) return GrB_SUCCESS; break;
            case TC_UINT64: if (spmv_fast<uint64_t>(add, mul, items, a)) return GrB_SUCCESS; break;
            case TC_BOOL:  if (spmv_fast_bool(add, mul, items, a)) return GrB_SUCCESS; break;
            default: break;
        }
        return gb_fail(GrB_PANIC, err, "mxv: internal dispatch error");
    }
#define GB_GEN2(XT_, ZT_) do { if (sparse_u) spmv_launch<XT_, ZT_, -1, -1, true, 8>(a); else spmv_launch<XT_, ZT_, -1, -1, false, 8>(a); return GrB_SUCCESS; } while (0)
    if (xt == zt) {
        switch (xt) {
#define GB_GEN(TC, T) case TC: GB_GEN2(T, T);
            GB_GEN(TC_BOOL, bool) GB_GEN(TC_INT8, int8_t) GB_GEN(TC_INT16, int16_t) GB_GEN(TC_INT32, int32_t) GB_GEN(TC_INT64, int64_t)
            GB_GEN(TC_UINT8, uint8_t) GB_GEN(TC_UINT16, uint16_t) GB_GEN(TC_UINT32, uint32_t) GB_GEN(TC_UINT64, uint64_t)
            GB_GEN(TC_FP32, float) GB_GEN(TC_FP64, double)
#undef GB_GEN
        }
    } else if (zt == TC_BOOL) {
        switch (xt) {
#define GB_GEN(TC, T) case TC: GB_GEN2(T, bool);
            GB_GEN(TC_INT8, int8_t) GB_GEN(TC_INT16, int16_t) GB_GEN(TC_INT32, int32_t) GB_GEN(TC_INT64, int64_t)
            GB_GEN(TC_UINT8, uint8_t) GB_GEN(TC_UINT16, uint16_t) GB_GEN(TC_UINT32, uint32_t) GB_GEN(TC_UINT64, uint64_t)
            GB_GEN(TC_FP32, float) GB_GEN(TC_FP64, double)
#undef GB_GEN
        }
    }
#undef GB_GEN2
    return gb_fail(GrB_DOMAIN_MISMATCH, err, "mxv: unsupported semiring domains (x=%d, z=%d)", xt, zt);
}

// ------------------------------------------------------------------ masked pull with early exit (BFS-shaped calls)
// w<mask> = A (+).(x) u for monoids with a terminal value (LOR, LAND, ANY): one warp per row, rows the mask
// rules out are skipped entirely (their entries are never read), and a row stops as soon as its monoid
// saturates -- the BFS step `A.mxv(q, mask=visited, desc=RC, semiring=LOR_LAND)` of
// /root/reference/tests/test_descriptor.py:13-30 touches only the unvisited rows and, for each, only the
// entries up to the first frontier hit.  Output: T restricted to the rows the mask lets through.
struct PullArgs {
    const uint32_t *rowptr; const uint32_t *col; const void *aval; int64_t nrows;
    const void *uval; const uint8_t *upres;
    const void *mval; const uint8_t *mpres; int mtc; int mask_comp, mask_struct;
    void *tval; uint8_t *tpres;
    int add_op, mul_op, flip;
    int has_long; int64_t long_cap;
    uint32_t *long_rows; int *long_count;                      // work list of the rows left to the CTA-per-row kernel
};
template <typename ZT> __device__ __forceinline__ bool monoid_saturated(int add, ZT v) {
    switch (add) {
        case OP_LOR: return v != (ZT)0;
        case OP_LAND: return v == (ZT)0;
        case OP_ANY: return true;
        default: return false;
    }
}
template <typename T> __device__ __forceinline__ T shfl_idx_t(T v, int src) {
    if constexpr (sizeof(T) == 8) { long long x = reinterpret_cast<long long &>(v); x = __shfl_sync(0xffffffffu, x, src); return reinterpret_cast<T &>(x); }
    else if constexpr (sizeof(T) == 4) { int x = reinterpret_cast<int &>(v); x = __shfl_sync(0xffffffffu, x, src); return reinterpret_cast<T &>(x); }
    else { int x = (int)v; x = __shfl_sync(0xffffffffu, x, src); return (T)x; }
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
constexpr uint32_t PULL_LONG = 4096;          // rows longer than this go to the CTA-per-row kernel

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
static GrB_Info spmv_masked_pull_dispatch(int xt, int zt, const PullArgs &a, std::string *err) {
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
        if (a.opres) a.opres[i] = op;
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
    // the multiply sees (x = matrix entry, y = u entry) for mxv and the reverse for vxm; for the
    // operators that ignore one side the flip folds into the operator itself
    int kmul = mul; bool kflip = flip;
    if (flip && (mul == OP_FIRST || mul == OP_SECOND)) { kmul = mul == OP_FIRST ? OP_SECOND : OP_FIRST; kflip = false; }
    if (flip && (mul == OP_TIMES || mul == OP_PLUS || mul == OP_MIN || mul == OP_MAX || mul == OP_PAIR || mul == OP_LAND || mul == OP_LOR ||
                 mul == OP_LXOR || mul == OP_EQ || mul == OP_NE || mul == OP_ISEQ || mul == OP_ISNE)) kflip = false;     // commutative
    const bool need_a = kflip ? op_uses_y(kmul) : op_uses_x(kmul);
    const bool need_u = kflip ? op_uses_x(kmul) : op_uses_y(kmul);
    const bool sparse_u = u->dpres != nullptr;
    const bool fast_sr = !kflip && spmv_is_fast(xt, zt, add, kmul, false);      // a compile-time specialised semiring
    const bool fast = fast_sr && !sparse_u;
    { const char *e = getenv("B200GRB_SPMV_ITEMS"); const int v = e ? atoi(e) : 8; g_items_fast = (v == 16 || v == 4) ? v : 8; }
    const int tile = SPMV_THREADS * (fast ? g_items_fast : g_items_generic);
    GB_TRY(spmv_plan(c, tile, err));

    void *a_cast = nullptr, *u_cast = nullptr;
    const void *aval = c.val, *uval = u->dval;
    if (need_a && A->type->code != xt) { GB_TRY(dev_cast_values(&a_cast, xt, c.val, A->type->code, c.nnz, err)); aval = a_cast; }
    if (need_u && u->type->code != xt) { GB_TRY(dev_cast_values(&u_cast, xt, u->dval, u->type->code, (int64_t)u->n, err)); uval = u_cast; }
    // run-time-operator kernels always read both operands: give them something readable of the right type
    if (!fast && !need_a && A->type->code != xt) { GB_TRY(dev_cast_values(&a_cast, xt, c.val, A->type->code, c.nnz, err)); aval = a_cast; }
    if (!fast && !need_u && u->type->code != xt) { GB_TRY(dev_cast_values(&u_cast, xt, u->dval, u->type->code, (int64_t)u->n, err)); uval = u_cast; }

    const int64_t n = (int64_t)out_n;
    const size_t zsz = (size_t)tc_size(zt);
    void *tval = nullptr; uint8_t *tpres = nullptr;
    GB_TRY(dmalloc(&tval, (size_t)n * zsz + 16, err));
    GB_TRY(dmalloc((void **)&tpres, (size_t)n + 16, err));

    // mask + saturating monoid (BFS-shaped): skip masked-out rows, stop rows at the first hit
    const bool use_pull = mask != nullptr && (add == OP_LOR || add == OP_LAND || add == OP_ANY) && c.nnz > 0 && getenv("B200GRB_NO_PULL") == nullptr;
    if (use_pull) {
        // the run-time-operator kernel reads both operands: make sure both are of the operand type
        if (aval == c.val && A->type->code != xt) { GB_TRY(dev_cast_values(&a_cast, xt, c.val, A->type->code, c.nnz, err)); aval = a_cast; }
        if (uval == u->dval && u->type->code != xt) { GB_TRY(dev_cast_values(&u_cast, xt, u->dval, u->type->code, (int64_t)u->n, err)); uval = u_cast; }
        PullArgs pa{};
        pa.rowptr = c.rowptr32; pa.col = c.col; pa.aval = aval; pa.nrows = c.nrows; pa.uval = uval; pa.upres = u->dpres;
        pa.mval = mask->dval; pa.mpres = mask->dpres; pa.mtc = mask->type->code; pa.mask_comp = f.mask_comp; pa.mask_struct = f.mask_struct;
        pa.tval = tval; pa.tpres = tpres; pa.add_op = add; pa.mul_op = kmul; pa.flip = kflip;
        // rows longer than PULL_LONG: at most nnz / PULL_LONG of them
        pa.long_cap = c.nnz / (int64_t)PULL_LONG + 1; pa.has_long = c.nnz > (int64_t)PULL_LONG;
        GB_TRY(dalloc(&pa.long_rows, (size_t)pa.long_cap + 1, err));
        GB_TRY(dalloc(&pa.long_count, 4, err));
        CU_TRY(cudaMemsetAsync(pa.long_count, 0, sizeof(int), G.stream), err);
        GrB_Info r = spmv_masked_pull_dispatch(xt, zt, pa, err);
        dfree(pa.long_rows); dfree(pa.long_count);
        if (r != GrB_SUCCESS) { dfree(tval); dfree(tpres); dfree(a_cast); dfree(u_cast); return r; }
    }
    // dense u + specialised semiring: warp-independent run kernel on the cached run plan
    const bool run_ok = !use_pull && (fast_sr || xt == zt || zt == TC_BOOL);     // specialised or run-time operators; dense or sparse u
    bool use_run = run_ok && c.nnz >= 4096;
    if (const char *e = getenv("B200GRB_SPMV_RUN")) use_run = run_ok && c.nnz > 0 && atoi(e) != 0;
    if (use_pull) {
        // done above
    } else if (use_run) {
        GB_TRY(spmv_run_plan(c, err));
        CU_TRY(cudaMemsetAsync(tval, 0, (size_t)n * zsz, G.stream), err);
        if (sparse_u) CU_TRY(cudaMemsetAsync(tpres, 0, (size_t)n, G.stream), err);     // presence follows u: written row by row
        else CU_TRY(cudaMemcpyAsync(tpres, c.pres_tmpl, (size_t)n, cudaMemcpyDeviceToDevice, G.stream), err);
        RunArgs ra{};
        ra.col = c.col; ra.aval = aval; ra.uval = uval; ra.headw = c.run_headw; ra.lane_rank = c.run_lane; ra.run_base = c.run_base;
        ra.nzrow = c.nzrow; ra.rowptr = c.rowptr32; ra.nruns = c.nruns; ra.nnz = c.nnz; ra.tval = tval;
        ra.tail_row = c.run_tail_row; ra.tail_last = c.run_tail_last;
        ra.add_op = add; ra.mul_op = kmul; ra.flip = kflip;
        GB_TRY(dmalloc(&ra.head_val, (size_t)c.nruns * zsz + 16, err));
        GB_TRY(dmalloc(&ra.tail_val, (size_t)c.nruns * zsz + 16, err));
        if (sparse_u) {
            ra.upres = u->dpres; ra.tpres = tpres;
            GB_TRY(dmalloc((void **)&ra.head_has, (size_t)c.nruns + 16, err));
            GB_TRY(dmalloc((void **)&ra.tail_has, (size_t)c.nruns + 16, err));
        }
        void *u_perm = nullptr; size_t hot_bytes = 0;
        // hot-column table: on by default for large matrices whose gathers are concentrated (R-MAT-like);
        // B200GRB_SPMV_HOT=0 disables it, =<KB> forces a table size
        const char *hot_env = getenv("B200GRB_SPMV_HOT");
        int hot_kb = hot_env ? atoi(hot_env) : 160;
        if (fast && need_u && hot_kb > 0 && c.nnz >= ((int64_t)1 << 20) && c.ncols >= (1 << 16)) {
            GB_TRY(spmv_hot_plan(c, err));
            if (!hot_env && c.hot_cover < 0.25) hot_kb = 0;
        } else hot_kb = 0;
        if (hot_kb > 0) {
            const size_t xsz = (size_t)tc_size(xt);
            GB_TRY(dmalloc(&u_perm, (size_t)c.hused * xsz + 16, err));
            if (c.hused > 0) { permute_u_kernel<<<hgrid(c.hused), 256, 0, G.stream>>>(c.hperm, (const uint8_t *)uval, (uint8_t *)u_perm, (int)xsz, c.hused); GB_LAUNCHED(); }
            ra.col = c.hcol; ra.uval = u_perm; hot_bytes = (size_t)hot_kb << 10;
        }
        const bool ok = fast_sr ? spmv_run_dispatch(xt, add, kmul, ra, hot_bytes, c.hused) : spmv_run_generic(xt, zt, ra);
        dfree(u_perm); dfree(ra.head_val); dfree(ra.tail_val); dfree(ra.head_has); dfree(ra.tail_has);
        if (!ok) { dfree(tval); dfree(tpres); dfree(a_cast); dfree(u_cast); return gb_fail(GrB_PANIC, err, "mxv: internal dispatch error"); }
    } else if (c.nnz == 0) {
        clear_presence_kernel<<<grid_for(n), 256, 0, G.stream>>>(tpres, n); GB_LAUNCHED();
    } else {
        SpmvArgs a{};
        a.rowptr = c.rowptr32; a.col = c.col; a.aval = aval; a.tile_row = c.tile_row; a.ntiles = c.ntiles;
        a.nrows = c.nrows; a.nnz = c.nnz; a.uval = uval; a.upres = u->dpres; a.tval = tval; a.tpres = tpres;
        a.add_op = add; a.mul_op = kmul; a.flip = kflip; a.tile = tile;
        GB_TRY(dmalloc(&a.head_val, (size_t)c.ntiles * zsz + 16, err));
        GB_TRY(dmalloc(&a.tail_val, (size_t)c.ntiles * zsz + 16, err));
        GB_TRY(dmalloc((void **)&a.head_has, (size_t)c.ntiles + 16, err));
        GB_TRY(dmalloc((void **)&a.tail_has, (size_t)c.ntiles + 16, err));
        GB_TRY(dalloc(&a.tail_row, (size_t)c.ntiles, err));
        // dense u on a large matrix with a specialised semiring: hot-column plan + shared-memory table
        void *u_perm = nullptr;
        g_hot.on = false;
        const char *hot_env = getenv("B200GRB_SPMV_HOT");
        if (fast && need_u && tile == SPMV_THREADS * 8 && c.nnz >= ((int64_t)1 << 20) && c.ncols >= (1 << 16) && hot_env && atoi(hot_env) > 0) {
            GB_TRY(spmv_hot_plan(c, err));
            const size_t xsz = (size_t)tc_size(xt);
            GB_TRY(dmalloc(&u_perm, (size_t)c.hused * xsz + 16, err));
            if (c.hused > 0) { permute_u_kernel<<<hgrid(c.hused), 256, 0, G.stream>>>(c.hperm, (const uint8_t *)uval, (uint8_t *)u_perm, (int)xsz, c.hused); GB_LAUNCHED(); }
            a.col = c.hcol; a.uval = u_perm;
            g_hot.on = true; g_hot.hused = c.hused; g_hot.table_bytes = (size_t)atoi(hot_env) << 10;
            if (const char *e = getenv("B200GRB_HOT_GROUPS")) g_hot.groups = atoi(e); else g_hot.groups = 4;
        }
        if (SPMV_PHASE_TIMERS && getenv("B200GRB_SPMV_DEBUG")) { GB_TRY(dalloc(&a.dbg, 8, err)); CU_TRY(cudaMemsetAsync(a.dbg, 0, 64, G.stream), err); }
        GrB_Info r = spmv_dispatch(xt, zt, add, kmul, sparse_u, a, err);
        if (a.dbg) {
            unsigned long long h[5];
            cudaMemcpyAsync(h, a.dbg, 40, cudaMemcpyDeviceToHost, G.stream); cudaStreamSynchronize(G.stream);
            if (h[4]) fprintf(stderr, "[spmv phases, avg cycles per tile] load+issue %.0f | rows+gather %.0f | fold+scan %.0f | carry+store %.0f | tiles %llu\n",
                              (double)h[0] / h[4], (double)h[1] / h[4], (double)h[2] / h[4], (double)h[3] / h[4], h[4]);
            dfree(a.dbg);
        }
        g_hot.on = false;
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
        // a full w stays full under an accumulator (and a mask that does not replace): no presence bytes, and
        // the next sweep sees a dense operand (SSSP: v = min(v, A' min.+ v))
        const bool out_full = !w_empty && w->dpres == nullptr && accum != nullptr && !(mask && f.replace);
        if (!out_full) GB_TRY(dmalloc((void **)&fa.opres, (size_t)n + 16, err));
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
