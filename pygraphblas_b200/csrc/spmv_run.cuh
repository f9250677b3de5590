// The run kernels and their launcher, as templates; instantiated by spmv_run.cu (FP32 / FP64 semirings),
// spmv_run_int.cu (integer / BOOL semirings) and spmv_run_generic.cu (run-time operator codes).
#pragma once
#include "spmv_args.cuh"

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

