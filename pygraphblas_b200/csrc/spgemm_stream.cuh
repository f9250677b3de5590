// spgemm_stream.cuh -- masked SpGEMM, the chunked classes:  T<M> = A (+).(x) B  where products are kept only on M's pattern.
//
// The triangle kernel C<L> = L (+.pair) L of BASELINE.json configs[3] makes 8e9 products for 1e7 outputs: 96 % of the
// products miss the mask.  The kernel is therefore built around the MISS path:
//
//   * membership filter: a BITMAP of the mask row's columns in shared memory (exact, bit j, when ncols fits; a one-hash
//     Bloom filter otherwise) -- one LDS and a bit test per product instead of a hash probe loop;
//   * B rows are streamed by whole warps with coalesced 128-bit loads of four column ids per lane (one A entry per
//     warp at a time; B rows longer than LONG_ROW positions are cut into 128-position blocks dealt to all the warps);
//   * the survivors (true hits, ~5 % of the products, + filter false positives) are compacted with ballots into a
//     per-warp queue and resolved 32 at a time by all lanes: a shared-memory hash of the mask row (short mask rows) or a
//     dense column -> position map in HBM (hub mask rows) gives the output position in one probe, then the product is
//     combined into a shared-memory accumulator (short mask rows) or straight into HBM;
//   * CTAs are persistent and take blocks of CONSECUTIVE chunks (a chunk = row i + a flop-bounded slice of A(i,:)), so
//     the filter of a hub row is built once and reused by all its chunks; rows are changed by clearing exactly the
//     bits that were set.
//
// Algorithmic bytes: flops x (4 + b_B) for the streamed B entries -- they come from the L2 when B fits
// (scale 20: 63 MB of column ids), which is what the coalesced 128-bit loads are for.
#pragma once

struct StreamArgs {
    GemmArgs g;
    const int32_t *chunk_row; const uint32_t *chunk_idx; const uint32_t *chunk_cnt; int64_t nchunks;
    void *t_words;                  // nnz(M) accumulator words (pre-set to the monoid identity)
    unsigned int *queue;            // next block of chunks
    int bm_log2;                    // bitmap bits = 1 << bm_log2
    int exact;                      // 1: bit index = column id (ncols <= bitmap bits); 0: one multiplicative hash
    int vals_cap;                   // mask rows up to this long accumulate in shared memory (0: straight into HBM)
    int table;                      // vals_cap > 0: slots of the shared-memory hash (column -> position in the mask row), 2 x vals_cap
    int32_t *spa_slot;              // vals_cap == 0: per-CTA dense map column -> mask position in HBM (ncols entries each, -1 when idle)
    int grab;                       // consecutive chunks per queue grab
};

constexpr uint32_t STREAM_LONG_ROW = 1024;      // B rows longer than this are spread over all the warps of the CTA
constexpr int STREAM_QCAP = 64;                 // survivor queue entries per warp

__host__ __device__ inline size_t stream_var_smem(int bm_log2, int table, int vals_cap, size_t wsize) {
    const size_t head = ((((size_t)1 << bm_log2) / 8 + (size_t)table * 6 + 7) & ~(size_t)7) + (size_t)vals_cap * (wsize + 1);
    return (head + 15) & ~(size_t)15;
}
template <int NT, typename XT, typename ZT> __host__ __device__ constexpr size_t stream_fixed_smem() {
    // per-batch A entry arrays + survivor queues + long-row list
    return (size_t)NT * (4 + 4 + sizeof(XT)) + (size_t)(NT / 32) * STREAM_QCAP * 8 + (size_t)NT * 2 + 64;
}

template <int NT, typename XT, typename ZT, int ADD, int MUL>
__global__ void __launch_bounds__(NT, NT == 1024 ? 1 : 4) masked_stream_kernel(const StreamArgs sa) {
    typedef typename SlotWord<ZT>::W W;
    constexpr int NW = NT / 32;
    const GemmArgs &p = sa.g;
    const int add = ADD >= 0 ? ADD : p.add_op;
    const int mul = MUL >= 0 ? MUL : p.mul_op;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [bitmap words][hash keys u32 x table][hash slot u16 x table][vals W x cap][found u8 x cap][s_av][s_bs][s_len][queues][long list]
    const uint32_t bm_words = 1u << (sa.bm_log2 - 5);
    uint32_t *bm = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *keys = bm + bm_words;
    uint16_t *slot = reinterpret_cast<uint16_t *>(keys + sa.table);
    W *vals = reinterpret_cast<W *>(smem_raw + (((size_t)bm_words * 4 + (size_t)sa.table * 6 + 7) & ~(size_t)7));
    uint8_t *found = reinterpret_cast<uint8_t *>(vals + sa.vals_cap);
    unsigned char *rest = smem_raw + stream_var_smem(sa.bm_log2, sa.table, sa.vals_cap, sizeof(W));
    const uint32_t tmask = (uint32_t)sa.table - 1u;
    const int hshift = sa.table ? __clz(sa.table) + 1 : 0;
    int32_t *gslot = sa.spa_slot ? sa.spa_slot + (size_t)blockIdx.x * p.ncols : nullptr;
    XT *s_av = reinterpret_cast<XT *>(rest);
    uint32_t *s_bs = reinterpret_cast<uint32_t *>(rest + (size_t)NT * sizeof(XT));
    uint32_t *s_len = s_bs + NT;
    uint2 *s_q = reinterpret_cast<uint2 *>(s_len + NT);
    uint16_t *s_long = reinterpret_cast<uint16_t *>(s_q + NW * STREAM_QCAP);
    __shared__ unsigned int s_next, s_nlong;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int bm_shift = 32 - sa.bm_log2;
    const XT *aval = static_cast<const XT *>(p.a_val), *bval = static_cast<const XT *>(p.b_val);
    W *tw = static_cast<W *>(sa.t_words);
    const W ident = pack_slot<ZT>(monoid_identity<ZT>(add));
    uint2 *myq = s_q + warp * STREAM_QCAP;

    auto bit_of = [&](uint32_t j) -> uint32_t { return sa.exact ? j : ((j * 2654435761u) >> bm_shift); };

    for (uint32_t w = tid; w < bm_words; w += NT) bm[w] = 0u;
    for (int t = tid; t < sa.table; t += NT) keys[t] = EMPTY_KEY;
    int64_t cur_row = -1; uint32_t ms = 0, me = 0; bool local = false; bool split = false;
    __syncthreads();

    // ---- leaving a row: its shared-memory accumulators go to HBM, its bits are cleared
    auto leave_row = [&]() {
        if (cur_row < 0) return;
        const int mlen = (int)(me - ms);
        if (local) {
            if (!split) { for (int q = tid; q < mlen; q += NT) { tw[ms + q] = vals[q]; p.t_found[ms + q] = found[q]; } }
            else for (int q = tid; q < mlen; q += NT) if (found[q]) { atomic_combine<ZT>(&tw[ms + q], unpack_slot<ZT>(vals[q]), add); p.t_found[ms + q] = 1; }
        }
        for (uint32_t q = ms + tid; q < me; q += NT) {
            const uint32_t j = p.m_col[q], b = bit_of(j); bm[b >> 5] = 0u;      // whole words: all their bits are this row's
            if (gslot) gslot[j] = -1;
        }
        if (local) for (int t = tid; t < sa.table; t += NT) keys[t] = EMPTY_KEY;
        __syncthreads();
    };
    auto enter_row = [&](int64_t row, bool row_is_split) {
        cur_row = row; ms = p.m_ptr[row]; me = p.m_ptr[row + 1]; split = row_is_split;
        const int mlen = (int)(me - ms);
        local = mlen <= sa.vals_cap;
        if (local) for (int q = tid; q < mlen; q += NT) { vals[q] = ident; found[q] = 0; }
        for (uint32_t q = ms + tid; q < me; q += NT) {
            const uint32_t j = p.m_col[q], b = bit_of(j);
            atomicOr(&bm[b >> 5], 1u << (b & 31));
            if (local) {                                        // mask columns are unique: plain insertion
                uint32_t h = hash_col(j, hshift);
                while (atomicCAS(&keys[h], EMPTY_KEY, j) != EMPTY_KEY) h = (h + 1) & tmask;
                slot[h] = (uint16_t)(q - ms);
            } else gslot[j] = (int32_t)q;
        }
        __syncthreads();
    };

    // ---- survivors: column j of B position pos passed the filter; find it in the mask row and combine
    auto resolve = [&](uint32_t j, uint32_t pos, XT av) {
        uint32_t lo;
        if (local) {
            uint32_t h = hash_col(j, hshift);
            while (true) {
                const uint32_t kk = keys[h];
                if (kk == j) break;
                if (kk == EMPTY_KEY) return;                   // a false positive of the filter
                h = (h + 1) & tmask;
            }
            lo = ms + slot[h];
        } else {
            const int32_t q = gslot[j];                        // one L2 access per survivor (exact filter: always a hit)
            if (q < 0) return;
            lo = (uint32_t)q;
        }
        if (!p.m_struct && sc_cast(sc_load(p.m_tc, p.m_val, lo), p.m_tc, TC_BOOL).u == 0) return;
        const XT bv = p.need_b ? gload<XT>(bval + pos) : (XT)1;
        const ZT prod = MulApply<XT, ZT>::f(mul, av, bv);
        if (local) { atomic_combine<ZT>(&vals[lo - ms], prod, add); found[lo - ms] = 1; }
        else { atomic_combine<ZT>(&tw[lo], prod, add); p.t_found[lo] = 1; }
    };
    int qn = 0;                                                 // entries in this warp's queue (warp-uniform)
    auto drain = [&](int count, XT av) {                        // the newest `count` entries, one per lane
        __syncwarp();
        if (lane < count) { const uint2 e = myq[qn - count + lane]; resolve(e.x, e.y, av); }
        qn -= count;
        __syncwarp();
    };
    auto push = [&](bool hit, uint32_t j, uint32_t pos, XT av) {
        const uint32_t ball = __ballot_sync(0xffffffffu, hit);
        if (ball) {
            if (hit) myq[qn + __popc(ball & lt_mask)] = make_uint2(j, pos);
            qn += __popc(ball);
            if (qn >= 32) drain(32, av);
        }
    };
    // ---- one warp streams positions [s, e) of B's column array; the trip count is warp-uniform (every lane takes part in the ballots)
    auto stream = [&](uint32_t s, uint32_t e, XT av) {
        const uint32_t p0 = s & ~3u;
        const uint32_t iters = (e - p0 + 127u) >> 7;
        for (uint32_t it = 0; it < iters; ++it) {
            const uint32_t pb = p0 + it * 128u + (uint32_t)lane * 4u;
            const bool act = pb < e;
            uint4 c = make_uint4(0u, 0u, 0u, 0u);
            if (act) c = __ldg(reinterpret_cast<const uint4 *>(p.b_col + pb));        // arrays are padded: reading past e is safe
            const uint32_t jj[4] = {c.x, c.y, c.z, c.w};
            const bool inner = pb >= s && pb + 4u <= e;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t b = bit_of(jj[i]);
                bool hit = act && ((bm[b >> 5] >> (b & 31)) & 1u);
                if (!inner) hit = hit && (pb + i >= s) && (pb + i < e);
                push(hit, jj[i], pb + i, av);
            }
        }
    };

    while (true) {
        if (tid == 0) s_next = atomicAdd(sa.queue, (unsigned int)sa.grab);
        __syncthreads();
        const int64_t c_begin = s_next;
        __syncthreads();
        if (c_begin >= sa.nchunks) break;
        const int64_t c_end = min(sa.nchunks, c_begin + (int64_t)sa.grab);
        for (int64_t ch = c_begin; ch < c_end; ++ch) {
            const int64_t row = sa.chunk_row[ch];
            const uint32_t part = sa.chunk_idx[ch], nparts = sa.chunk_cnt[ch];
            if (row != cur_row) { leave_row(); enter_row(row, nparts > 1); }
            const uint32_t as = p.a_ptr[row], ae = p.a_ptr[row + 1], alen = ae - as;
            const uint32_t c0 = as + (uint32_t)(((uint64_t)alen * part) / nparts), c1 = as + (uint32_t)(((uint64_t)alen * (part + 1)) / nparts);
            for (uint32_t base = c0; base < c1; base += NT) {
                // the batch's A entries: B row range and A value, one per thread
                const uint32_t pa = base + tid;
                uint32_t bs = 0, len = 0; XT av = (XT)1;
                if (pa < c1) {
                    const uint32_t k = __ldg(p.a_col + pa);
                    bs = __ldg(p.b_ptr + k); len = __ldg(p.b_ptr + k + 1) - bs;
                    if (p.need_a) av = gload<XT>(aval + pa);
                }
                if (tid == 0) s_nlong = 0;
                __syncthreads();
                s_bs[tid] = bs; s_len[tid] = len; s_av[tid] = av;
                if (len > STREAM_LONG_ROW) s_long[atomicAdd(&s_nlong, 1u)] = (uint16_t)tid;
                __syncthreads();
                const int nent = (int)min((uint32_t)NT, c1 - base);
                // short B rows: one warp per A entry
                for (int e = warp; e < nent; e += NW) {
                    const uint32_t l = s_len[e];
                    if (l == 0 || l > STREAM_LONG_ROW) continue;
                    const uint32_t b0 = s_bs[e]; const XT a = s_av[e];
                    stream(b0, b0 + l, a);
                    if (qn) drain(qn, a);                      // the queue is per A entry (one A value)
                }
                // long B rows: 128-position blocks dealt to all the warps
                const int nlong = (int)s_nlong;
                for (int x = 0; x < nlong; ++x) {
                    const int e = s_long[x];
                    const uint32_t b0 = s_bs[e], l = s_len[e]; const XT a = s_av[e];
                    const uint32_t p0 = b0 & ~3u, nblk = (b0 + l - p0 + 127u) >> 7;
                    for (uint32_t blk = warp; blk < nblk; blk += NW) {
                        const uint32_t s = max(b0, p0 + blk * 128u), en = min(b0 + l, p0 + (blk + 1u) * 128u);
                        stream(s, en, a);
                    }
                    if (qn) drain(qn, a);
                }
                __syncthreads();
            }
        }
    }
    leave_row();
}
