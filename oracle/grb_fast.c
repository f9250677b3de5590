/* grb_fast.c -- typed OpenMP CPU kernels for the benchmark's CPU baseline.  TEST
 * INFRASTRUCTURE ONLY (bench.py `cpu_baseline` and `--impl reference` legs, tests/).
 *
 * SuiteSparse:GraphBLAS (the library behind /root/reference/pygraphblas/matrix.py:2574,2716)
 * cannot be built here, so the "reference on the host cores" arm is this port: the same
 * row-parallel CSR dot-product SpMV and Gustavson/masked-dot SpGEMM that SuiteSparse's
 * mxm uses for these shapes, one typed loop per benchmarked
 * semiring, parallelised over rows with OpenMP.  Each kernel is validated against the
 * generic oracle (grb_oracle.c) in tests/test_oracle.py.
 */
#include <stdint.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>

int fast_num_threads(void) { return omp_get_max_threads(); }
/* use n threads (bench.py passes os.cpu_count(): torchrun exports OMP_NUM_THREADS=1) */
void fast_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* w = A (+.x) u, PLUS_TIMES FP32, dense u; present[r] = row non-empty */
void fast_spmv_plus_times_f32(int64_t nrows, const int64_t *ptr, const uint32_t *col, const float *val,
                              const float *u, float *w, uint8_t *present) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t r = 0; r < nrows; ++r) {
        float acc = 0.0f;
        for (int64_t k = ptr[r]; k < ptr[r + 1]; ++k) acc += val[k] * u[col[k]];
        w[r] = acc; present[r] = ptr[r + 1] > ptr[r];
    }
}

/* The same SpMV on a "plan": the CSR is copied once into buffers that are first touched by the thread that will
 * stream them (nnz-balanced contiguous row ranges, one per thread), so that on a multi-socket host every thread
 * reads its slice of A from its own NUMA node; u gathers are prefetched 16 entries ahead.  bench.py measures both
 * variants and reports the faster one: the CPU baseline gets its best configuration. */
typedef struct { int nt; int64_t nrows; int64_t *row0; int64_t **ptr; uint32_t **col; float **val; } fast_plan_f32;

void *fast_spmv_plan_f32(int64_t nrows, const int64_t *ptr, const uint32_t *col, const float *val, int nthreads) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    fast_plan_f32 *pl = (fast_plan_f32 *)calloc(1, sizeof(fast_plan_f32));
    pl->nt = nthreads; pl->nrows = nrows;
    pl->row0 = (int64_t *)calloc((size_t)nthreads + 1, sizeof(int64_t));
    pl->ptr = (int64_t **)calloc((size_t)nthreads, sizeof(void *));
    pl->col = (uint32_t **)calloc((size_t)nthreads, sizeof(void *));
    pl->val = (float **)calloc((size_t)nthreads, sizeof(void *));
    const int64_t nnz = ptr[nrows];
    for (int t = 0; t <= nthreads; ++t) {                 /* first row whose offset reaches the t-th share of nnz */
        const int64_t target = t == nthreads ? nnz + 1 : nnz * t / nthreads;
        int64_t a = 0, b = nrows;
        while (a < b) { const int64_t m = (a + b) >> 1; if (ptr[m] < target) a = m + 1; else b = m; }
        pl->row0[t] = t == 0 ? 0 : (t == nthreads ? nrows : a);
    }
    /* one partition per loop iteration, static schedule: with a full team iteration t runs on thread t both here and in
     * the run, which is what places the pages; a smaller team still covers every partition */
#pragma omp parallel for schedule(static, 1) num_threads(nthreads)
    for (int t = 0; t < nthreads; ++t) {
        const int64_t r0 = pl->row0[t], r1 = pl->row0[t + 1], k0 = ptr[r0], k1 = ptr[r1], nr = r1 - r0, nk = k1 - k0;
        int64_t *lp = (int64_t *)malloc(((size_t)nr + 1) * sizeof(int64_t));
        uint32_t *lc = (uint32_t *)malloc(((size_t)nk + 16) * sizeof(uint32_t));
        float *lv = (float *)malloc(((size_t)nk + 16) * sizeof(float));
        for (int64_t r = 0; r <= nr; ++r) lp[r] = ptr[r0 + r] - k0;
        for (int64_t k = 0; k < nk; ++k) { lc[k] = col[k0 + k]; lv[k] = val[k0 + k]; }
        for (int64_t k = nk; k < nk + 16; ++k) { lc[k] = nk ? lc[nk - 1] : 0; lv[k] = 0.0f; }
        pl->ptr[t] = lp; pl->col[t] = lc; pl->val[t] = lv;
    }
    return pl;
}
void fast_spmv_plan_run_f32(void *plan, const float *u, float *w, uint8_t *present) {
    fast_plan_f32 *pl = (fast_plan_f32 *)plan;
#pragma omp parallel for schedule(static, 1) num_threads(pl->nt)
    for (int t = 0; t < pl->nt; ++t) {
        const int64_t r0 = pl->row0[t], nr = pl->row0[t + 1] - r0;
        const int64_t *lp = pl->ptr[t]; const uint32_t *lc = pl->col[t]; const float *lv = pl->val[t];
        for (int64_t r = 0; r < nr; ++r) {
            float acc = 0.0f;
            const int64_t ke = lp[r + 1];
            for (int64_t k = lp[r]; k < ke; ++k) {
                __builtin_prefetch(&u[lc[k + 16]], 0, 1);
                acc += lv[k] * u[lc[k]];
            }
            w[r0 + r] = acc; present[r0 + r] = ke > lp[r];
        }
    }
}
void fast_spmv_plan_free_f32(void *plan) {
    fast_plan_f32 *pl = (fast_plan_f32 *)plan;
    if (!pl) return;
    for (int t = 0; t < pl->nt; ++t) { free(pl->ptr[t]); free(pl->col[t]); free(pl->val[t]); }
    free(pl->ptr); free(pl->col); free(pl->val); free(pl->row0); free(pl);
}

/* w = A (+.x) u, PLUS_TIMES FP64, dense u */
void fast_spmv_plus_times_f64(int64_t nrows, const int64_t *ptr, const uint32_t *col, const double *val,
                              const double *u, double *w, uint8_t *present) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t r = 0; r < nrows; ++r) {
        double acc = 0.0;
        for (int64_t k = ptr[r]; k < ptr[r + 1]; ++k) acc += val[k] * u[col[k]];
        w[r] = acc; present[r] = ptr[r + 1] > ptr[r];
    }
}

/* MIN_PLUS FP32 with accum MIN into w (SSSP sweep): w[r] = min(w[r], min_k a(r,k) + u[k]) */
void fast_spmv_min_plus_f32_accum(int64_t nrows, const int64_t *ptr, const uint32_t *col, const float *val,
                                  const float *u, float *w) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t r = 0; r < nrows; ++r) {
        float acc = INFINITY;
        for (int64_t k = ptr[r]; k < ptr[r + 1]; ++k) acc = fminf(acc, val[k] + u[col[k]]);
        w[r] = fminf(w[r], acc);
    }
}

/* BFS step: next<!visited, replace> = A lor.land frontier (pattern A, byte maps); returns |next| */
int64_t fast_bfs_step(int64_t nrows, const int64_t *ptr, const uint32_t *col,
                      const uint8_t *frontier, const uint8_t *visited, uint8_t *next) {
    int64_t count = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : count)
    for (int64_t r = 0; r < nrows; ++r) {
        uint8_t hit = 0;
        if (!visited[r])
            for (int64_t k = ptr[r]; k < ptr[r + 1]; ++k) if (frontier[col[k]]) { hit = 1; break; }
        next[r] = hit; count += hit;
    }
    return count;
}

/* masked SpGEMM C<L> = L (+.pair) L' form used by triangle counting: for every mask entry (i,j)
 * count |L(i,:) ^ L(j,:)| by sorted merge; cval[k] for mask entry k, chas[k] = count > 0.
 * (dot3 method of SuiteSparse for C<M> = A*B' with a sparse structural mask.) */
void fast_masked_dot_plus_pair_i64(int64_t nrows, const int64_t *mptr, const uint32_t *mcol,
                                   const int64_t *aptr, const uint32_t *acol,
                                   const int64_t *bptr, const uint32_t *bcol,
                                   int64_t *cval, uint8_t *chas) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nrows; ++i) {
        for (int64_t p = mptr[i]; p < mptr[i + 1]; ++p) {
            const int64_t j = mcol[p];
            int64_t x = aptr[i], xe = aptr[i + 1], y = bptr[j], ye = bptr[j + 1], c = 0;
            while (x < xe && y < ye) {
                const uint32_t cx = acol[x], cy = bcol[y];
                c += cx == cy; x += cx <= cy; y += cy <= cx;
            }
            cval[p] = c; chas[p] = c > 0;
        }
    }
}

/* masked Gustavson C<M> = A (+.pair) B (structural mask M, row-wise saxpy with the mask row
 * scattered into a per-thread dense marker): cval/chas per mask entry. */
void fast_masked_saxpy_plus_pair_i64(int64_t nrows, int64_t ncols, const int64_t *mptr, const uint32_t *mcol,
                                     const int64_t *aptr, const uint32_t *acol,
                                     const int64_t *bptr, const uint32_t *bcol,
                                     int64_t *cval, uint8_t *chas) {
#pragma omp parallel
    {
        int64_t *slot = malloc((size_t)ncols * sizeof(int64_t));
        for (int64_t j = 0; j < ncols; ++j) slot[j] = -1;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < nrows; ++i) {
            for (int64_t p = mptr[i]; p < mptr[i + 1]; ++p) { slot[mcol[p]] = p; cval[p] = 0; }
            for (int64_t pa = aptr[i]; pa < aptr[i + 1]; ++pa) {
                const int64_t k = acol[pa];
                for (int64_t pb = bptr[k]; pb < bptr[k + 1]; ++pb) { const int64_t s = slot[bcol[pb]]; if (s >= 0) cval[s]++; }
            }
            for (int64_t p = mptr[i]; p < mptr[i + 1]; ++p) { slot[mcol[p]] = -1; chas[p] = cval[p] > 0; }
        }
        free(slot);
    }
}

/* unmasked Gustavson SpGEMM, PLUS_SECOND FP32 (C = A*B, value = sum of B's values):
 * pass 1 (symbolic) fills cptr[r+1] = nnz(C(r,:)); the caller prefix-sums; pass 2 fills ccol/cval. */
void fast_spgemm_symbolic(int64_t nrows, int64_t ncols, const int64_t *aptr, const uint32_t *acol,
                          const int64_t *bptr, const uint32_t *bcol, int64_t *cptr) {
    cptr[0] = 0;
#pragma omp parallel
    {
        int64_t *mark = malloc((size_t)ncols * sizeof(int64_t));
        for (int64_t j = 0; j < ncols; ++j) mark[j] = -1;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < nrows; ++i) {
            int64_t c = 0;
            for (int64_t pa = aptr[i]; pa < aptr[i + 1]; ++pa) {
                const int64_t k = acol[pa];
                for (int64_t pb = bptr[k]; pb < bptr[k + 1]; ++pb) { const uint32_t j = bcol[pb]; if (mark[j] != i) { mark[j] = i; ++c; } }
            }
            cptr[i + 1] = c;
        }
        free(mark);
    }
}

static int cmp_u32(const void *p, const void *q) { uint32_t a = *(const uint32_t *)p, b = *(const uint32_t *)q; return a < b ? -1 : a > b; }

void fast_spgemm_numeric_plus_second_f32(int64_t nrows, int64_t ncols, const int64_t *aptr, const uint32_t *acol,
                                         const int64_t *bptr, const uint32_t *bcol, const float *bval,
                                         const int64_t *cptr, uint32_t *ccol, float *cval) {
#pragma omp parallel
    {
        int64_t *mark = malloc((size_t)ncols * sizeof(int64_t));
        float *acc = malloc((size_t)ncols * sizeof(float));
        for (int64_t j = 0; j < ncols; ++j) mark[j] = -1;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < nrows; ++i) {
            int64_t c = cptr[i];
            for (int64_t pa = aptr[i]; pa < aptr[i + 1]; ++pa) {
                const int64_t k = acol[pa];
                for (int64_t pb = bptr[k]; pb < bptr[k + 1]; ++pb) {
                    const uint32_t j = bcol[pb];
                    if (mark[j] != i) { mark[j] = i; acc[j] = bval[pb]; ccol[c++] = j; }
                    else acc[j] += bval[pb];
                }
            }
            qsort(ccol + cptr[i], (size_t)(c - cptr[i]), sizeof(uint32_t), cmp_u32);
            for (int64_t p = cptr[i]; p < c; ++p) cval[p] = acc[ccol[p]];
        }
        free(mark); free(acc);
    }
}
