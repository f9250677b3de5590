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
