/* TEST INFRASTRUCTURE ONLY -- the parity oracle.  Never imported, linked or executed
 * by the product path (graphlily_b200/, include/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * A plain-C restatement of the CPU algorithms GraphLily ships for the hot path
 * (its compute_reference_results() family and the host IO helpers the apps run
 * first).  Each function cites the reference file:line it follows under
 * /root/reference/graphlily.  Statement order and fp32 evaluation order follow the
 * reference exactly so results are bit-identical to it; build WITHOUT -ffast-math,
 * -march or fp contraction (oracle/Makefile: gcc -O3 -std=c11 -ffp-contract=off).
 *
 * PARITY PINNING: this restatement is pinned against (1) the reference's golden
 * vectors in tests/test_io.cpp (tests/test_oracle_golden.py), (2) the reference's
 * own code compiled from /root/reference into oracle/_ref/libgraphlily_ref.so
 * (oracle/ref_driver.cpp) on seeded inputs, bit-for-bit (tests/test_oracle_vs_ref.py),
 * and (3) committed fixtures under tests/golden/ produced by that library
 * (tests/golden/make_golden.py).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* global.h:83-87 OperationType, :103-107 MaskType */
enum { OP_MUL_ADD = 0, OP_LOGICAL_AND_OR = 1, OP_ADD_MIN = 2 };
enum { MASK_NONE = 0, MASK_WRITE_TO_ZERO = 1, MASK_WRITE_TO_ONE = 2 };

/* global.h:80 FLOAT_INF = 999999999 (converted to fp32 where the reference compares floats). */
static const float ORACLE_FLOAT_INF = 999999999.0f;

/* -------------------------------------------------------------------------- SpMV
 * spmv_module.h:478-510.  y[r] starts at semiring.zero; per CSR row, in CSR order:
 *   kMulAdd       : y += a * x[c]                       (:493-495)
 *   kLogicalAndOr : y = y || (a && x[c])  -> 0.0f/1.0f  (:496-499)
 *   kAddMin       : y = min(y, a + x[c])  (no clamp)    (:500-503)
 * then the mask overload, spmv_module.h:513-532: compares against literal 0 and
 * writes literal 0; any mask_type other than WriteToZero takes the else branch. */
int oracle_spmv(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                const float *data, int op, float zero, int mask_type, const float *x,
                const float *mask, float *y) {
    (void)ncols;
    for (uint32_t r = 0; r < nrows; r++) y[r] = zero;
    switch (op) {
    case OP_MUL_ADD:
        for (uint32_t r = 0; r < nrows; r++)
            for (uint32_t i = indptr[r]; i < indptr[r + 1]; i++) {
                float prod = data[i] * x[indices[i]];
                y[r] += prod;
            }
        break;
    case OP_LOGICAL_AND_OR:
        for (uint32_t r = 0; r < nrows; r++)
            for (uint32_t i = indptr[r]; i < indptr[r + 1]; i++)
                y[r] = (float)(y[r] || (data[i] && x[indices[i]]));
        break;
    case OP_ADD_MIN:
        for (uint32_t r = 0; r < nrows; r++)
            for (uint32_t i = indptr[r]; i < indptr[r + 1]; i++) {
                float s = data[i] + x[indices[i]];
                y[r] = (s < y[r]) ? s : y[r]; /* std::min(y, s): returns y unless s < y */
            }
        break;
    default:
        return 1;
    }
    if (mask_type != MASK_NONE) {
        if (mask_type == MASK_WRITE_TO_ZERO) {
            for (uint32_t i = 0; i < nrows; i++) if (mask[i] != 0) y[i] = 0;
        } else {
            for (uint32_t i = 0; i < nrows; i++) if (mask[i] == 0) y[i] = 0;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------ SpMSpV
 * spmspv_module.h:446-520.  CSC matrix; x is the list of (index,val) active entries in
 * list order.  y starts at semiring.zero; kAddMin clamps at FLOAT_INF (:482-491); the mask
 * compares against semiring.zero and masked rows get semiring.zero (:499-516). */
int oracle_spmspv(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                  const float *data, int op, float zero, int mask_type, const uint32_t *x_idx,
                  const float *x_val, uint32_t x_nnz, const float *mask, float *y) {
    (void)ncols;
    for (uint32_t r = 0; r < nrows; r++) y[r] = zero;
    for (uint32_t k = 0; k < x_nnz; k++) {
        float v = x_val[k];
        uint32_t c = x_idx[k];
        for (uint32_t i = indptr[c]; i < indptr[c + 1]; i++) {
            uint32_t r = indices[i];
            float a = data[i];
            float incr;
            switch (op) {
            case OP_MUL_ADD:
                incr = a * v;
                y[r] += incr;
                break;
            case OP_LOGICAL_AND_OR:
                incr = (float)(a && v);
                y[r] = (float)(y[r] || incr);
                break;
            case OP_ADD_MIN:
                if (a > ORACLE_FLOAT_INF || v > ORACLE_FLOAT_INF) {
                    incr = ORACLE_FLOAT_INF;
                } else {
                    incr = a + v;
                    if (incr > ORACLE_FLOAT_INF) incr = ORACLE_FLOAT_INF;
                }
                y[r] = (y[r] < incr) ? y[r] : incr;
                break;
            default:
                return 1;
            }
        }
    }
    for (uint32_t i = 0; i < nrows; i++) {
        int off;
        switch (mask_type) {
        case MASK_NONE: off = 0; break;
        case MASK_WRITE_TO_ONE: off = (mask[i] == zero); break;
        case MASK_WRITE_TO_ZERO: off = (mask[i] != zero); break;
        default: off = 1; break;
        }
        if (off) y[i] = zero;
    }
    return 0;
}

/* --------------------------------------------------------------------- apply ops */
/* add_scalar_vector_dense_module.h:195-204 */
int oracle_ewise_add(const float *in, float *out, uint32_t len, float val) {
    for (uint32_t i = 0; i < len; i++) out[i] = in[i] + val;
    return 0;
}

/* assign_vector_dense_module.h:223-246; kNoMask is an error (print + exit in the reference). */
int oracle_assign_dense(const float *mask, float *inout, uint32_t len, float val, int mask_type) {
    if (mask_type == MASK_WRITE_TO_ZERO) {
        for (uint32_t i = 0; i < len; i++) if (mask[i] == 0) inout[i] = val;
    } else if (mask_type == MASK_WRITE_TO_ONE) {
        for (uint32_t i = 0; i < len; i++) if (mask[i] != 0) inout[i] = val;
    } else {
        return 1;
    }
    return 0;
}

/* assign_vector_sparse_module.h:306-315 (BFS mode) */
int oracle_assign_sparse(const uint32_t *m_idx, uint32_t m_nnz, float *inout, float val) {
    for (uint32_t i = 0; i < m_nnz; i++) inout[m_idx[i]] = val;
    return 0;
}

/* assign_vector_sparse_module.h:318-335 (SSSP mode): relax in list order, emit improved
 * entries in the same order.  Returns the new-frontier count (the {count,0} head). */
int oracle_assign_sparse_relax(const uint32_t *m_idx, const float *m_val, uint32_t m_nnz, float *inout,
                               uint32_t *nf_idx, float *nf_val) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < m_nnz; i++) {
        if (inout[m_idx[i]] > m_val[i]) {
            inout[m_idx[i]] = m_val[i];
            nf_idx[n] = m_idx[i];
            nf_val[n] = m_val[i];
            n++;
        }
    }
    return (int)n;
}

/* ---------------------------------------------------------------------------- IO */
/* data_loader.h:108-144: counting-sort transpose, row order preserved inside a column. */
int oracle_csr2csc(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                   const float *data, uint32_t *o_indptr, uint32_t *o_indices, float *o_data) {
    uint32_t nnz = indptr[nrows];
    uint32_t *consumed = (uint32_t *)calloc(ncols ? ncols : 1, sizeof(uint32_t));
    if (!consumed) return 2;
    for (uint32_t c = 0; c <= ncols; c++) o_indptr[c] = 0;
    for (uint32_t n = 0; n < nnz; n++) o_indptr[indices[n] + 1]++;
    for (uint32_t c = 0; c < ncols; c++) o_indptr[c + 1] += o_indptr[c];
    for (uint32_t r = 0; r < nrows; r++)
        for (uint32_t i = indptr[r]; i < indptr[r + 1]; i++) {
            uint32_t c = indices[i];
            uint32_t dest = o_indptr[c] + consumed[c]++;
            o_indices[dest] = r;
            o_data[dest] = data[i];
        }
    free(consumed);
    return 0;
}

/* data_formatter.h:19-33: pad rows (repeat last indptr) and cols up to multiples.
 * o_indptr must hold round_up(nrows,row_div)+1 entries. */
int oracle_round_dim(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, uint32_t row_div,
                     uint32_t col_div, uint32_t *o_indptr, uint32_t *out_dims) {
    uint32_t nr = nrows, nc = ncols;
    memcpy(o_indptr, indptr, sizeof(uint32_t) * ((size_t)nrows + 1));
    if (nrows % row_div != 0) {
        uint32_t pad = row_div - nrows % row_div;
        for (uint32_t i = 0; i < pad; i++) o_indptr[nrows + 1 + i] = indptr[nrows];
        nr += pad;
    }
    if (ncols % col_div != 0) nc += col_div - ncols % col_div;
    out_dims[0] = nr;
    out_dims[1] = nc;
    return 0;
}

/* data_formatter.h:37-51: data[i] = 1.0 / (#nnz in column of i), double divide -> float. */
int oracle_normalize_outdegree(uint32_t nrows, uint32_t ncols, const uint32_t *indptr,
                               const uint32_t *indices, float *data) {
    uint32_t nnz = indptr[nrows];
    uint32_t *cnt = (uint32_t *)calloc(ncols ? ncols : 1, sizeof(uint32_t));
    if (!cnt) return 2;
    for (uint32_t i = 0; i < nnz; i++) cnt[indices[i]]++;
    for (uint32_t i = 0; i < nnz; i++) data[i] = (float)(1.0 / cnt[indices[i]]);
    free(cnt);
    return 0;
}

/* sssp.h:16-62 _preprocess, restated in O(nnz) (the reference uses vector::insert, O(N*nnz)).
 * All weights <- 1 (:18-20); then every row is meant to get a weight-0 diagonal.  The reference
 * reads `start` from the indptr entry it has already rewritten but `end` from the one it has
 * not (:31-32, :58), so after k insertions in earlier rows it scans only the first len-k
 * elements of a row of length len.  That is observable behaviour and is reproduced here:
 *   len == k (incl. a truly empty row when k == 0): diagonal inserted at the row start (:33-36);
 *   len <  k: row left untouched (the for loop at :39 does not execute);
 *   else scan elements 0..len-k-1: existing diagonal zeroed in place (:41-43); else inserted
 *   before the first column > row (:44-48), or before the LAST SCANNED element (:49-53).
 * With every diagonal present (how the benches generate graphs) k stays 0 and the function
 * just zeroes the diagonal weights.  Outputs sized nnz + nrows; returns the new nnz. */
int64_t oracle_sssp_preprocess(uint32_t nrows, uint32_t ncols, const uint32_t *indptr,
                               const uint32_t *indices, const float *data, uint32_t *o_indptr,
                               uint32_t *o_indices, float *o_data) {
    (void)ncols;
    (void)data;
    uint64_t k = 0; /* insertions so far */
    uint32_t w = 0;
    o_indptr[0] = 0;
    for (uint32_t r = 0; r < nrows; r++) {
        uint32_t s = indptr[r], e = indptr[r + 1];
        uint64_t len = (uint64_t)e - s;
        int64_t ins = -1;  /* source index before which (r, 0) is inserted; -1 = none */
        int64_t zero_at = -1; /* source index of an existing diagonal to zero */
        if (len == k) {
            ins = s;
        } else if (len > k) {
            uint32_t scan_end = (uint32_t)(e - k);
            for (uint32_t i = s; i < scan_end; i++) {
                if (indices[i] == r) { zero_at = i; break; }
                if (indices[i] > r) { ins = i; break; }
                if (i == scan_end - 1) { ins = i; break; }
            }
        }
        if (ins == (int64_t)s && len == k) { o_indices[w] = r; o_data[w] = 0.0f; w++; }
        for (uint32_t i = s; i < e; i++) {
            if (len != k && ins == (int64_t)i) { o_indices[w] = r; o_data[w] = 0.0f; w++; }
            o_indices[w] = indices[i];
            o_data[w] = (zero_at == (int64_t)i) ? 0.0f : 1.0f;
            w++;
        }
        if (ins >= 0) k++;
        o_indptr[r + 1] = w;
    }
    return (int64_t)w;
}

/* ------------------------------------------------------------------------- apps
 * All three take the matrix AFTER the app's load_and_format_matrix preprocessing
 * (bfs.h:84-97, pagerank.h:60-73, sssp.h:128-143), square n x n. */

/* bfs.h:350-360: input=zero(0), input[src]=1, distance=0, distance[src]=1;
 * per iter: input = SpMV_ref(input, distance) [Logical, WriteToZero]; DenseAssign_ref(input,
 * distance, n, iter+1) [WriteToOne]. */
int oracle_bfs(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data,
               uint32_t source, uint32_t iters, float *distance) {
    float *in = (float *)calloc(n ? n : 1, sizeof(float));
    float *nx = (float *)malloc(sizeof(float) * (n ? n : 1));
    if (!in || !nx) return 2;
    for (uint32_t i = 0; i < n; i++) distance[i] = 0;
    in[source] = 1;
    distance[source] = 1;
    for (uint32_t it = 1; it <= iters; it++) {
        oracle_spmv(n, n, indptr, indices, data, OP_LOGICAL_AND_OR, 0.0f, MASK_WRITE_TO_ZERO, in, distance, nx);
        float *t = in; in = nx; nx = t;
        oracle_assign_dense(in, distance, n, (float)(it + 1), MASK_WRITE_TO_ONE);
    }
    free(in);
    free(nx);
    return 0;
}

/* pagerank.h:150-159: rank0 = float(1.0/n); rank = SpMV_ref(rank); rank += (1-damping)/n
 * with the teleport term evaluated in float (val_t damping; `(1 - damping) / n`). */
int oracle_pagerank(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data,
                    float damping, uint32_t iters, float *rank) {
    float *nx = (float *)malloc(sizeof(float) * (n ? n : 1));
    if (!nx) return 2;
    float r0 = (float)(1.0 / n);
    float teleport = (1 - damping) / n;
    for (uint32_t i = 0; i < n; i++) rank[i] = r0;
    for (uint32_t it = 1; it <= iters; it++) {
        oracle_spmv(n, n, indptr, indices, data, OP_MUL_ADD, 0.0f, MASK_NONE, rank, 0, nx);
        oracle_ewise_add(nx, rank, n, teleport);
    }
    free(nx);
    return 0;
}

/* sssp.h:246-253: dist = zero (255), dist[src]=0; dist = SpMV_ref(dist) [Tropical, no mask]. */
int oracle_sssp(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data,
                uint32_t source, uint32_t iters, float zero, float *dist) {
    float *nx = (float *)malloc(sizeof(float) * (n ? n : 1));
    if (!nx) return 2;
    for (uint32_t i = 0; i < n; i++) dist[i] = zero;
    dist[source] = 0;
    for (uint32_t it = 1; it <= iters; it++) {
        oracle_spmv(n, n, indptr, indices, data, OP_ADD_MIN, zero, MASK_NONE, dist, 0, nx);
        memcpy(dist, nx, sizeof(float) * n);
    }
    free(nx);
    return 0;
}

/* ------------------------------------------------------------- CPU baseline timer
 * Best-of-`reps` wall-clock seconds of one oracle_spmv() call (1 thread, as the reference). */
double oracle_spmv_timed(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                         const float *data, int op, float zero, const float *x, float *y, int reps) {
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        oracle_spmv(nrows, ncols, indptr, indices, data, op, zero, MASK_NONE, x, 0, y);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (s < best) best = s;
    }
    return best;
}
