/* TEST INFRASTRUCTURE ONLY -- C prototypes of oracle/liboracle.so (oracle.c) for the C++ tests
 * under tests/cpp.  Never included by the product (graphlily_b200/, include/). */
#ifndef GRAPHLILY_B200_ORACLE_H_
#define GRAPHLILY_B200_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int oracle_spmv(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const float *data,
                int op, float zero, int mask_type, const float *x, const float *mask, float *y);
int oracle_spmspv(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const float *data,
                  int op, float zero, int mask_type, const uint32_t *x_idx, const float *x_val, uint32_t x_nnz,
                  const float *mask, float *y);
int oracle_ewise_add(const float *in, float *out, uint32_t len, float val);
int oracle_assign_dense(const float *mask, float *inout, uint32_t len, float val, int mask_type);
int oracle_assign_sparse(const uint32_t *m_idx, uint32_t m_nnz, float *inout, float val);
int oracle_assign_sparse_relax(const uint32_t *m_idx, const float *m_val, uint32_t m_nnz, float *inout,
                               uint32_t *nf_idx, float *nf_val);
int oracle_csr2csc(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices, const float *data,
                   uint32_t *o_indptr, uint32_t *o_indices, float *o_data);
int oracle_round_dim(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, uint32_t row_div, uint32_t col_div,
                     uint32_t *o_indptr, uint32_t *out_dims);
int oracle_normalize_outdegree(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                               float *data);
int64_t oracle_sssp_preprocess(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                               const float *data, uint32_t *o_indptr, uint32_t *o_indices, float *o_data);
int oracle_bfs(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data, uint32_t source,
               uint32_t iters, float *distance);
int oracle_pagerank(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data, float damping,
                    uint32_t iters, float *rank);
int oracle_sssp(uint32_t n, const uint32_t *indptr, const uint32_t *indices, const float *data, uint32_t source,
                uint32_t iters, float zero, float *dist);
double oracle_spmv_timed(uint32_t nrows, uint32_t ncols, const uint32_t *indptr, const uint32_t *indices,
                         const float *data, int op, float zero, const float *x, float *y, int reps);
#ifdef __cplusplus
}
#endif
#endif
