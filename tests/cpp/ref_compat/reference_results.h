// TEST INFRASTRUCTURE -- never part of the product.
//
// The reference's own tests and benchmark drivers call compute_reference_results on the module / app
// classes (/root/reference/tests/test_app.cpp:66,100,121, tests/test_module_spmv_spmspv.cpp:124-126,
// tests/test_module_apply.cpp:72-258, benchmark/bench_*.cpp).  include/graphlily DECLARES those methods
// and this header DEFINES them over the parity oracle (oracle/oracle.h, liboracle.so: the plain-C
// restatement pinned bit-for-bit to the compiled reference by tests/test_oracle_vs_ref.py), so that those
// files compile UNMODIFIED (g++ -include reference_results.h) and check the CUDA path against the CPU path.
#ifndef GLB_REF_COMPAT_REFERENCE_RESULTS_H_
#define GLB_REF_COMPAT_REFERENCE_RESULTS_H_

#include "graphlily/app/bfs.h"
#include "graphlily/app/pagerank.h"
#include "graphlily/app/sssp.h"
#include "graphlily/module/add_scalar_vector_dense_module.h"
#include "graphlily/module/assign_vector_dense_module.h"
#include "graphlily/module/assign_vector_sparse_module.h"
#include "graphlily/module/spmspv_module.h"
#include "graphlily/module/spmv_module.h"

extern "C" {
#include "../../../oracle/oracle.h"
}

namespace graphlily {
namespace module {

template <typename M, typename V>
aligned_dense_float_vec_t SpMVModule<M, V>::compute_reference_results(aligned_dense_float_vec_t &vector) {
    const CSRMatrix<float> &m = csr_matrix_float_;
    aligned_dense_float_vec_t y(m.num_rows);
    oracle_spmv(m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), int(semiring_.op),
                float(semiring_.zero), GLB_MASK_NONE, vector.data(), nullptr, y.data());
    return y;
}

template <typename M, typename V>
aligned_dense_float_vec_t SpMVModule<M, V>::compute_reference_results(aligned_dense_float_vec_t &vector,
                                                                       aligned_dense_float_vec_t &mask) {
    const CSRMatrix<float> &m = csr_matrix_float_;
    aligned_dense_float_vec_t y(m.num_rows);
    oracle_spmv(m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), int(semiring_.op),
                float(semiring_.zero), int(mask_type_), vector.data(), mask.data(), y.data());
    return y;
}

template <typename M, typename V, typename IV>
aligned_dense_float_vec_t SpMSpVModule<M, V, IV>::compute_reference_results(aligned_sparse_float_vec_t &vector,
                                                                            aligned_dense_float_vec_t &mask) {
    const CSCMatrix<float> &m = csc_matrix_float_;
    const uint32_t nnz = vector[0].index;
    std::vector<uint32_t> idx(nnz);
    std::vector<float> val(nnz);
    for (uint32_t i = 0; i < nnz; i++) { idx[i] = vector[i + 1].index; val[i] = vector[i + 1].val; }
    aligned_dense_float_vec_t y(m.num_rows);
    oracle_spmspv(m.num_rows, m.num_cols, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), int(semiring_.op),
                  float(semiring_.zero), int(mask_type_), idx.data(), val.data(), nnz,
                  mask_type_ == kNoMask ? nullptr : mask.data(), y.data());
    return y;
}

template <typename V>
aligned_dense_float_vec_t eWiseAddModule<V>::compute_reference_results(aligned_dense_float_vec_t const &in, uint32_t len,
                                                                       float val) {
    aligned_dense_float_vec_t out(len);
    oracle_ewise_add(in.data(), out.data(), len, val);
    return out;
}

template <typename V>
void AssignVectorDenseModule<V>::compute_reference_results(aligned_dense_float_vec_t &mask, aligned_dense_float_vec_t &inout,
                                                           uint32_t len, float val) {
    oracle_assign_dense(mask.data(), inout.data(), len, val, int(mask_type_));
}

template <typename V, typename SV>
void AssignVectorSparseModule<V, SV>::compute_reference_results(aligned_sparse_float_vec_t &mask,
                                                                aligned_dense_float_vec_t &inout, float val) {
    const uint32_t nnz = mask[0].index;
    std::vector<uint32_t> idx(nnz);
    for (uint32_t i = 0; i < nnz; i++) idx[i] = mask[i + 1].index;
    oracle_assign_sparse(idx.data(), nnz, inout.data(), val);
}

template <typename V, typename SV>
void AssignVectorSparseModule<V, SV>::compute_reference_results(aligned_sparse_float_vec_t &mask,
                                                                aligned_dense_float_vec_t &inout,
                                                                aligned_sparse_float_vec_t &new_frontier) {
    const uint32_t nnz = mask[0].index;
    std::vector<uint32_t> idx(nnz), nf_i(nnz + 1);
    std::vector<float> val(nnz), nf_v(nnz + 1);
    for (uint32_t i = 0; i < nnz; i++) { idx[i] = mask[i + 1].index; val[i] = mask[i + 1].val; }
    const int cnt = oracle_assign_sparse_relax(idx.data(), val.data(), nnz, inout.data(), nf_i.data(), nf_v.data());
    new_frontier.clear();   // assign_vector_sparse_module.h:324-334: head {count, 0} first
    idx_float_t head;
    head.index = uint32_t(cnt);
    head.val = 0;
    new_frontier.push_back(head);
    for (int i = 0; i < cnt; i++) {
        idx_float_t e;
        e.index = nf_i[i];
        e.val = nf_v[i];
        new_frontier.push_back(e);
    }
}

}  // namespace module

namespace app {

inline aligned_dense_float_vec_t BFS::compute_reference_results(uint32_t source, uint32_t num_iterations) {
    const io::CSRMatrix<float> &m = SpMV_->host_matrix();
    aligned_dense_float_vec_t out(m.num_rows);
    oracle_bfs(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), source, num_iterations, out.data());
    return out;
}

inline aligned_dense_float_vec_t PageRank::compute_reference_results(float damping, uint32_t num_iterations) {
    const io::CSRMatrix<float> &m = SpMV_->host_matrix();
    aligned_dense_float_vec_t out(m.num_rows);
    oracle_pagerank(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), damping, num_iterations,
                    out.data());
    return out;
}

inline aligned_dense_float_vec_t SSSP::compute_reference_results(uint32_t source, uint32_t num_iterations) {
    const io::CSRMatrix<float> &m = SpMV_->host_matrix();
    aligned_dense_float_vec_t out(m.num_rows);
    oracle_sssp(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), source, num_iterations,
                float(TropicalSemiring.zero), out.data());
    return out;
}

}  // namespace app
}  // namespace graphlily

#endif  // GLB_REF_COMPAT_REFERENCE_RESULTS_H_
