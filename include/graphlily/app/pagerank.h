// graphlily-b200: PageRank by power iteration, rank <- d * A_norm * rank + (1 - d) / N.
//
// Same surface and constants as /root/reference/graphlily/app/pagerank.h:17-159: values
// float(1.0 / colcount) then * damping in float (:60-73, data_formatter.h:37-51), rank0 =
// float(1.0 / N) (:151), teleport (1 - damping) / N evaluated in float (:156); dangling columns
// and padded rows get no special handling.  With fused() on (default) the eWiseAdd of :86-88 rides
// in the SpMV write-back and vector / results swap instead of being bound to each other.
#ifndef GRAPHLILY_PAGERANK_H_
#define GRAPHLILY_PAGERANK_H_

#include <utility>

#include "graphlily/app/module_collection.h"
#include "graphlily/io/data_formatter.h"
#include "graphlily/io/data_loader.h"
#include "graphlily/module/add_scalar_vector_dense_module.h"
#include "graphlily/module/spmv_module.h"

namespace graphlily {
namespace app {

class PageRank : public app::ModuleCollection {
private:
    module::SpMVModule<graphlily::val_t, graphlily::val_t> *SpMV_;
    module::eWiseAddModule<graphlily::val_t> *eWiseAdd_;
    uint32_t matrix_num_rows_ = 0, matrix_num_cols_ = 0;
    uint32_t num_channels_, spmv_out_buf_len_, vec_buf_len_;
    graphlily::SemiringType semiring_ = graphlily::ArithmeticSemiring;
    bool fused_ = true;
    using aligned_dense_vec_t = graphlily::aligned_dense_vec_t;

public:
    PageRank(uint32_t num_channels, uint32_t spmv_out_buf_len, uint32_t vec_buf_len)
        : num_channels_(num_channels), spmv_out_buf_len_(spmv_out_buf_len), vec_buf_len_(vec_buf_len) {
        SpMV_ = new module::SpMVModule<graphlily::val_t, graphlily::val_t>(num_channels_, spmv_out_buf_len_, vec_buf_len_);
        SpMV_->set_semiring(semiring_);
        SpMV_->set_mask_type(graphlily::kNoMask);
        add_module(SpMV_);
        eWiseAdd_ = new module::eWiseAddModule<graphlily::val_t>();
        add_module(eWiseAdd_);
    }

    void set_fused(bool fused) { fused_ = fused; }
    uint32_t get_nnz() { return SpMV_->get_nnz(); }
    uint32_t get_num_rows() { return matrix_num_rows_; }

    void load_and_format_matrix(graphlily::io::CSRMatrix<float> csr_matrix, float damping, bool skip_empty_rows) {
        graphlily::io::util_round_csr_matrix_dim(csr_matrix, num_channels_ * graphlily::pack_size,
                                                 num_channels_ * graphlily::pack_size);
        graphlily::io::util_normalize_csr_matrix_by_outdegree(csr_matrix);
        for (auto &x : csr_matrix.adj_data) x = x * damping;
        SpMV_->load_and_format_matrix(csr_matrix, skip_empty_rows);
        matrix_num_rows_ = SpMV_->get_num_rows();
        matrix_num_cols_ = SpMV_->get_num_cols();
        assert(matrix_num_rows_ == matrix_num_cols_);
    }
    void load_and_format_matrix(std::string csr_float_npz_path, float damping, bool skip_empty_rows) {
        load_and_format_matrix(graphlily::io::load_csr_matrix_from_float_npz(csr_float_npz_path), damping, skip_empty_rows);
    }

    void send_matrix_host_to_device() {
        drop_recorded_sequences();
        if (world_ > 1) make_cuts(SpMV_->host_matrix().adj_indptr, matrix_num_rows_);
        SpMV_->send_matrix_host_to_device(row_begin(), row_end(matrix_num_rows_));
        if (exchange_) {   // the vectors of the loop live in the exchange: every rank sees every slice
            assert(exchange_->size() == matrix_num_rows_ && exchange_->vectors() >= 2);
            SpMV_->vector_buf = exchange_->buffer(0);
            SpMV_->results_buf = exchange_->buffer(1);
        }
    }

    aligned_dense_vec_t pull(float damping, uint32_t num_iterations) {
        const uint32_t n = matrix_num_rows_;
        const float teleport = (1 - damping) / n;
        if (exchange_) exchange_->barrier();   // no rank overwrites vectors a peer still reads from the previous run
        SpMV_->set_vector_constant(graphlily::val_t(float(1.0 / n)));  // rank0 = 1 / N (pagerank.h:81-82), built on the device
        if (fused_ || exchange_) {
            DeviceBuffer vec = SpMV_->vector_buf, res = SpMV_->results_buf;
            replay({2, key_of(SpMV_->device_matrix()), key_of(teleport), num_iterations, key_of(vec.ptr()), key_of(res.ptr())}, [&] {
                // (the scalar travels as a word of val_t: the float itself, or its unsigned / Q8.24 conversion)
                glb_spmv_epilogue_t ep = {1, graphlily::val_container(graphlily::val_t(teleport)), nullptr, 0.0f, 0};
                if (exchange_) {
                    std::vector<glb_spmv_epilogue_t> eps(num_iterations, ep);
                    SpMV_->iterate_exchange(*exchange_, vec, DeviceBuffer(), res, eps.data(), int(num_iterations));
                    return;
                }
                DeviceBuffer v = vec, r = res;
                for (uint32_t iter = 1; iter <= num_iterations; iter++) {
                    SpMV_->run_fused(v, DeviceBuffer(), r, &ep);
                    std::swap(v, r);
                }
            });
            if (num_iterations % 2) std::swap(SpMV_->vector_buf, SpMV_->results_buf);
        } else {
            eWiseAdd_->bind_in_buf(SpMV_->results_buf);
            eWiseAdd_->bind_out_buf(SpMV_->vector_buf);
            for (uint32_t iter = 1; iter <= num_iterations; iter++) {
                SpMV_->run();
                eWiseAdd_->run(n, graphlily::val_t(teleport));
            }
        }
        return SpMV_->send_vector_device_to_host();
    }

    // compute_reference_results (reference: app/pagerank.h:150-159): declared for the reference's callers, defined only by
    // the test adapter tests/cpp/ref_compat/reference_results.h (oracle/); the product has no CPU path.
    aligned_dense_float_vec_t compute_reference_results(float damping, uint32_t num_iterations);
};

}  // namespace app
}  // namespace graphlily

#endif  // GRAPHLILY_PAGERANK_H_
