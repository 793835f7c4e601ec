// graphlily-b200: single-source shortest paths over the min-plus semiring (pull, push, pull_push).
//
// Same surface and results as /root/reference/graphlily/app/sssp.h:68-253: _preprocess (:16-62)
// sets every weight to 1 and gives every row a zero-weight diagonal; dist0 = TropicalSemiring.zero
// (255), dist[source] = 0; one iteration is dist <- A min.+ dist.  Unreached vertices stay exactly
// 255 because min(255, 1 + 255) = 255.  The push -> pull switch copies the distance vector on the
// device (:226-233 does the same with copy_buffer_device_to_device).
#ifndef GRAPHLILY_SSSP_H_
#define GRAPHLILY_SSSP_H_

#include <utility>

#include "graphlily/app/module_collection.h"
#include "graphlily/io/data_formatter.h"
#include "graphlily/io/data_loader.h"
#include "graphlily/module/add_scalar_vector_dense_module.h"
#include "graphlily/module/assign_vector_sparse_module.h"
#include "graphlily/module/spmspv_module.h"
#include "graphlily/module/spmv_module.h"

namespace graphlily {
namespace app {

namespace detail {

// sssp.h:16-62 in one O(nnz) pass with the same output as the reference's in-place insertions,
// including its quirk: while walking row r it reads the row END from the not yet shifted indptr
// (sssp.h:31-32), so after k insertions it scans only the first len - k entries of the row, treats
// a row with len == k as empty, and inserts before the last scanned entry when no larger column
// was seen (sssp.h:49-52).
inline void sssp_preprocess(graphlily::io::CSRMatrix<float> &m) {
    const uint32_t num_rows = uint32_t(m.adj_indptr.size() - 1);
    std::vector<float> data;
    std::vector<uint32_t> indices, indptr(size_t(num_rows) + 1, 0);
    data.reserve(m.adj_data.size() + num_rows);
    indices.reserve(m.adj_indices.size() + num_rows);
    uint64_t k = 0;  // insertions so far
    for (uint32_t r = 0; r < num_rows; r++) {
        const uint64_t s = m.adj_indptr[r], e = m.adj_indptr[r + 1], len = e - s;
        int64_t insert_at = -1, zero_at = -1;
        if (len == k) {
            insert_at = int64_t(s);
        } else if (len > k) {
            const uint64_t scan_end = e - k;
            for (uint64_t i = s; i < scan_end; i++) {
                const uint32_t c = m.adj_indices[i];
                if (c == r) { zero_at = int64_t(i); break; }
                if (c > r || i == scan_end - 1) { insert_at = int64_t(i); break; }
            }
        }
        if (len == 0 && insert_at >= 0) {
            indices.push_back(r);
            data.push_back(0.0f);
        }
        for (uint64_t i = s; i < e; i++) {
            if (insert_at == int64_t(i)) {
                indices.push_back(r);
                data.push_back(0.0f);
            }
            indices.push_back(m.adj_indices[i]);
            data.push_back(zero_at == int64_t(i) ? 0.0f : 1.0f);
        }
        if (insert_at >= 0) k++;
        indptr[r + 1] = uint32_t(indices.size());
    }
    m.adj_data.swap(data);
    m.adj_indices.swap(indices);
    m.adj_indptr.swap(indptr);
}

}  // namespace detail

class SSSP : public app::ModuleCollection {
private:
    module::SpMVModule<graphlily::val_t, graphlily::val_t> *SpMV_;
    module::SpMSpVModule<graphlily::val_t, graphlily::val_t, graphlily::idx_val_t> *SpMSpV_;
    module::AssignVectorSparseModule<graphlily::val_t, graphlily::idx_val_t> *SparseAssign_;
    module::eWiseAddModule<graphlily::val_t> *eWiseAdd_;
    uint32_t matrix_num_rows_ = 0, matrix_num_cols_ = 0;
    uint32_t num_channels_, spmv_out_buf_len_, spmspv_out_buf_len_, vec_buf_len_;
    graphlily::SemiringType semiring_ = graphlily::TropicalSemiring;
    bool fused_ = true;
    uint32_t push_iterations_ = 0;
    bool push_iterations_device_ = false;
    DeviceBuffer frontier2_buf_;   // second frontier list of the fused push levels
    using aligned_dense_vec_t = graphlily::aligned_dense_vec_t;
    using aligned_sparse_vec_t = graphlily::aligned_sparse_vec_t;

    void pull_loop(uint32_t first_iter, uint32_t num_iterations) {
        if (fused_ || exchange_) {
            DeviceBuffer vec = SpMV_->vector_buf, res = SpMV_->results_buf;
            const uint32_t count = num_iterations >= first_iter ? num_iterations - first_iter + 1 : 0;
            replay({3, key_of(SpMV_->device_matrix()), count, key_of(vec.ptr()), key_of(res.ptr())}, [&] {
                if (exchange_) {
                    SpMV_->iterate_exchange(*exchange_, vec, DeviceBuffer(), res, nullptr, int(count));
                    return;
                }
                DeviceBuffer v = vec, r = res;
                for (uint32_t k = 0; k < count; k++) {
                    SpMV_->run_fused(v, DeviceBuffer(), r, nullptr);
                    std::swap(v, r);
                }
            });
            if (count % 2) std::swap(SpMV_->vector_buf, SpMV_->results_buf);
        } else {
            eWiseAdd_->bind_in_buf(SpMV_->results_buf);
            eWiseAdd_->bind_out_buf(SpMV_->vector_buf);
            for (uint32_t iter = first_iter; iter <= num_iterations; iter++) {
                SpMV_->run();
                eWiseAdd_->run(matrix_num_rows_, 0);
            }
        }
    }

    void push_setup(uint32_t source) {
        if (exchange_) {
            exchange_->barrier();
            SpMV_->home_buffers();
        }
        SpMSpV_->home_lists();
        SpMSpV_->set_vector_single(source, 0);                  // the source frontier (sssp.h:169-171), built on the device
        SpMSpV_->set_mask_constant(semiring_.zero, source, 0);  // distance (sssp.h:172-176), built on the device
        SparseAssign_->bind_mask_buf(SpMSpV_->results_buf);
        SparseAssign_->bind_inout_buf(SpMSpV_->mask_buf);
        SparseAssign_->bind_new_frontier_buf(SpMSpV_->vector_buf);
    }

    // The two lists the frontier alternates between in the fused push levels (the kernel reads the old
    // frontier while it appends the new one, so they cannot be one buffer as in sssp.h:185-187).
    void frontier_lists(DeviceBuffer lists[2]) {
        const size_t bytes = sizeof(graphlily::idx_val_t) * (size_t(matrix_num_cols_) + 1);
        if (!frontier2_buf_.valid() || frontier2_buf_.bytes() != bytes) frontier2_buf_ = DeviceBuffer(runtime_, bytes);
        lists[0] = SpMSpV_->vector_buf;
        lists[1] = frontier2_buf_;
    }
    // One push level in ONE launch: SpMSpV lists[(level - 1) & 1] -> results, with the relax of the distance
    // vector and the new frontier -> lists[level & 1] (sssp.h:178-190) fused into the kernel.
    void push_level_fused(const DeviceBuffer lists[2], uint32_t level, const glb_spmspv_next_t *next) {
        glb_spmspv_epilogue_t ep = {GLB_SPMSPV_EP_RELAX, SpMSpV_->mask_buf.f32(), 0.0f, lists[level & 1].sparse()};
        SpMSpV_->run_fused(lists[(level - 1) & 1], SpMSpV_->results_buf, &ep, next);
    }

    // pull_push as ONE recorded sequence with the direction decided on the device (sssp.h:210-214); the level
    // that stops pushing copies the distance vector into the input of the first pull level (sssp.h:219-222).
    aligned_dense_vec_t pull_push_device(uint32_t source, uint32_t num_iterations, float threshold) {
        const uint32_t n = matrix_num_rows_;
        push_setup(source);
        DeviceBuffer dist = SpMSpV_->mask_buf, lists[2], dense[2];
        frontier_lists(lists);
        if (!SpMV_->vector_buf.valid() || SpMV_->vector_buf.bytes() != sizeof(graphlily::val_t) * size_t(n))
            SpMV_->vector_buf = DeviceBuffer(runtime_, sizeof(graphlily::val_t) * size_t(n));
        SpMV_->home_buffers();
        dense[0] = SpMV_->vector_buf;
        dense[1] = SpMV_->results_buf;
        SpMSpV_->reset_push_levels();
        replay({5, key_of(SpMV_->device_matrix()), num_iterations, key_of(threshold), key_of(lists[0].ptr()), key_of(lists[1].ptr()),
                key_of(SpMSpV_->results_buf.ptr()), key_of(dist.ptr()), key_of(dense[0].ptr()), key_of(dense[1].ptr())}, [&] {
            std::vector<uint64_t> cond(num_iterations + 2, 0);
            for (uint32_t level = 2; level <= num_iterations; level++) cond[level] = cond_create();
            auto push_level = [&](uint32_t level) {
                glb_spmspv_next_t next = {int(level + 1 >= num_iterations), threshold, n, cond[level + 1], GLB_SPMSPV_DENSE_COPY,
                                          dense[(level + 1) & 1].f32(), dist.f32(), n};
                push_level_fused(lists, level, level < num_iterations ? &next : nullptr);
            };
            push_level(1);
            for (uint32_t level = 2; level <= num_iterations; level++)
                branch(cond[level], [&] { push_level(level); },
                       [&] { SpMV_->run_fused(dense[level & 1], DeviceBuffer(), dense[(level + 1) & 1], nullptr); });
        });
        push_iterations_device_ = true;
        // the last level is always a pull level (sssp.h:214: iter < num_iterations): its output is the result
        SpMV_->vector_buf = dense[(num_iterations + 1) & 1];
        SpMV_->results_buf = dense[num_iterations & 1];
        return SpMV_->send_vector_device_to_host();
    }

public:
    SSSP(uint32_t num_channels, uint32_t spmv_out_buf_len, uint32_t spmspv_out_buf_len, uint32_t vec_buf_len)
        : num_channels_(num_channels), spmv_out_buf_len_(spmv_out_buf_len), spmspv_out_buf_len_(spmspv_out_buf_len),
          vec_buf_len_(vec_buf_len) {
        SpMV_ = new module::SpMVModule<graphlily::val_t, graphlily::val_t>(num_channels_, spmv_out_buf_len_, vec_buf_len_);
        SpMV_->set_semiring(semiring_);
        SpMV_->set_mask_type(graphlily::kNoMask);
        add_module(SpMV_);
        SpMSpV_ = new module::SpMSpVModule<graphlily::val_t, graphlily::val_t, graphlily::idx_val_t>(spmspv_out_buf_len_);
        SpMSpV_->set_semiring(semiring_);
        SpMSpV_->set_mask_type(graphlily::kNoMask);
        add_module(SpMSpV_);
        SparseAssign_ = new module::AssignVectorSparseModule<graphlily::val_t, graphlily::idx_val_t>(true);
        add_module(SparseAssign_);
        eWiseAdd_ = new module::eWiseAddModule<graphlily::val_t>();
        add_module(eWiseAdd_);
    }

    void set_fused(bool fused) { fused_ = fused; }
    uint32_t get_nnz() { return SpMV_->get_nnz(); }
    uint32_t get_num_rows() { return matrix_num_rows_; }
    uint32_t get_push_iterations() { return push_iterations_device_ ? SpMSpV_->push_levels() : push_iterations_; }

    void load_and_format_matrix(graphlily::io::CSRMatrix<float> csr_matrix, bool skip_empty_rows) {
        detail::sssp_preprocess(csr_matrix);
        graphlily::io::util_round_csr_matrix_dim(csr_matrix, num_channels_ * graphlily::pack_size,
                                                 num_channels_ * graphlily::pack_size);
        graphlily::io::CSCMatrix<float> csc_matrix = graphlily::io::csr2csc(csr_matrix);
        SpMV_->load_and_format_matrix(csr_matrix, skip_empty_rows);
        SpMSpV_->load_and_format_matrix(csc_matrix);
        matrix_num_rows_ = SpMV_->get_num_rows();
        matrix_num_cols_ = SpMV_->get_num_cols();
        assert(matrix_num_rows_ == matrix_num_cols_);
    }
    void load_and_format_matrix(std::string csr_float_npz_path, bool skip_empty_rows) {
        load_and_format_matrix(graphlily::io::load_csr_matrix_from_float_npz(csr_float_npz_path), skip_empty_rows);
    }

    void send_matrix_host_to_device() {
        drop_recorded_sequences();
        if (world_ > 1) make_cuts(SpMV_->host_matrix().adj_indptr, matrix_num_rows_);
        SpMV_->send_matrix_host_to_device(row_begin(), row_end(matrix_num_rows_));
        if (exchange_) {
            assert(exchange_->size() == matrix_num_rows_ && exchange_->vectors() >= 2);
            SpMV_->vector_buf = exchange_->buffer(0);
            SpMV_->results_buf = exchange_->buffer(1);
        }
        SpMSpV_->send_matrix_host_to_device(row_begin(), row_end(matrix_num_rows_));
    }

    aligned_dense_vec_t pull(uint32_t source, uint32_t num_iterations) {
        if (exchange_) exchange_->barrier();
        SpMV_->set_vector_constant(semiring_.zero, source, 0);  // sssp.h:153-156
        pull_loop(1, num_iterations);
        return SpMV_->send_vector_device_to_host();
    }

    aligned_dense_vec_t push(uint32_t source, uint32_t num_iterations) {
        push_setup(source);
        if (!fused_ || world_ > 1 || !SpMSpV_->fused_levels_available()) {   // (sharded: the frontier exchange sits between the SpMSpV and the relax; the fused levels compute in fp32)
            for (uint32_t iter = 1; iter <= num_iterations; iter++) {
                SpMSpV_->run();
                exchange_frontier(SpMV_, SpMSpV_->results_buf, semiring_.zero, matrix_num_rows_);
                SparseAssign_->run();
            }
            return SpMSpV_->send_mask_device_to_host();
        }
        DeviceBuffer lists[2];
        frontier_lists(lists);
        replay({4, key_of(SpMV_->device_matrix()), num_iterations, key_of(lists[0].ptr()), key_of(lists[1].ptr()),
                key_of(SpMSpV_->results_buf.ptr()), key_of(SpMSpV_->mask_buf.ptr())}, [&] {
            for (uint32_t level = 1; level <= num_iterations; level++) push_level_fused(lists, level, nullptr);
        });
        return SpMSpV_->send_mask_device_to_host();
    }

    aligned_dense_vec_t pull_push(uint32_t source, uint32_t num_iterations, float threshold = 0.05) {
        if (fused_ && use_graphs_ && num_iterations >= 2 && world_ == 1 && SpMSpV_->fused_levels_available())
            return pull_push_device(source, num_iterations, threshold);
        push_iterations_device_ = false;
        const uint32_t n = matrix_num_rows_;
        push_setup(source);
        uint32_t iter = 1;
        uint32_t vector_nnz;
        do {
            SpMSpV_->run();
            exchange_frontier(SpMV_, SpMSpV_->results_buf, semiring_.zero, n);
            SparseAssign_->run();
            vector_nnz = SpMSpV_->get_results_nnz();
            iter++;
        } while (iter < num_iterations && (float(vector_nnz) / n < threshold));
        push_iterations_ = iter - 1;
        // switch from push to pull: the distance vector becomes the SpMV input (device copy)
        SpMV_->home_buffers();
        if (!SpMV_->vector_buf.valid() || SpMV_->vector_buf.bytes() < sizeof(graphlily::val_t) * n)
            SpMV_->vector_buf = DeviceBuffer(runtime_, sizeof(graphlily::val_t) * n);
        SpMV_->copy_buffer_device_to_device(SpMSpV_->mask_buf, SpMV_->vector_buf, sizeof(graphlily::val_t) * n);
        pull_loop(iter, num_iterations);
        return SpMV_->send_vector_device_to_host();
    }

    // compute_reference_results (reference: app/sssp.h:245-253): declared for the reference's callers, defined only by
    // the test adapter tests/cpp/ref_compat/reference_results.h (oracle/); the product has no CPU path.
    aligned_dense_float_vec_t compute_reference_results(uint32_t source, uint32_t num_iterations);
};

}  // namespace app
}  // namespace graphlily

#endif  // GRAPHLILY_SSSP_H_
