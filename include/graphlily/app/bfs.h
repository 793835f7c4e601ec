// graphlily-b200: level-synchronous BFS (pull, push, direction-optimising pull_push).
//
// Same class and method surface as /root/reference/graphlily/app/bfs.h:20-360 and the same
// results: distance[source] = 1, a vertex reached in iteration k gets k + 1, unreached stays 0
// (:108-110,123); matrix values forced to 1, dimensions padded to 128 (:84-97).  Differences in
// HOW, not in what:
//   * with fused() on (default) one pull level is ONE launch: the eWiseAdd copy and the dense
//     assign of :117-124 ride in the SpMV write-back (glb_spmv_fused) and vector / results swap;
//     set_fused(false) replays the reference's three launches per level literally;
//   * the push loop swaps the two sparse lists instead of copying results -> vector (:147-151);
//   * the push -> pull switch densifies the frontier on the device (glb_sparse_to_dense) instead of
//     the host round trip of :195-201.
// compute_reference_results (:350-360) is not in the product; the tests link oracle/ for it.
#ifndef GRAPHLILY_BFS_H_
#define GRAPHLILY_BFS_H_

#include <utility>

#include "graphlily/app/module_collection.h"
#include "graphlily/io/data_formatter.h"
#include "graphlily/io/data_loader.h"
#include "graphlily/module/add_scalar_vector_dense_module.h"
#include "graphlily/module/assign_vector_dense_module.h"
#include "graphlily/module/assign_vector_sparse_module.h"
#include "graphlily/module/spmspv_module.h"
#include "graphlily/module/spmv_module.h"

namespace graphlily {
namespace app {

using graphlily::io::CSCMatrix;
using graphlily::io::CSRMatrix;

class BFS : public app::ModuleCollection {
private:
    module::SpMVModule<graphlily::val_t, graphlily::val_t> *SpMV_;
    module::AssignVectorDenseModule<graphlily::val_t> *DenseAssign_;
    module::SpMSpVModule<graphlily::val_t, graphlily::val_t, graphlily::idx_val_t> *SpMSpV_;
    module::AssignVectorSparseModule<graphlily::val_t, graphlily::idx_val_t> *SparseAssign_;
    module::eWiseAddModule<graphlily::val_t> *eWiseAdd_;
    uint32_t matrix_num_rows_ = 0, matrix_num_cols_ = 0;
    uint32_t num_channels_, spmv_out_buf_len_, spmspv_out_buf_len_, vec_buf_len_;
    graphlily::SemiringType semiring_ = graphlily::LogicalSemiring;
    bool fused_ = true;
    uint32_t push_iterations_ = 0;
    bool push_iterations_device_ = false;

    using aligned_dense_vec_t = graphlily::aligned_dense_vec_t;
    using aligned_sparse_vec_t = graphlily::aligned_sparse_vec_t;

    void pull_loop(uint32_t first_iter, uint32_t num_iterations) {
        const uint32_t n = matrix_num_rows_;
        if (fused_ || exchange_) {
            DeviceBuffer vec = SpMV_->vector_buf, res = SpMV_->results_buf, mask = SpMV_->mask_buf;
            replay({1, key_of(SpMV_->device_matrix()), first_iter, num_iterations, key_of(vec.ptr()), key_of(res.ptr()), key_of(mask.ptr())}, [&] {
                if (exchange_) {   // the whole loop in one call; the distance stays row-local until the end
                    std::vector<glb_spmv_epilogue_t> eps;
                    for (uint32_t iter = first_iter; iter <= num_iterations; iter++)
                        eps.push_back({0, 0.0f, mask.f32(), float(iter + 1), GLB_MASK_WRITE_TO_ONE});
                    SpMV_->iterate_exchange(*exchange_, vec, mask, res, eps.data(), int(eps.size()));
                    return;
                }
                DeviceBuffer v = vec, r = res;
                for (uint32_t iter = first_iter; iter <= num_iterations; iter++) {
                    glb_spmv_epilogue_t ep = {0, 0.0f, mask.f32(), graphlily::val_container(graphlily::val_t(float(iter + 1))), GLB_MASK_WRITE_TO_ONE};
                    SpMV_->run_fused(v, mask, r, &ep);
                    std::swap(v, r);
                }
            });
            if (num_iterations >= first_iter && (num_iterations - first_iter + 1) % 2) std::swap(SpMV_->vector_buf, SpMV_->results_buf);
        } else {
            DenseAssign_->bind_mask_buf(SpMV_->vector_buf);
            DenseAssign_->bind_inout_buf(SpMV_->mask_buf);
            eWiseAdd_->bind_in_buf(SpMV_->results_buf);
            eWiseAdd_->bind_out_buf(SpMV_->vector_buf);
            for (uint32_t iter = first_iter; iter <= num_iterations; iter++) {
                SpMV_->run();
                eWiseAdd_->run(n, 0);
                DenseAssign_->run(n, iter + 1);
            }
        }
    }

    void push_setup(uint32_t source) {
        if (exchange_) {
            exchange_->barrier();   // no rank overwrites vectors a peer still reads from the previous run
            SpMV_->home_buffers();
        }
        SpMSpV_->home_lists();
        SpMSpV_->set_vector_single(source, 1);     // one source vertex (bfs.h:131-135), built on the device
        SpMSpV_->set_mask_constant(0, source, 1);  // distance: 0 but distance[source] = 1, built on the device
        SparseAssign_->bind_inout_buf(SpMSpV_->mask_buf);
    }

    // One push level in ONE launch: SpMSpV lists[(level - 1) & 1] -> lists[level & 1] with the sparse assign
    // distance[row] = level + 1 (bfs.h:147-151) fused into the kernel.
    void push_level_fused(const DeviceBuffer lists[2], uint32_t level, const glb_spmspv_next_t *next) {
        glb_spmspv_epilogue_t ep = {GLB_SPMSPV_EP_ASSIGN, SpMSpV_->mask_buf.f32(), float(level + 1), nullptr};
        SpMSpV_->run_fused(lists[(level - 1) & 1], lists[level & 1], &ep, next);
    }

    // The two dense vectors the pull levels ping-pong through, zero-filled: the push level that stops pushing
    // scatters its frontier into one of them.
    void dense_pair(DeviceBuffer dense[2]) {
        const uint32_t n = matrix_num_rows_;
        if (!SpMV_->vector_buf.valid() || SpMV_->vector_buf.bytes() != sizeof(graphlily::val_t) * size_t(n))
            SpMV_->vector_buf = DeviceBuffer(runtime_, sizeof(graphlily::val_t) * size_t(n));
        SpMV_->home_buffers();
        dense[0] = SpMV_->vector_buf;
        dense[1] = SpMV_->results_buf;
        for (int i = 0; i < 2; i++) GLB_CHECK(glb_buffer_fill_f32(runtime_->ctx(), dense[i].f32(), semiring_.zero, n));
    }

    // pull_push as ONE recorded sequence with the direction decided on the device (bfs.h:186-190): every push
    // level is one launch that also sets the IF / ELSE condition of the next level; no host synchronisation
    // between the start vectors and the read-back of the result.
    aligned_dense_vec_t pull_push_device(uint32_t source, uint32_t num_iterations, float threshold) {
        const uint32_t n = matrix_num_rows_;
        push_setup(source);
        DeviceBuffer dist = SpMSpV_->mask_buf;
        SpMV_->bind_mask_buf(dist);
        DeviceBuffer dense[2], lists[2] = {SpMSpV_->vector_buf, SpMSpV_->results_buf};
        dense_pair(dense);
        SpMSpV_->reset_push_levels();
        replay({4, key_of(SpMV_->device_matrix()), num_iterations, key_of(threshold), key_of(lists[0].ptr()), key_of(lists[1].ptr()),
                key_of(dist.ptr()), key_of(dense[0].ptr()), key_of(dense[1].ptr())}, [&] {
            std::vector<uint64_t> cond(num_iterations + 2, 0);
            for (uint32_t level = 2; level <= num_iterations; level++) cond[level] = cond_create();
            auto push_level = [&](uint32_t level) {
                glb_spmspv_next_t next = {int(level + 1 >= num_iterations), threshold, n, cond[level + 1], GLB_SPMSPV_DENSE_SCATTER,
                                          dense[(level + 1) & 1].f32(), nullptr, n};
                push_level_fused(lists, level, level < num_iterations ? &next : nullptr);
            };
            auto pull_level = [&](uint32_t level) {
                glb_spmv_epilogue_t ep = {0, 0.0f, dist.f32(), float(level + 1), GLB_MASK_WRITE_TO_ONE};
                SpMV_->run_fused(dense[level & 1], dist, dense[(level + 1) & 1], &ep);
            };
            push_level(1);
            for (uint32_t level = 2; level <= num_iterations; level++)
                branch(cond[level], [&] { push_level(level); }, [&] { pull_level(level); });
        });
        push_iterations_device_ = true;
        return SpMSpV_->send_mask_device_to_host();
    }

    void push_step(uint32_t iter) {
        SpMSpV_->run();
        exchange_frontier(SpMV_, SpMSpV_->results_buf, semiring_.zero, matrix_num_rows_);  // (row-sharded runs)
        std::swap(SpMSpV_->vector_buf, SpMSpV_->results_buf);  // the new frontier is the next input
        SparseAssign_->bind_mask_buf(SpMSpV_->vector_buf);
        SparseAssign_->run(iter + 1);
    }

public:
    BFS(uint32_t num_channels, uint32_t spmv_out_buf_len, uint32_t spmspv_out_buf_len, uint32_t vec_buf_len)
        : num_channels_(num_channels), spmv_out_buf_len_(spmv_out_buf_len), spmspv_out_buf_len_(spmspv_out_buf_len),
          vec_buf_len_(vec_buf_len) {
        SpMV_ = new module::SpMVModule<graphlily::val_t, graphlily::val_t>(num_channels_, spmv_out_buf_len_, vec_buf_len_);
        SpMV_->set_semiring(semiring_);
        SpMV_->set_mask_type(graphlily::kMaskWriteToZero);
        add_module(SpMV_);
        DenseAssign_ = new module::AssignVectorDenseModule<graphlily::val_t>();
        DenseAssign_->set_mask_type(graphlily::kMaskWriteToOne);
        add_module(DenseAssign_);
        SpMSpV_ = new module::SpMSpVModule<graphlily::val_t, graphlily::val_t, graphlily::idx_val_t>(spmspv_out_buf_len_);
        SpMSpV_->set_semiring(semiring_);
        SpMSpV_->set_mask_type(graphlily::kMaskWriteToZero);
        add_module(SpMSpV_);
        SparseAssign_ = new module::AssignVectorSparseModule<graphlily::val_t, graphlily::idx_val_t>(false);
        add_module(SparseAssign_);
        eWiseAdd_ = new module::eWiseAddModule<graphlily::val_t>();
        add_module(eWiseAdd_);
    }

    void set_fused(bool fused) { fused_ = fused; }
    uint32_t get_nnz() { return SpMV_->get_nnz(); }
    uint32_t get_num_rows() { return matrix_num_rows_; }
    // levels the last pull_push spent pushing; with the decision on the device it is read back on demand
    uint32_t get_push_iterations() { return push_iterations_device_ ? SpMSpV_->push_levels() : push_iterations_; }

    void load_and_format_matrix(CSRMatrix<float> csr_matrix, bool skip_empty_rows) {
        graphlily::io::util_round_csr_matrix_dim(csr_matrix, num_channels_ * graphlily::pack_size,
                                                 num_channels_ * graphlily::pack_size);
        for (auto &x : csr_matrix.adj_data) x = 1;
        CSCMatrix<float> csc_matrix = graphlily::io::csr2csc(csr_matrix);
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
        if (exchange_) {   // vector / results / distance of the pull loop live in the exchange
            assert(exchange_->size() == matrix_num_rows_ && exchange_->vectors() >= 3);
            SpMV_->vector_buf = exchange_->buffer(0);
            SpMV_->results_buf = exchange_->buffer(1);
            SpMV_->mask_buf = exchange_->buffer(2);
        }
        SpMSpV_->send_matrix_host_to_device(row_begin(), row_end(matrix_num_rows_));
        if (exchange_) SpMSpV_->mask_buf = exchange_->buffer(2);   // the distance vector both directions update
    }

    aligned_dense_vec_t pull(uint32_t source, uint32_t num_iterations) {
        if (exchange_) exchange_->barrier();   // no rank overwrites vectors a peer still reads from the previous run
        // input = zero, input[source] = 1; distance = 0, distance[source] = 1 (bfs.h:108-112): built on
        // the device instead of uploaded
        SpMV_->set_vector_constant(semiring_.zero, source, 1);
        SpMV_->set_mask_constant(0, source, 1);
        pull_loop(1, num_iterations);
        if (exchange_)   // the distance was updated shard by shard: complete it on every rank
            exchange_->allgather(2, row_begin(), row_end(matrix_num_rows_) - row_begin());
        return SpMV_->send_mask_device_to_host();
    }

    aligned_dense_vec_t push(uint32_t source, uint32_t num_iterations) {
        push_setup(source);
        if (!fused_ || world_ > 1 || !SpMSpV_->fused_levels_available()) {   // (sharded: the frontier exchange sits between the SpMSpV and the assign; the fused levels compute in fp32)
            for (uint32_t iter = 1; iter <= num_iterations; iter++) push_step(iter);
            return SpMSpV_->send_mask_device_to_host();
        }
        DeviceBuffer lists[2] = {SpMSpV_->vector_buf, SpMSpV_->results_buf};
        replay({3, key_of(SpMV_->device_matrix()), num_iterations, key_of(lists[0].ptr()), key_of(lists[1].ptr()),
                key_of(SpMSpV_->mask_buf.ptr())}, [&] {
            for (uint32_t level = 1; level <= num_iterations; level++) push_level_fused(lists, level, nullptr);
        });
        SpMSpV_->vector_buf = lists[num_iterations & 1];
        SpMSpV_->results_buf = lists[(num_iterations + 1) & 1];
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
            push_step(iter);
            vector_nnz = SpMSpV_->get_vector_nnz();
            iter++;
        } while (iter < num_iterations && (float(vector_nnz) / n < threshold));
        push_iterations_ = iter - 1;
        // switch from push to pull: the last frontier becomes the dense SpMV input, on the device
        SpMV_->bind_mask_buf(SpMSpV_->mask_buf);
        if (world_ == 1) {   // (sharded: the frontier already sits in SpMV_->vector_buf, it travelled as a dense vector)
            SpMV_->home_buffers();
            if (!SpMV_->vector_buf.valid() || SpMV_->vector_buf.bytes() < sizeof(graphlily::val_t) * n)
                SpMV_->vector_buf = DeviceBuffer(runtime_, sizeof(graphlily::val_t) * n);
            GLB_CHECK(glb_sparse_to_dense(runtime_->ctx(), SpMSpV_->vector_buf.sparse(), SpMV_->vector_buf.f32(), n,
                                          graphlily::LogicalSemiring.zero));
        }
        pull_loop(iter, num_iterations);
        if (exchange_)   // the pull levels updated the distance shard by shard
            exchange_->allgather(SpMSpV_->mask_buf.exchange_vector(), row_begin(), row_end(n) - row_begin());
        return SpMSpV_->send_mask_device_to_host();  // the mask of SpMV on the host is not valid
    }

    // compute_reference_results (reference: app/bfs.h:350-360): declared for the reference's callers, defined only by
    // the test adapter tests/cpp/ref_compat/reference_results.h (oracle/); the product has no CPU path.
    aligned_dense_float_vec_t compute_reference_results(uint32_t source, uint32_t num_iterations);
};

}  // namespace app
}  // namespace graphlily

#endif  // GRAPHLILY_BFS_H_
