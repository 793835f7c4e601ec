// graphlily-b200: assign by a sparse index list (BFS mode) / relax and emit the new frontier (SSSP mode).
//
// Same public surface as /root/reference/graphlily/module/assign_vector_sparse_module.h:17-210:
// run(val) is overlay mode 5 (glb_assign_sparse), run() is mode 6 (glb_assign_sparse_relax); the
// constructor flag selects which one is legal, misuse prints and exits as in the reference.
#ifndef GRAPHLILY_ASSIGN_VECTOR_SPARSE_MODULE_H_
#define GRAPHLILY_ASSIGN_VECTOR_SPARSE_MODULE_H_

#include "graphlily/global.h"
#include "graphlily/module/base_module.h"

namespace graphlily {
namespace module {

template <typename vector_data_t, typename sparse_vector_data_t>
class AssignVectorSparseModule : public BaseModule {
    static_assert(sizeof(sparse_vector_data_t) == sizeof(glb_idx_val_t), "idx_val_t must match the C ABI");
private:
    bool generate_new_frontier_;
    using aligned_dense_vec_t = std::vector<vector_data_t, aligned_allocator<vector_data_t>>;
    using aligned_sparse_vec_t = std::vector<sparse_vector_data_t, aligned_allocator<sparse_vector_data_t>>;
    aligned_sparse_vec_t mask_, new_frontier_;
    aligned_dense_vec_t inout_;

    void require_new_frontier(bool wanted) {
        if (generate_new_frontier_ != wanted) {
            std::cerr << "[ERROR]: this->generate_new_frontier_ should be " << (wanted ? "true" : "false") << std::endl;
            exit(EXIT_FAILURE);
        }
    }
    uint32_t count_of(const DeviceBuffer &list) {
        uint32_t n = 0;
        GLB_CHECK(glb_sparse_count(ctx(), list.sparse(), &n));
        return n;
    }

public:
    // Device buffers
    DeviceBuffer mask_buf;
    DeviceBuffer inout_buf;
    DeviceBuffer new_frontier_buf;

    explicit AssignVectorSparseModule(bool generate_new_frontier)
        : BaseModule("overlay"), generate_new_frontier_(generate_new_frontier) {}

    void send_mask_host_to_device(aligned_sparse_vec_t &mask) {
        mask_ = mask;
        mask_buf = upload(mask_);
        if (generate_new_frontier_) {  // assign_vector_sparse_module.h:232-247: same capacity as the mask
            new_frontier_.assign(mask_.size(), sparse_vector_data_t{0, vector_data_t(0)});
            new_frontier_buf = upload(new_frontier_);
        }
    }
    void send_inout_host_to_device(aligned_dense_vec_t &inout) {
        inout_ = inout;
        inout_buf = upload(inout_);
    }
    void bind_mask_buf(DeviceBuffer src_buf) { mask_buf = src_buf; }
    void bind_inout_buf(DeviceBuffer src_buf) { inout_buf = src_buf; }
    void bind_new_frontier_buf(DeviceBuffer src_buf) {
        require_new_frontier(true);
        new_frontier_buf = src_buf;
    }

    // BFS mode: inout[mask[i].index] = val
    void run(vector_data_t val) {
        require_new_frontier(false);
        using VT = graphlily::val_traits<vector_data_t>;
        if (VT::id == GLB_VAL_F32) GLB_CHECK(glb_assign_sparse(ctx(), mask_buf.sparse(), inout_buf.f32(), float(val)));
        else GLB_CHECK(glb_assign_sparse_vt(ctx(), VT::id, mask_buf.sparse(), inout_buf.ptr(), VT::bits(val)));
        end_run();
    }
    // SSSP mode: relax and emit the new frontier
    void run() {
        require_new_frontier(true);
        using VT = graphlily::val_traits<vector_data_t>;
        if (VT::id == GLB_VAL_F32)
            GLB_CHECK(glb_assign_sparse_relax(ctx(), mask_buf.sparse(), inout_buf.f32(), new_frontier_buf.sparse()));
        else
            GLB_CHECK(glb_assign_sparse_relax_vt(ctx(), VT::id, mask_buf.sparse(), inout_buf.ptr(), new_frontier_buf.sparse()));
        end_run();
    }

    aligned_sparse_vec_t send_mask_device_to_host() {
        download(mask_, mask_buf, size_t(count_of(mask_buf)) + 1);
        return mask_;
    }
    aligned_dense_vec_t send_inout_device_to_host() {
        download(inout_, inout_buf, inout_buf.bytes() / sizeof(vector_data_t));
        return inout_;
    }
    aligned_sparse_vec_t send_new_frontier_device_to_host() {
        require_new_frontier(true);
        download(new_frontier_, new_frontier_buf, size_t(count_of(new_frontier_buf)) + 1);
        return new_frontier_;
    }

    // compute_reference_results (reference: assign_vector_sparse_module.h:306-335): declared for the reference's callers, defined only by
    // the test adapter tests/cpp/ref_compat/reference_results.h (oracle/); the product has no CPU path.
    void compute_reference_results(graphlily::aligned_sparse_float_vec_t &mask, graphlily::aligned_dense_float_vec_t &inout,
                                   float val);
    void compute_reference_results(graphlily::aligned_sparse_float_vec_t &mask, graphlily::aligned_dense_float_vec_t &inout,
                                   graphlily::aligned_sparse_float_vec_t &new_frontier);
};

}  // namespace module
}  // namespace graphlily

#endif  // GRAPHLILY_ASSIGN_VECTOR_SPARSE_MODULE_H_
