// graphlily-b200: masked scalar assign into a dense vector.
//
// Same public surface as /root/reference/graphlily/module/assign_vector_dense_module.h:17-164;
// run() becomes glb_assign_dense.  kNoMask is an error: the reference prints and exits (:88-95).
#ifndef GRAPHLILY_ASSIGN_VECTOR_DENSE_MODULE_H_
#define GRAPHLILY_ASSIGN_VECTOR_DENSE_MODULE_H_

#include "graphlily/global.h"
#include "graphlily/module/base_module.h"

namespace graphlily {
namespace module {

template <typename vector_data_t>
class AssignVectorDenseModule : public BaseModule {
private:
    graphlily::MaskType mask_type_ = graphlily::kNoMask;
    using aligned_dense_vec_t = std::vector<vector_data_t, aligned_allocator<vector_data_t>>;
    aligned_dense_vec_t mask_, inout_;

public:
    // Device buffers
    DeviceBuffer mask_buf;
    DeviceBuffer inout_buf;

    AssignVectorDenseModule() : BaseModule("overlay") {}

    void set_mask_type(graphlily::MaskType mask_type) {
        if (mask_type == graphlily::kNoMask) {
            std::cerr << "Please set the mask type" << std::endl;
            exit(EXIT_FAILURE);
        }
        mask_type_ = mask_type;
    }

    void send_mask_host_to_device(aligned_dense_vec_t &mask) {
        mask_ = mask;
        mask_buf = upload(mask_);
    }
    void send_inout_host_to_device(aligned_dense_vec_t &inout) {
        inout_ = inout;
        inout_buf = upload(inout_);
    }
    void bind_mask_buf(DeviceBuffer src_buf) { mask_buf = src_buf; }
    void bind_inout_buf(DeviceBuffer src_buf) { inout_buf = src_buf; }

    void run(uint32_t len, vector_data_t val) {
        using VT = graphlily::val_traits<vector_data_t>;
        if (VT::id == GLB_VAL_F32) GLB_CHECK(glb_assign_dense(ctx(), mask_buf.f32(), inout_buf.f32(), len, float(val), mask_type_));
        else GLB_CHECK(glb_assign_dense_vt(ctx(), VT::id, mask_buf.ptr(), inout_buf.ptr(), len, VT::bits(val), mask_type_));
        end_run();
    }

    aligned_dense_vec_t send_mask_device_to_host() {
        download(mask_, mask_buf, mask_buf.bytes() / sizeof(vector_data_t));
        return mask_;
    }
    aligned_dense_vec_t send_inout_device_to_host() {
        download(inout_, inout_buf, inout_buf.bytes() / sizeof(vector_data_t));
        return inout_;
    }

    // compute_reference_results (reference: assign_vector_dense_module.h:223-246): declared for the reference's callers, defined only by
    // the test adapter tests/cpp/ref_compat/reference_results.h (oracle/); the product has no CPU path.
    void compute_reference_results(graphlily::aligned_dense_float_vec_t &mask, graphlily::aligned_dense_float_vec_t &inout,
                                   uint32_t len, float val);
    graphlily::MaskType mask_type() const { return mask_type_; }
};

}  // namespace module
}  // namespace graphlily

#endif  // GRAPHLILY_ASSIGN_VECTOR_DENSE_MODULE_H_
