// graphlily-b200: eWiseAdd, out[i] = in[i] + val.
//
// Same public surface as /root/reference/graphlily/module/add_scalar_vector_dense_module.h:17-138;
// run() (:180-192) becomes glb_ewise_add.  The apps also use it with val = 0 as the device copy
// results -> vector (app/bfs.h:122).
#ifndef GRAPHLILY_ADD_SCALAR_VECTOR_DENSE_MODULE_H_
#define GRAPHLILY_ADD_SCALAR_VECTOR_DENSE_MODULE_H_

#include "graphlily/global.h"
#include "graphlily/module/base_module.h"

namespace graphlily {
namespace module {

template <typename vector_data_t>
class eWiseAddModule : public BaseModule {
private:
    using aligned_dense_vec_t = std::vector<vector_data_t, aligned_allocator<vector_data_t>>;
    aligned_dense_vec_t in_, out_;

public:
    // Device buffers
    DeviceBuffer in_buf;
    DeviceBuffer out_buf;

    eWiseAddModule() : BaseModule("overlay") {}

    void send_in_host_to_device(aligned_dense_vec_t &in) {
        in_ = in;
        in_buf = upload(in_);
    }
    void allocate_out_buf(uint32_t len) {
        out_.assign(len, vector_data_t(0));
        out_buf = DeviceBuffer(runtime_, sizeof(vector_data_t) * len);
    }
    void bind_in_buf(DeviceBuffer src_buf) { in_buf = src_buf; }
    void bind_out_buf(DeviceBuffer src_buf) { out_buf = src_buf; }

    void run(uint32_t len, vector_data_t val) {
        using VT = graphlily::val_traits<vector_data_t>;
        if (VT::id == GLB_VAL_F32) GLB_CHECK(glb_ewise_add(ctx(), in_buf.f32(), out_buf.f32(), len, float(val)));
        else GLB_CHECK(glb_ewise_add_vt(ctx(), VT::id, in_buf.ptr(), out_buf.ptr(), len, VT::bits(val)));
        end_run();
    }

    aligned_dense_vec_t send_out_device_to_host() {
        download(out_, out_buf, out_buf.bytes() / sizeof(vector_data_t));
        return out_;
    }

    // compute_reference_results (reference: add_scalar_vector_dense_module.h:195-204): declared for the reference's callers, defined only by
    // the test adapter tests/cpp/ref_compat/reference_results.h (oracle/); the product has no CPU path.
    graphlily::aligned_dense_float_vec_t compute_reference_results(graphlily::aligned_dense_float_vec_t const &in, uint32_t len,
                                                                   float val);
};

}  // namespace module
}  // namespace graphlily

#endif  // GRAPHLILY_ADD_SCALAR_VECTOR_DENSE_MODULE_H_
