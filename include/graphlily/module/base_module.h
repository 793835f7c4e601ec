// graphlily-b200: common base of the operator modules.
//
// Mirrors /root/reference/graphlily/module/base_module.h:10-133.  The OpenCL device / context /
// kernel / queue quartet becomes one shared graphlily::Runtime (CUDA device + stream); the
// `overlay` kernel name and the xclbin path are kept in the signatures and ignored: the kernels are
// in libgraphlily_b200.so.  set_unused_args / set_mode (the overlay's argument and mode plumbing,
// :88-101) have no counterpart: every overlay mode is its own C-ABI entry point.
#ifndef GRAPHLILY_BASE_MODULE_H_
#define GRAPHLILY_BASE_MODULE_H_

#include <cassert>
#include <memory>
#include <string>

#include "graphlily/global.h"
#include "graphlily/runtime.h"

namespace graphlily {
namespace module {

class BaseModule {
protected:
    std::string kernel_name_;
    std::string target_ = "hw";
    std::shared_ptr<Runtime> runtime_;
    // The reference's run() is setArg + enqueueTask + finish (spmv_module.h:471-475): callers -- its
    // benchmark drivers time `for (...) run();` with no synchronisation of their own -- rely on run()
    // returning after the kernel has.  run() keeps that; set_async_run(true) makes it return after
    // enqueueing (the apps' own loops do not go through run(): they record launch sequences).
    bool async_run_ = false;
    void end_run() {
        if (!async_run_) runtime_->finish();
    }

    glb_ctx_t ctx() const {
        if (!runtime_) {
            std::cerr << "[ERROR]: set_up_runtime was not called" << std::endl;
            exit(EXIT_FAILURE);
        }
        return runtime_->ctx();
    }
    template <typename vec_t>
    DeviceBuffer upload(const vec_t &host) {
        DeviceBuffer buf(runtime_, host.size() * sizeof(host[0]));
        GLB_CHECK(glb_buffer_h2d(ctx(), buf.ptr(), host.data(), host.size() * sizeof(host[0])));
        return buf;
    }
    // A vector that is constant but for one element, built on the device in one launch instead of
    // uploading n values.  A buffer of the right size is refilled in place, so the app loops see the
    // same addresses call after call and their recorded launch sequences stay valid.
    DeviceBuffer constant_on_device(const DeviceBuffer &reuse, uint32_t n, float value, bool has_index = false,
                                    uint32_t index = 0, float index_value = 0) {
        DeviceBuffer buf = (reuse.valid() && reuse.bytes() == sizeof(float) * size_t(n))
                               ? reuse : DeviceBuffer(runtime_, sizeof(float) * size_t(n));
        if (has_index) GLB_CHECK(glb_buffer_fill_one_f32(ctx(), buf.f32(), value, n, index, index_value));
        else GLB_CHECK(glb_buffer_fill_f32(ctx(), buf.f32(), value, n));
        return buf;
    }
    template <typename vec_t>
    void download(vec_t &host, const DeviceBuffer &buf, size_t count) {
        host.resize(count);
        GLB_CHECK(glb_buffer_d2h(ctx(), host.data(), buf.ptr(), count * sizeof(host[0])));
    }

public:
    explicit BaseModule(std::string kernel_name) : kernel_name_(kernel_name) {}
    virtual ~BaseModule() {}

    std::string get_kernel_name() { return kernel_name_; }

    // Share an existing runtime (what ModuleCollection does with its context and queues).
    void set_runtime(std::shared_ptr<Runtime> runtime) { runtime_ = runtime; }
    std::shared_ptr<Runtime> get_runtime() { return runtime_; }
    void set_async_run(bool on) { async_run_ = on; }

    void set_target(std::string target) {
        assert(target == "sw_emu" || target == "hw_emu" || target == "hw");
        target_ = target;
    }

    // Copy the contents of a buffer into another buffer without going through the host.
    void copy_buffer_device_to_device(DeviceBuffer src, DeviceBuffer dst, size_t bytes) {
        GLB_CHECK(glb_buffer_d2d(ctx(), dst.ptr(), src.ptr(), bytes));
        runtime_->finish();
    }

    // base_module.h:106-133: the path named a bitstream; here it is ignored.
    void set_up_runtime(std::string /*xclbin_file_path*/) {
        if (!runtime_) runtime_ = Runtime::create_from_env();
    }
};

}  // namespace module
}  // namespace graphlily

#endif  // GRAPHLILY_BASE_MODULE_H_
