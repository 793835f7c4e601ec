// TEST INFRASTRUCTURE: the few OpenCL / XRT names the reference's tests touch directly
// (/root/reference/tests/test_module_apply.cpp:236-258 builds a cl::Buffer over host memory by hand,
// binds it to a module and migrates it back), mapped onto the CUDA runtime of include/graphlily so that
// the file compiles unmodified.  cl::Buffer = a graphlily::DeviceBuffer shadowing the host vector
// (CL_MEM_USE_HOST_PTR: uploaded at creation, copied back by enqueueMigrateMemObjects).
#ifndef GLB_REF_COMPAT_OPENCL_H_
#define GLB_REF_COMPAT_OPENCL_H_
#include <memory>
#include <vector>

#include "graphlily/runtime.h"

typedef struct { unsigned flags; void *obj; void *param; } cl_mem_ext_ptr_t;
#define CL_MEM_EXT_PTR_XILINX 0
#define CL_MEM_USE_HOST_PTR 0
#define CL_MEM_READ_WRITE 0
#define CL_MIGRATE_MEM_OBJECT_HOST 0

namespace graphlily {
const unsigned HBM[34] = {};   // bank ids of the U280 (global.h:110-143): meaningless here
}  // namespace graphlily

namespace cl {
inline std::shared_ptr<graphlily::Runtime> &compat_runtime() {
    static std::shared_ptr<graphlily::Runtime> rt = graphlily::Runtime::create_from_env();
    return rt;
}
struct Device {};
struct Context {
    Context() {}
    Context(const Device &, void *, void *, void *) {}
};
struct Buffer {
    graphlily::DeviceBuffer dev;
    void *host = nullptr;
    size_t bytes = 0;
    Buffer() {}
    Buffer(const Context &, unsigned long, size_t n, cl_mem_ext_ptr_t *ext) : dev(compat_runtime(), n), host(ext->obj), bytes(n) {
        GLB_CHECK(glb_buffer_h2d(compat_runtime()->ctx(), dev.ptr(), host, bytes));
    }
    operator graphlily::DeviceBuffer() const { return dev; }
};
struct CommandQueue {
    CommandQueue(const Context &, const Device &) {}
    void enqueueMigrateMemObjects(const std::vector<Buffer> &bufs, int) {
        GLB_CHECK(glb_device_sync(compat_runtime()->ctx()));   // the module ran on its own stream
        for (const Buffer &b : bufs) GLB_CHECK(glb_buffer_d2h(compat_runtime()->ctx(), b.host, b.dev.ptr(), b.bytes));
    }
    void finish() { GLB_CHECK(glb_device_sync(compat_runtime()->ctx())); }
};
}  // namespace cl

namespace graphlily {
inline cl::Device find_device() { return cl::Device(); }
}  // namespace graphlily
#endif
