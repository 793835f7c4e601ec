// graphlily-b200: what the OpenCL runtime objects of the reference become.
//
//   cl::Device + cl::Context + cl::CommandQueue  ->  graphlily::Runtime   (one CUDA device + one stream)
//   cl::Buffer                                   ->  graphlily::DeviceBuffer (shared-ownership handle)
//   OCL_CHECK (xcl2.hpp:40-46: print + exit)     ->  graphlily::check
//
// Reference: /root/reference/graphlily/module/base_module.h:14-21,106-133 and the public cl::Buffer
// members of every module, which the apps alias across modules (app/bfs.h:113-116).
#ifndef GRAPHLILY_RUNTIME_H_
#define GRAPHLILY_RUNTIME_H_

#include <cstdio>
#include <cstdlib>
#include <memory>

#include "graphlily_b200.h"

namespace graphlily {

// The reference aborts the process on any runtime error; the drivers rely on that.
inline void check(int rc, const char *what) {
    if (rc != GLB_OK) {
        std::fprintf(stderr, "%s: glb error %d: %s\n", what, rc, glb_last_error());
        std::exit(EXIT_FAILURE);
    }
}
#define GLB_CHECK(call) ::graphlily::check((call), #call)

class Runtime {
public:
    explicit Runtime(int device = 0) { GLB_CHECK(glb_ctx_create(device, nullptr, &ctx_)); }
    ~Runtime() { glb_ctx_destroy(ctx_); }
    Runtime(const Runtime &) = delete;
    Runtime &operator=(const Runtime &) = delete;
    glb_ctx_t ctx() const { return ctx_; }
    void finish() { GLB_CHECK(glb_ctx_sync(ctx_)); }  // command_queue_.finish()
    // Device chosen by GRAPHLILY_B200_DEVICE (default 0); the xclbin path of the reference is ignored.
    static std::shared_ptr<Runtime> create_from_env() {
        const char *d = std::getenv("GRAPHLILY_B200_DEVICE");
        return std::make_shared<Runtime>(d ? std::atoi(d) : 0);
    }
private:
    glb_ctx_t ctx_ = nullptr;
};

class DeviceBuffer {
public:
    DeviceBuffer() {}
    DeviceBuffer(const std::shared_ptr<Runtime> &rt, size_t bytes) : block_(std::make_shared<Block>(rt, bytes)) {}
    void *ptr() const { return block_ ? block_->ptr : nullptr; }
    float *f32() const { return static_cast<float *>(ptr()); }
    glb_idx_val_t *sparse() const { return static_cast<glb_idx_val_t *>(ptr()); }
    size_t bytes() const { return block_ ? block_->bytes : 0; }
    bool valid() const { return bool(block_); }
    bool operator==(const DeviceBuffer &o) const { return block_ == o.block_; }
private:
    struct Block {
        std::shared_ptr<Runtime> rt;
        void *ptr = nullptr;
        size_t bytes = 0;
        Block(const std::shared_ptr<Runtime> &r, size_t b) : rt(r), bytes(b) { GLB_CHECK(glb_buffer_alloc(rt->ctx(), b, &ptr)); }
        ~Block() { glb_buffer_free(rt->ctx(), ptr); }
    };
    std::shared_ptr<Block> block_;
};

}  // namespace graphlily

#endif  // GRAPHLILY_RUNTIME_H_
