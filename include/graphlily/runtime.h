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
    // non-owning handle over device memory that belongs to something else (a vector of an Exchange)
    static DeviceBuffer view(const std::shared_ptr<Runtime> &rt, void *ptr, size_t bytes, int exchange_vector = -1) {
        DeviceBuffer b;
        b.block_ = std::make_shared<Block>(rt, ptr, bytes, exchange_vector);
        return b;
    }
    int exchange_vector() const { return block_ ? block_->exchange_vector : -1; }
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
        bool owned = true;
        int exchange_vector = -1;
        Block(const std::shared_ptr<Runtime> &r, size_t b) : rt(r), bytes(b) { GLB_CHECK(glb_buffer_alloc(rt->ctx(), b, &ptr)); }
        Block(const std::shared_ptr<Runtime> &r, void *p, size_t b, int xv) : rt(r), ptr(p), bytes(b), owned(false), exchange_vector(xv) {}
        ~Block() { if (owned) glb_buffer_free(rt->ctx(), ptr); }
    };
    std::shared_ptr<Block> block_;
};

// Vectors of a row-sharded run, mapped in every rank's process (one process per GPU): the C ABI's peer
// exchange over CUDA IPC (glb_xchg_create / _export / _connect) -- no framework in the data path.  The
// host exchanges the 64-byte handles of all ranks by any means it has (a pipe, MPI, a file).
class Exchange {
public:
    Exchange(const std::shared_ptr<Runtime> &rt, uint32_t n_floats, int n_vectors) : rt_(rt), n_(n_floats), n_vectors_(n_vectors) {
        GLB_CHECK(glb_xchg_create(rt_->ctx(), n_floats, n_vectors, &xc_));
    }
    // The same vectors behind an NVSwitch multicast object the library creates (glb_xchg_mc_open / _bind): a finished
    // slice leaves the GPU once and the switch replicates it.  Rank 0 passes fd_in = -1 and receives the object's file
    // descriptor in *fd_out, which the host hands to the other ranks' processes (SCM_RIGHTS over a Unix socket); they
    // pass their copy as fd_in.  After a host barrier every rank calls bind(), and after another one the exchange is usable.
    Exchange(const std::shared_ptr<Runtime> &rt, uint32_t n_floats, int n_vectors, int rank, int world, int fd_in, int *fd_out)
        : rt_(rt), n_(n_floats), n_vectors_(n_vectors), rank_(rank), world_(world) {
        GLB_CHECK(glb_xchg_mc_open(rt_->ctx(), n_floats, n_vectors, rank, world, fd_in, &xc_, fd_out));
    }
    void bind() { GLB_CHECK(glb_xchg_mc_bind(xc_)); }
    static bool multicast_supported(const std::shared_ptr<Runtime> &rt) {
        int s = 0;
        GLB_CHECK(glb_xchg_mc_supported(rt->ctx(), &s));
        return s != 0;
    }
    bool has_multicast() const { return glb_xchg_has_multicast(xc_) != 0; }
    ~Exchange() { glb_xchg_destroy(xc_); }
    Exchange(const Exchange &) = delete;
    Exchange &operator=(const Exchange &) = delete;
    void export_handle(void *handle64) { GLB_CHECK(glb_xchg_export(xc_, handle64)); }
    // handles: world x GLB_IPC_HANDLE_BYTES, in rank order
    void connect(int rank, int world, const void *handles) {
        GLB_CHECK(glb_xchg_connect(xc_, rank, world, handles));
        rank_ = rank;
        world_ = world;
    }
    DeviceBuffer buffer(int which) {
        float *p = nullptr;
        GLB_CHECK(glb_xchg_vector(xc_, which, &p));
        return DeviceBuffer::view(rt_, p, sizeof(float) * size_t(n_), which);
    }
    void barrier() { GLB_CHECK(glb_xchg_barrier(rt_->ctx(), xc_)); }
    void allgather(int which, size_t offset, size_t count) { GLB_CHECK(glb_xchg_allgather(rt_->ctx(), xc_, which, offset, count)); }
    glb_xchg_t handle() const { return xc_; }
    uint32_t size() const { return n_; }
    int vectors() const { return n_vectors_; }
    int rank() const { return rank_; }
    int world() const { return world_; }
private:
    std::shared_ptr<Runtime> rt_;
    glb_xchg_t xc_ = nullptr;
    uint32_t n_;
    int n_vectors_, rank_ = 0, world_ = 1;
};

}  // namespace graphlily

#endif  // GRAPHLILY_RUNTIME_H_
