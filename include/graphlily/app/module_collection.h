// graphlily-b200: a set of modules sharing one runtime.
//
// Mirrors /root/reference/graphlily/app/module_collection.h:15-114: owns the modules, and
// set_up_runtime gives every module the same device context (there: one cl::Context with a kernel
// and an out-of-order queue per module; here: one CUDA device + one stream, so launches of
// different modules are ordered without the finish() calls the reference needs).
#ifndef GRAPHLILY_MODULE_COLLECTION_H_
#define GRAPHLILY_MODULE_COLLECTION_H_

#include <algorithm>
#include <cassert>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "graphlily/global.h"
#include "graphlily/module/base_module.h"

namespace graphlily {
namespace app {

using namespace module;

class ModuleCollection {
protected:
    std::vector<BaseModule *> modules_;
    uint32_t num_modules_ = 0;
    std::vector<std::string> kernel_names_;
    std::string target_ = "hw";
    std::shared_ptr<Runtime> runtime_;

    // Row-range sharding over the ranks of one box (one process per GPU): this process owns the
    // rows [cuts_[rank], cuts_[rank + 1]) of the CSR, cut where the nnz prefix crosses r / world (32-row
    // aligned); the slices of every iteration's result meet over the exchange.  The push direction shards the CSC by
    // the same row ranges and exchanges its frontier as a dense vector (exchange_frontier).
    int rank_ = 0, world_ = 1;
    Exchange *exchange_ = nullptr;
    std::vector<uint32_t> cuts_;
    void make_cuts(const std::vector<uint32_t> &indptr, uint32_t n) {
        cuts_.assign(size_t(world_) + 1, 0);
        const uint64_t nnz = indptr[n];
        for (int r = 1; r < world_; r++) {
            const uint64_t want = nnz * uint64_t(r) / uint64_t(world_);
            const uint32_t row = uint32_t(std::lower_bound(indptr.begin(), indptr.begin() + n + 1, uint32_t(want)) - indptr.begin());
            cuts_[r] = std::max(cuts_[r - 1], std::min(n, row) / 32 * 32);
        }
        cuts_[world_] = n;
    }
    uint32_t row_begin() const { return cuts_.empty() ? 0 : cuts_[rank_]; }
    uint32_t row_end(uint32_t n) const { return cuts_.empty() ? n : cuts_[rank_ + 1]; }

    // Row-sharded push: every rank's SpMSpV listed only the rows it owns.  The lists meet as ONE dense
    // vector -- own slice reset and scattered, slices exchanged, the full vector listed again on every rank into the
    // same list buffer -- so what follows (the sparse assign on the replicated distance vector, the next SpMSpV) sees
    // the whole frontier.  The dense vector ALTERNATES between the SpMV module's two exchange vectors (idle while the
    // push direction runs): the exchange's signal / wait says "every rank has written its slice of step k", not "every
    // rank has finished reading step k", so a fast rank's slice of step k + 1 must not land in the vector a slow rank
    // is still listing.  After the call the fresh frontier (dense) is spmv->vector_buf.
    template <typename SpMVT>
    void exchange_frontier(SpMVT *spmv, const DeviceBuffer &list, float zero, uint32_t n) {
        if (world_ == 1) return;
        DeviceBuffer dense = spmv->results_buf;
        assert(dense.exchange_vector() >= 0 && "the SpMV module's vectors must live in the exchange");
        GLB_CHECK(glb_sparse_to_dense_rows(runtime_->ctx(), list.sparse(), dense.f32(), row_begin(), row_end(n), zero));
        exchange_->allgather(dense.exchange_vector(), row_begin(), row_end(n) - row_begin());
        GLB_CHECK(glb_dense_to_sparse(runtime_->ctx(), dense.f32(), n, zero, list.sparse()));
        std::swap(spmv->vector_buf, spmv->results_buf);
    }

    // Launch replay.  The iteration loop of an app is a fixed launch sequence (same buffers, same
    // per-iteration scalars): it is recorded once per key as a CUDA graph (glb_graph_begin / _end)
    // and replayed with one glb_graph_launch afterwards -- a 7-level BFS is 21 launches of 5-50 us
    // kernels that one host thread cannot enqueue fast enough one by one.
    bool use_graphs_ = true;
    std::vector<std::pair<std::vector<uint64_t>, glb_graph_t>> graphs_;
    static constexpr size_t kGraphCache = 8;

    void replay(const std::vector<uint64_t> &key, const std::function<void()> &launches) {
        if (!use_graphs_) { launches(); return; }
        for (auto &kv : graphs_)
            if (kv.first == key) { GLB_CHECK(glb_graph_launch(runtime_->ctx(), kv.second)); return; }
        if (graphs_.size() >= kGraphCache) { glb_graph_destroy(graphs_.front().second); graphs_.erase(graphs_.begin()); }
        glb_graph_t g = nullptr;
        GLB_CHECK(glb_graph_begin(runtime_->ctx()));
        launches();
        GLB_CHECK(glb_graph_end(runtime_->ctx(), &g));
        graphs_.emplace_back(key, g);
        GLB_CHECK(glb_graph_launch(runtime_->ctx(), g));
    }
    // a new matrix went to the device: sequences recorded with the old one are void
    void drop_recorded_sequences() {
        for (auto &kv : graphs_) glb_graph_destroy(kv.second);
        graphs_.clear();
    }
    // A branch of the recorded sequence decided on the device (glb_graph_cond_create / glb_graph_branch_*)
    uint64_t cond_create() {
        uint64_t c = 0;
        GLB_CHECK(glb_graph_cond_create(runtime_->ctx(), &c));
        return c;
    }
    void branch(uint64_t cond, const std::function<void()> &if_arm, const std::function<void()> &else_arm) {
        GLB_CHECK(glb_graph_branch_begin(runtime_->ctx(), cond));
        if_arm();
        GLB_CHECK(glb_graph_branch_else(runtime_->ctx()));
        else_arm();
        GLB_CHECK(glb_graph_branch_end(runtime_->ctx()));
    }
    static uint64_t key_of(const void *p) { return uint64_t(reinterpret_cast<uintptr_t>(p)); }
    static uint64_t key_of(float v) { uint32_t b; memcpy(&b, &v, sizeof(b)); return b; }

public:
    ModuleCollection() {}
    ModuleCollection(const ModuleCollection &) = delete;
    ModuleCollection &operator=(const ModuleCollection &) = delete;
    virtual ~ModuleCollection() {
        for (auto &kv : graphs_) glb_graph_destroy(kv.second);
        for (size_t i = 0; i < num_modules_; i++) delete modules_[i];
    }
    void set_use_graphs(bool on) { use_graphs_ = on; }
    // Call before send_matrix_host_to_device.  `exchange`: connected, >= 3 vectors of the padded dimension
    // (0 / 1: the SpMV vector / results, 2: its mask); the collection does not own it.
    void set_sharding(int rank, int world, Exchange *exchange) {
        assert(rank >= 0 && rank < world && (world == 1 || exchange != nullptr));
        rank_ = rank;
        world_ = world;
        exchange_ = exchange;
    }

    void add_module(BaseModule *module) {
        modules_.push_back(module);
        kernel_names_.push_back(module->get_kernel_name());
        num_modules_++;
    }

    void set_target(std::string target) {
        assert(target == "sw_emu" || target == "hw_emu" || target == "hw");
        target_ = target;
    }

    // The path named the bitstream in the reference; ignored here.
    void set_up_runtime(std::string /*xclbin_file_path*/) {
        if (!runtime_) runtime_ = Runtime::create_from_env();
        for (size_t i = 0; i < num_modules_; i++) modules_[i]->set_runtime(runtime_);
    }
    void set_runtime(std::shared_ptr<Runtime> runtime) {
        runtime_ = runtime;
        for (size_t i = 0; i < num_modules_; i++) modules_[i]->set_runtime(runtime_);
    }
    void finish() { runtime_->finish(); }
};

}  // namespace app
}  // namespace graphlily

#endif  // GRAPHLILY_MODULE_COLLECTION_H_
