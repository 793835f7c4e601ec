// Runtime part of the C ABI: error text, context (device + stream), buffers, NCCL binding.
// Replaces what BaseModule::set_up_runtime / cl::Buffer / enqueueMigrateMemObjects do in the
// reference (/root/reference/graphlily/module/base_module.h:82-133).
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "glb_internal.h"

static thread_local char g_err[1024] = "";

void glb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int glb_version(void) { return GLB_VERSION; }

const char *glb_last_error(void) { return g_err; }

int glb_device_count(int *count) {
    GLB_REQUIRE(count, "count is NULL");
    *count = 0;
    GLB_CUDA(cudaGetDeviceCount(count));
    return GLB_OK;
}

int glb_ctx_create(int device, void *cuda_stream, glb_ctx_t *out) {
    GLB_REQUIRE(out, "out is NULL");
    *out = nullptr;
    int n = 0;
    GLB_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        glb_set_error("glb_ctx_create: device %d not available (%d CUDA devices); there is no CPU fallback", device, n);
        return GLB_ECUDA;
    }
    GLB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GLB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        glb_set_error("glb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                      prop.minor);
        return GLB_ECUDA;
    }
    glb_ctx_t ctx = new glb_ctx_s();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    // Vector buffers come from the device's stream-ordered pool and go back to it without a
    // device synchronisation: cudaMalloc / cudaFree cost milliseconds (measured 7-500 ms for a
    // 12 MB vector on a busy context) and the apps allocate start vectors on every call.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
        ctx->owns_stream = false;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            glb_set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
            delete ctx;
            return GLB_ECUDA;
        }
        ctx->owns_stream = true;
    }
    *out = ctx;
    return GLB_OK;
}

}  // extern "C"

static void ctx_free(glb_ctx_t ctx);
void glb_ctx_retain(glb_ctx_t ctx) { ctx->children++; }
void glb_ctx_release(glb_ctx_t ctx) {
    if (--ctx->children == 0 && ctx->destroyed) ctx_free(ctx);
}

extern "C" {

int glb_ctx_destroy(glb_ctx_t ctx) {
    if (!ctx || ctx->destroyed) return GLB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->destroyed = true;
    if (ctx->children == 0) ctx_free(ctx);
    return GLB_OK;
}

}  // extern "C"

static void ctx_free(glb_ctx_t ctx) {
    cudaSetDevice(ctx->device);
    glb_comm_destroy(ctx);
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->stream);
    for (cudaEvent_t e : ctx->stage_ev) if (e) cudaEventDestroy(e);
    for (auto &row : ctx->pipe_ev)
        for (cudaEvent_t e : row) if (e) cudaEventDestroy(e);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->branch_stream) cudaStreamDestroy(ctx->branch_stream);
    if (ctx->split_ev_head) cudaEventDestroy(ctx->split_ev_head);
    if (ctx->pipe_pushed) cudaEventDestroy(ctx->pipe_pushed);
    if (ctx->pusher_stream) cudaStreamDestroy(ctx->pusher_stream);
    if (ctx->pusher_ev_fork) cudaEventDestroy(ctx->pusher_ev_fork);
    if (ctx->pusher_ev_done) cudaEventDestroy(ctx->pusher_ev_done);
    for (int s = 0; s < GLB_MAX_SPLIT; ++s) {
        if (ctx->split_stream[s]) cudaStreamDestroy(ctx->split_stream[s]);
        if (ctx->split_ev_main[s]) cudaEventDestroy(ctx->split_ev_main[s]);
        if (ctx->split_ev_done[s]) cudaEventDestroy(ctx->split_ev_done[s]);
    }
    if (ctx->staging) cudaFreeHost(ctx->staging);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" {

int glb_ctx_sync(glb_ctx_t ctx) {
    GLB_REQUIRE(ctx, "ctx is NULL");
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GLB_OK;
}

int glb_device_sync(glb_ctx_t ctx) {
    GLB_REQUIRE(ctx, "ctx is NULL");
    GLB_CUDA(cudaSetDevice(ctx->device));
    GLB_CUDA(cudaDeviceSynchronize());
    return GLB_OK;
}

int glb_ctx_stream(glb_ctx_t ctx, void **cuda_stream) {
    GLB_REQUIRE(ctx && cuda_stream, "NULL argument");
    *cuda_stream = ctx->stream;
    return GLB_OK;
}

int glb_ctx_kernel_timing(glb_ctx_t ctx, int enable) {
    GLB_REQUIRE(ctx, "ctx is NULL");
    ctx->timing = enable != 0;
    return GLB_OK;
}

int glb_ctx_kernel_timing_read(glb_ctx_t ctx, double out[3]) {
    GLB_REQUIRE(ctx && out, "NULL argument");
    out[0] = out[1] = out[2] = 0.0;
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i + 2 < ctx->timing_events.size(); i += 3) {
        float a = 0.f, b = 0.f;
        GLB_CUDA(cudaEventElapsedTime(&a, ctx->timing_events[i], ctx->timing_events[i + 1]));
        GLB_CUDA(cudaEventElapsedTime(&b, ctx->timing_events[i + 1], ctx->timing_events[i + 2]));
        out[0] += a;
        out[1] += b;
        out[2] += 1.0;
    }
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    ctx->timing_events.clear();
    return GLB_OK;
}

// ------------------------------------------------------------------------ buffers
int glb_buffer_alloc(glb_ctx_t ctx, size_t bytes, void **dptr) {
    GLB_REQUIRE(ctx && dptr, "NULL argument");
    *dptr = nullptr;
    GLB_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMallocAsync(dptr, bytes ? bytes : 4, ctx->stream);
    if (e != cudaSuccess) {
        glb_set_error("cudaMallocAsync(%zu): %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? GLB_ENOMEM : GLB_ECUDA;
    }
    return GLB_OK;
}

int glb_buffer_free(glb_ctx_t ctx, void *dptr) {
    GLB_REQUIRE(ctx, "ctx is NULL");
    if (!dptr) return GLB_OK;
    GLB_CUDA(cudaSetDevice(ctx->device));
    GLB_CUDA(cudaFreeAsync(dptr, ctx->stream));  // stream-ordered: safe after the kernels already enqueued
    return GLB_OK;
}

int glb_buffer_h2d_async(glb_ctx_t ctx, void *dst_dev, const void *src_host, size_t bytes) {
    GLB_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_host)), "NULL argument");
    if (bytes) GLB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return GLB_OK;
}

int glb_buffer_d2h_async(glb_ctx_t ctx, void *dst_host, const void *src_dev, size_t bytes) {
    GLB_REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), "NULL argument");
    if (bytes) GLB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return GLB_OK;
}

// Pageable host memory moves at 2-3 GB/s through the driver's bounce path; large blocking copies
// are staged through a page-locked buffer owned by the context instead (memcpy + full-speed DMA),
// in two halves so the memcpy of one half overlaps the DMA of the other.
static bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

static int staging_reserve(glb_ctx_t ctx, size_t bytes) {
    if (ctx->staging_bytes >= bytes) return GLB_OK;
    if (ctx->staging) cudaFreeHost(ctx->staging);
    ctx->staging = nullptr;
    ctx->staging_bytes = 0;
    cudaError_t e = cudaMallocHost(&ctx->staging, bytes);
    if (e != cudaSuccess) { glb_set_error("cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)); return GLB_ENOMEM; }
    ctx->staging_bytes = bytes;
    return GLB_OK;
}

constexpr size_t kStageMin = 256 * 1024, kStagePiece = 4u << 20;

int glb_buffer_h2d(glb_ctx_t ctx, void *dst_dev, const void *src_host, size_t bytes) {
    GLB_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_host)), "NULL argument");
    if (bytes >= kStageMin && is_pageable(src_host)) {
        int rc = staging_reserve(ctx, 2 * kStagePiece);
        if (rc) return rc;
        if (!ctx->stage_ev[0]) for (auto &e : ctx->stage_ev) GLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        size_t off = 0;
        for (int k = 0; off < bytes; ++k, off += kStagePiece) {
            const size_t nb = bytes - off < kStagePiece ? bytes - off : kStagePiece;
            char *slot = static_cast<char *>(ctx->staging) + (k & 1) * kStagePiece;
            if (k >= 2) GLB_CUDA(cudaEventSynchronize(ctx->stage_ev[k & 1]));  // the DMA that last read this slot
            memcpy(slot, static_cast<const char *>(src_host) + off, nb);
            GLB_CUDA(cudaMemcpyAsync(static_cast<char *>(dst_dev) + off, slot, nb, cudaMemcpyHostToDevice, ctx->stream));
            GLB_CUDA(cudaEventRecord(ctx->stage_ev[k & 1], ctx->stream));
        }
    } else {
        int rc = glb_buffer_h2d_async(ctx, dst_dev, src_host, bytes);
        if (rc) return rc;
    }
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GLB_OK;
}

int glb_buffer_d2h(glb_ctx_t ctx, void *dst_host, const void *src_dev, size_t bytes) {
    GLB_REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), "NULL argument");
    if (bytes >= kStageMin && is_pageable(dst_host)) {
        int rc = staging_reserve(ctx, 2 * kStagePiece);
        if (rc) return rc;
        if (!ctx->stage_ev[0]) for (auto &e : ctx->stage_ev) GLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const size_t n_pieces = (bytes + kStagePiece - 1) / kStagePiece;
        auto issue = [&](size_t k) -> cudaError_t {
            const size_t off = k * kStagePiece, nb = bytes - off < kStagePiece ? bytes - off : kStagePiece;
            char *slot = static_cast<char *>(ctx->staging) + (k & 1) * kStagePiece;
            cudaError_t e = cudaMemcpyAsync(slot, static_cast<const char *>(src_dev) + off, nb, cudaMemcpyDeviceToHost, ctx->stream);
            return e != cudaSuccess ? e : cudaEventRecord(ctx->stage_ev[k & 1], ctx->stream);
        };
        GLB_CUDA(issue(0));
        for (size_t k = 0; k < n_pieces; ++k) {
            if (k + 1 < n_pieces) GLB_CUDA(issue(k + 1));
            GLB_CUDA(cudaEventSynchronize(ctx->stage_ev[k & 1]));
            const size_t off = k * kStagePiece, nb = bytes - off < kStagePiece ? bytes - off : kStagePiece;
            memcpy(static_cast<char *>(dst_host) + off, static_cast<char *>(ctx->staging) + (k & 1) * kStagePiece, nb);
        }
        return GLB_OK;
    }
    int rc = glb_buffer_d2h_async(ctx, dst_host, src_dev, bytes);
    if (rc) return rc;
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GLB_OK;
}

int glb_buffer_d2d(glb_ctx_t ctx, void *dst_dev, const void *src_dev, size_t bytes) {
    GLB_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_dev)), "NULL argument");
    if (bytes) GLB_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return GLB_OK;
}

int glb_host_alloc(size_t bytes, void **hptr) {
    GLB_REQUIRE(hptr, "hptr is NULL");
    *hptr = nullptr;
    cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 4);
    if (e != cudaSuccess) {
        glb_set_error("cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
        return GLB_ENOMEM;
    }
    return GLB_OK;
}

int glb_host_free(void *hptr) {
    if (hptr) GLB_CUDA(cudaFreeHost(hptr));
    return GLB_OK;
}

// ------------------------------------------------------------------------ NCCL (dlopen)
// libnccl is resolved at run time so the library loads on hosts without NCCL; in a torch
// process dlopen("libnccl.so.2") returns the copy torch already mapped.
typedef struct { char internal[GLB_NCCL_UNIQUE_ID_BYTES]; } nccl_uid_t;
typedef int (*pfn_get_uid)(nccl_uid_t *);
typedef int (*pfn_init_rank)(void **, int, nccl_uid_t, int);
typedef int (*pfn_destroy)(void *);
typedef int (*pfn_allgather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*pfn_errstr)(int);

static struct {
    bool tried = false;
    void *handle = nullptr;
    pfn_get_uid get_uid = nullptr;
    pfn_init_rank init_rank = nullptr;
    pfn_destroy destroy = nullptr;
    pfn_allgather allgather = nullptr;
    pfn_errstr errstr = nullptr;
} g_nccl;

static bool nccl_load() {
    if (g_nccl.tried) return g_nccl.handle != nullptr;
    g_nccl.tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return false;
    g_nccl.get_uid = (pfn_get_uid)dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.init_rank = (pfn_init_rank)dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.destroy = (pfn_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.allgather = (pfn_allgather)dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.errstr = (pfn_errstr)dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.allgather) {
        dlclose(g_nccl.handle);
        g_nccl.handle = nullptr;
        return false;
    }
    return true;
}

#define GLB_NCCL(call)                                                                              \
    do {                                                                                            \
        int r__ = (call);                                                                           \
        if (r__ != 0) {                                                                             \
            glb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                              \
                          g_nccl.errstr ? g_nccl.errstr(r__) : "nccl error");                      \
            return GLB_ENCCL;                                                                       \
        }                                                                                           \
    } while (0)

int glb_nccl_available(void) { return nccl_load() ? 1 : 0; }

int glb_nccl_unique_id(void *id128) {
    GLB_REQUIRE(id128, "id is NULL");
    if (!nccl_load()) { glb_set_error("libnccl.so.2 not found"); return GLB_ENCCL; }
    nccl_uid_t id;
    GLB_NCCL(g_nccl.get_uid(&id));
    memcpy(id128, &id, sizeof(id));
    return GLB_OK;
}

int glb_comm_init(glb_ctx_t ctx, const void *id128, int rank, int nranks) {
    GLB_REQUIRE(ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
    if (!nccl_load()) { glb_set_error("libnccl.so.2 not found"); return GLB_ENCCL; }
    GLB_CUDA(cudaSetDevice(ctx->device));
    nccl_uid_t id;
    memcpy(&id, id128, sizeof(id));
    GLB_NCCL(g_nccl.init_rank(&ctx->nccl_comm, nranks, id, rank));
    ctx->nccl_rank = rank;
    ctx->nccl_nranks = nranks;
    return GLB_OK;
}

int glb_comm_destroy(glb_ctx_t ctx) {
    if (ctx && ctx->nccl_comm && g_nccl.destroy) {
        g_nccl.destroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return GLB_OK;
}

int glb_allgather_f32(glb_ctx_t ctx, float *buf, size_t count_per_rank) {
    GLB_REQUIRE(ctx && buf, "NULL argument");
    if (!ctx->nccl_comm) { glb_set_error("glb_allgather_f32: glb_comm_init was not called"); return GLB_ENCCL; }
    const float *send = buf + size_t(ctx->nccl_rank) * count_per_rank;
    GLB_NCCL(g_nccl.allgather(send, buf, count_per_rank, /*ncclFloat32*/ 7, ctx->nccl_comm, ctx->stream));
    return GLB_OK;
}

// ------------------------------------------------------------------------ launch replay
int glb_graph_begin(glb_ctx_t ctx) {
    GLB_REQUIRE(ctx, "ctx is NULL");
    GLB_REQUIRE(!ctx->timing, "kernel timing is on: switch it off before recording");
    GLB_CUDA(cudaSetDevice(ctx->device));
    GLB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    return GLB_OK;
}

int glb_graph_end(glb_ctx_t ctx, glb_graph_t *out) {
    GLB_REQUIRE(ctx && out, "NULL argument");
    *out = nullptr;
    if (ctx->in_branch) glb_graph_branch_end(ctx);  // a failed recording is being abandoned
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (e != cudaSuccess || !graph) {
        glb_set_error("glb_graph_end: recording failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return GLB_ECUDA;
    }
    glb_graph_t g = new glb_graph_s();
    g->device = ctx->device;
    g->ctx = ctx;
    glb_ctx_retain(ctx);
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        glb_set_error("glb_graph_end: cudaGraphInstantiate: %s", cudaGetErrorString(e));
        glb_ctx_release(ctx);
        delete g;
        return GLB_ECUDA;
    }
    *out = g;
    return GLB_OK;
}

// ---- branches inside a recorded sequence -------------------------------------------------------
// The direction-optimising apps decide per level, from the size of the frontier, whether the next level
// pushes or pulls (bfs.h:186-190).  The reference reads the count on the host; here the decision stays
// on the device: the kernel that ends a push level writes it into a conditional handle of the recorded
// graph (cudaGraphSetConditional) and an IF / ELSE node (CUDA 12.8 conditional nodes, two body graphs)
// runs the push arm or the pull arm of the next level.  While a body is being recorded the context's
// stream is a side stream capturing INTO that body graph, so every glb_* launch lands there unchanged.
static int capture_info(glb_ctx_t ctx, cudaStream_t st, cudaGraph_t *graph, const cudaGraphNode_t **deps, size_t *n_deps) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    GLB_CUDA(cudaStreamGetCaptureInfo(st, &status, &id, graph, deps, n_deps));
    if (status != cudaStreamCaptureStatusActive) {
        glb_set_error("no launch sequence is being recorded (glb_graph_begin)");
        return GLB_EINVAL;
    }
    return GLB_OK;
}

int glb_graph_cond_create(glb_ctx_t ctx, uint64_t *cond) {
    GLB_REQUIRE(ctx && cond, "NULL argument");
    GLB_REQUIRE(!ctx->in_branch, "conditions are created outside branch bodies");
    cudaGraph_t graph = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t n_deps = 0;
    int rc = capture_info(ctx, ctx->stream, &graph, &deps, &n_deps);
    if (rc) return rc;
    cudaGraphConditionalHandle h;
    GLB_CUDA(cudaGraphConditionalHandleCreate(&h, graph, 0, cudaGraphCondAssignDefault));  // 0 at every replay: the ELSE arm
    static_assert(sizeof(h) == sizeof(uint64_t), "conditional handle size");
    *cond = uint64_t(h);
    return GLB_OK;
}

int glb_graph_branch_begin(glb_ctx_t ctx, uint64_t cond) {
    GLB_REQUIRE(ctx && cond, "bad argument");
    GLB_REQUIRE(!ctx->in_branch, "branches do not nest");
    cudaGraph_t graph = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t n_deps = 0;
    int rc = capture_info(ctx, ctx->stream, &graph, &deps, &n_deps);
    if (rc) return rc;
    cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};  // (the union makes the default constructor unusable)
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = cudaGraphConditionalHandle(cond);
    p.conditional.type = cudaGraphCondTypeIf;
    p.conditional.size = 2;
    cudaGraphNode_t node = nullptr;
    GLB_CUDA(cudaGraphAddNode(&node, graph, deps, n_deps, &p));
    ctx->branch_body[0] = p.conditional.phGraph_out[0];
    ctx->branch_body[1] = p.conditional.phGraph_out[1];
    // what follows the branch in the outer sequence depends on the node
    GLB_CUDA(cudaStreamUpdateCaptureDependencies(ctx->stream, &node, 1, cudaStreamSetCaptureDependencies));
    if (!ctx->branch_stream) GLB_CUDA(cudaStreamCreateWithFlags(&ctx->branch_stream, cudaStreamNonBlocking));
    GLB_CUDA(cudaStreamBeginCaptureToGraph(ctx->branch_stream, ctx->branch_body[0], nullptr, nullptr, 0,
                                           cudaStreamCaptureModeThreadLocal));
    ctx->outer_stream = ctx->stream;
    ctx->stream = ctx->branch_stream;
    ctx->in_branch = 1;
    return GLB_OK;
}

int glb_graph_branch_else(glb_ctx_t ctx) {
    GLB_REQUIRE(ctx && ctx->in_branch == 1, "no IF body is being recorded");
    cudaGraph_t g = nullptr;
    GLB_CUDA(cudaStreamEndCapture(ctx->branch_stream, &g));
    GLB_CUDA(cudaStreamBeginCaptureToGraph(ctx->branch_stream, ctx->branch_body[1], nullptr, nullptr, 0,
                                           cudaStreamCaptureModeThreadLocal));
    ctx->in_branch = 2;
    return GLB_OK;
}

int glb_graph_branch_end(glb_ctx_t ctx) {
    GLB_REQUIRE(ctx && ctx->in_branch, "no branch is being recorded");
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->branch_stream, &g);
    ctx->stream = ctx->outer_stream;
    ctx->in_branch = 0;
    if (e != cudaSuccess) {
        glb_set_error("glb_graph_branch_end: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return GLB_ECUDA;
    }
    return GLB_OK;
}

int glb_graph_launch(glb_ctx_t ctx, glb_graph_t g) {
    GLB_REQUIRE(ctx && g && g->exec, "NULL argument");
    GLB_REQUIRE(g->device == ctx->device, "sequence was recorded on another device");
    GLB_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
    return GLB_OK;
}

int glb_graph_destroy(glb_graph_t g) {
    if (!g) return GLB_OK;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    glb_ctx_release(g->ctx);
    delete g;
    return GLB_OK;
}

}  // extern "C"
