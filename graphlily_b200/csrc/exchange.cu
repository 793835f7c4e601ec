// Exchange of a row-sharded run: one process per GPU, every rank owns a
// block [n_vectors x n floats][flag words] that all ranks map (CUDA IPC, or memory the host mapped
// itself, optionally with an NVSwitch multicast mapping).  A step's slice of y reaches the other
// ranks by
//   * multicast:  ONE kernel that copies the finished slice to the multicast address in 16-byte
//                 multimem.st stores (the switch replicates each store into every rank's block) and
//                 whose last CTA publishes the epoch -- xchg_push_signal_kernel;
//   * peer:       the SpMV write-back stores each row into every peer's block (spmv.cu) and one
//                 tiny kernel publishes the epoch -- xchg_signal_kernel;
//   * copies:     peer-to-peer cudaMemcpyAsync + the same signal kernel (glb_xchg_allgather without
//                 a multicast mapping);
// and the matching acquire is NOT a launch of its own inside an iteration: it sits at the head of
// the kernel that opens the next SpMV (glb_xchg_wait_head, exchange.cuh).  The epoch lives in
// device memory and is advanced by the publishing kernel, so a recorded launch sequence (CUDA
// graph) of many steps replays correctly.
#include <string.h>

#include "exchange.cuh"
#include "glb_internal.h"

namespace {

// d_state: [0] epoch, [1] CTA ticket of the push kernel, [2] timeout flag
__device__ __forceinline__ uint32_t next_epoch(uint32_t *state) {
    const uint32_t e = state[0] + 1;
    state[0] = e;
    return e;
}

// Publish (and optionally wait): one CTA of 32 threads.  All stores of this rank's slice were
// issued by kernels / copies that completed before this kernel started (stream order).
__global__ void xchg_signal_kernel(uint32_t *const *peer_flags, uint32_t *local_flags, uint32_t *mc_flags, int rank,
                                   int nranks, uint32_t *state, int do_wait) {
    __shared__ uint32_t s_epoch;
    const int p = int(threadIdx.x);
    if (p == 0) s_epoch = next_epoch(state);
    __syncthreads();
    const uint32_t epoch = s_epoch;
    if (mc_flags) {
        // Data that travelled as multicast stores is published by a multicast store too, so the flag
        // follows the data through the switch on every destination -- this rank included: its own
        // slice comes back through the switch as well, and nothing may overwrite it locally before
        // that copy has landed.  Hence waits cover all ranks, not only the peers.
        if (p == 0) {
            __threadfence_system();
            asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(mc_flags + rank), "r"(epoch) : "memory");
        }
    } else if (p < nranks) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[p] + rank), "r"(epoch) : "memory");
    }
    if (do_wait && p < nranks) glb_xchg_spin(local_flags + p, epoch, state + 2);
}

__global__ void xchg_wait_kernel(const GlbXchgWait w) { glb_xchg_wait_head(w); }

// One store, every rank: multimem.st on the multicast mapping of the blocks is replicated by the
// NVSwitch into all ranks' copies (the issuing rank's included), so a slice leaves this GPU once
// instead of once per peer.  16-byte stores; the unaligned head / tail go out as scalars.  The CTA
// that finishes last publishes the epoch: every CTA fences its stores system-wide before it takes
// a ticket, so the release store of the flag is ordered after all of them.
__global__ void __launch_bounds__(256) xchg_push_signal_kernel(const float *__restrict__ src, float *mc, size_t count,
                                                              uint32_t *mc_flags, int rank, uint32_t *state) {
    const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x, stride = size_t(gridDim.x) * blockDim.x;
    size_t head = (16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15;
    head = head / 4 < count ? head / 4 : count;
    const size_t n4 = (count - head) / 4;
    for (size_t i = tid; i < head; i += stride)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + i), "f"(src[i]) : "memory");
    const float4 *s4 = reinterpret_cast<const float4 *>(src + head);
    for (size_t i = tid; i < n4; i += stride) {
        const float4 v = s4[i];
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + head + 4 * i), "f"(v.x),
                     "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
    for (size_t i = head + 4 * n4 + tid; i < count; i += stride)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + i), "f"(src[i]) : "memory");
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(state + 1, 1u);
        if (ticket == gridDim.x - 1) {
            state[1] = 0;
            __threadfence_system();
            const uint32_t epoch = next_epoch(state);
            asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(mc_flags + rank), "r"(epoch) : "memory");
        }
    }
}

int alloc_device_state(glb_xchg_t xc) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&xc->d_peer_flags), sizeof(uint32_t *) * (GLB_MAX_PEERS + 1));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&xc->d_state), 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(xc->d_state, 0, 4 * sizeof(uint32_t));
    if (e != cudaSuccess) {
        glb_set_error("exchange: %s", cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? GLB_ENOMEM : GLB_ECUDA;
    }
    return GLB_OK;
}

}  // namespace

// ---- internal entry points (glb_internal.h) -----------------------------------------------------
GlbXchgWait glb_xchg_wait_desc(glb_xchg_t xc) {
    GlbXchgWait w;
    w.flags = xc->local_flags;
    w.state = xc->d_state;
    w.err = xc->d_state + 2;
    w.n = xc->nranks;
    return w;
}

int glb_xchg_preload() {
    cudaFuncAttributes fa;
    GLB_CUDA(cudaFuncGetAttributes(&fa, xchg_signal_kernel));
    GLB_CUDA(cudaFuncGetAttributes(&fa, xchg_wait_kernel));
    GLB_CUDA(cudaFuncGetAttributes(&fa, xchg_push_signal_kernel));
    return GLB_OK;
}

int glb_xchg_signal(glb_ctx_t ctx, glb_xchg_t xc, bool wait) {
    xchg_signal_kernel<<<1, 32, 0, ctx->stream>>>(xc->d_peer_flags, xc->local_flags, xc->mc_flags, xc->rank, xc->nranks,
                                                 xc->d_state, wait ? 1 : 0);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_xchg_wait_launch(glb_ctx_t ctx, const GlbXchgWait &w) {
    xchg_wait_kernel<<<1, 32, 0, ctx->stream>>>(w);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_xchg_wait(glb_ctx_t ctx, glb_xchg_t xc) { return glb_xchg_wait_launch(ctx, glb_xchg_wait_desc(xc)); }

// Slice [offset, offset + count) of vector `which` goes to every rank and the epoch is published;
// the acquire is left to the caller (glb_xchg_wait, or the head of the next SpMV).
int glb_xchg_push(glb_ctx_t ctx, glb_xchg_t xc, int which, size_t offset, size_t count) {
    const size_t at = size_t(which) * xc->n + offset;
    if (xc->mc && xc->nranks > 1) {
        size_t blocks = (count / 4 + 255) / 256 + 1;
        const size_t cap = size_t(ctx->num_sms) * 4;
        if (blocks > cap) blocks = cap;
        xchg_push_signal_kernel<<<unsigned(blocks), 256, 0, ctx->stream>>>(xc->local + at, xc->mc + at, count, xc->mc_flags,
                                                                         xc->rank, xc->d_state);
        GLB_CUDA(cudaGetLastError());
        return GLB_OK;
    }
    for (int r = 0; r < xc->nranks && count; ++r)
        if (r != xc->rank)
            GLB_CUDA(cudaMemcpyAsync(xc->peer[r] + at, xc->local + at, count * sizeof(float), cudaMemcpyDeviceToDevice,
                                     ctx->stream));
    return glb_xchg_signal(ctx, xc, false);
}

extern "C" {

int glb_xchg_signal_wait(glb_ctx_t ctx, glb_xchg_t xc) { return glb_xchg_signal(ctx, xc, true); }

int glb_xchg_create(glb_ctx_t ctx, uint32_t n_floats, int n_vectors, glb_xchg_t *out) {
    GLB_REQUIRE(ctx && out && n_floats > 0 && n_vectors >= 2 && n_vectors <= 4, "bad argument");
    *out = nullptr;
    GLB_CUDA(cudaSetDevice(ctx->device));
    glb_xchg_t xc = new glb_xchg_s();
    xc->ctx = ctx;
    glb_ctx_retain(ctx);
    xc->n = n_floats;
    xc->n_vectors = n_vectors;
    const size_t vec_bytes = glb_xchg_block_bytes(n_floats, n_vectors) - 256;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&xc->local), vec_bytes + 256);  // plain cudaMalloc: IPC-exportable
    if (e == cudaSuccess) e = cudaMemset(xc->local, 0, vec_bytes + 256);
    int rc = GLB_OK;
    if (e != cudaSuccess) {
        glb_set_error("glb_xchg_create: %s", cudaGetErrorString(e));
        rc = e == cudaErrorMemoryAllocation ? GLB_ENOMEM : GLB_ECUDA;
    }
    if (!rc) rc = alloc_device_state(xc);
    if (rc) {
        glb_xchg_destroy(xc);
        return rc;
    }
    xc->local_flags = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(xc->local) + vec_bytes);
    *out = xc;
    return GLB_OK;
}

int glb_xchg_export(glb_xchg_t xc, void *handle64) {
    GLB_REQUIRE(xc && handle64, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == GLB_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    GLB_CUDA(cudaIpcGetMemHandle(&h, xc->local));
    memcpy(handle64, &h, sizeof(h));
    return GLB_OK;
}

int glb_xchg_connect(glb_xchg_t xc, int rank, int nranks, const void *handles) {
    GLB_REQUIRE(xc && handles && nranks >= 1 && nranks <= GLB_MAX_PEERS + 1 && rank >= 0 && rank < nranks, "bad argument");
    GLB_CUDA(cudaSetDevice(xc->ctx->device));
    const size_t vec_bytes = glb_xchg_block_bytes(xc->n, xc->n_vectors) - 256;
    xc->rank = rank;
    xc->nranks = nranks;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            xc->peer[r] = xc->local;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const char *>(handles) + size_t(r) * GLB_IPC_HANDLE_BYTES, sizeof(h));
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                glb_set_error("glb_xchg_connect: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
                cudaGetLastError();
                return GLB_ECUDA;
            }
            xc->peer[r] = static_cast<float *>(p);
        }
        xc->peer_flags[r] = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(xc->peer[r]) + vec_bytes);
    }
    GLB_CUDA(cudaMemcpy(xc->d_peer_flags, xc->peer_flags, sizeof(uint32_t *) * (GLB_MAX_PEERS + 1), cudaMemcpyHostToDevice));
    xc->connected = true;
    return GLB_OK;
}

int glb_xchg_adopt(glb_ctx_t ctx, uint32_t n_floats, int n_vectors, int rank, int nranks, void *const *blocks,
                   void *multicast_block, glb_xchg_t *out) {
    GLB_REQUIRE(ctx && out && blocks && n_floats > 0 && n_vectors >= 2 && n_vectors <= 4, "bad argument");
    GLB_REQUIRE(nranks >= 1 && nranks <= GLB_MAX_PEERS + 1 && rank >= 0 && rank < nranks, "bad rank");
    *out = nullptr;
    GLB_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < nranks; ++r) GLB_REQUIRE(blocks[r], "NULL block");
    glb_xchg_t xc = new glb_xchg_s();
    xc->ctx = ctx;
    glb_ctx_retain(ctx);
    xc->n = n_floats;
    xc->n_vectors = n_vectors;
    xc->adopted = true;
    xc->rank = rank;
    xc->nranks = nranks;
    xc->mc = static_cast<float *>(multicast_block);
    const size_t vec_bytes = glb_xchg_block_bytes(n_floats, n_vectors) - 256;
    for (int r = 0; r < nranks; ++r) {
        xc->peer[r] = static_cast<float *>(blocks[r]);
        xc->peer_flags[r] = reinterpret_cast<uint32_t *>(static_cast<char *>(blocks[r]) + vec_bytes);
    }
    xc->local = xc->peer[rank];
    xc->local_flags = xc->peer_flags[rank];
    if (xc->mc) xc->mc_flags = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(xc->mc) + vec_bytes);
    int rc = alloc_device_state(xc);
    if (!rc && cudaMemcpy(xc->d_peer_flags, xc->peer_flags, sizeof(uint32_t *) * (GLB_MAX_PEERS + 1), cudaMemcpyHostToDevice) !=
                   cudaSuccess) {
        glb_set_error("glb_xchg_adopt: %s", cudaGetErrorString(cudaGetLastError()));
        rc = GLB_ECUDA;
    }
    if (rc) {
        glb_xchg_destroy(xc);
        return rc;
    }
    xc->connected = true;
    *out = xc;
    return GLB_OK;
}

size_t glb_xchg_block_bytes(uint32_t n_floats, int n_vectors) {
    return ((size_t(n_floats) * size_t(n_vectors) * sizeof(float) + 255) & ~size_t(255)) + 256;
}

int glb_xchg_has_multicast(glb_xchg_t xc) { return xc && xc->mc ? 1 : 0; }

int glb_xchg_vector(glb_xchg_t xc, int which, float **local_ptr) {
    GLB_REQUIRE(xc && local_ptr && which >= 0 && which < xc->n_vectors, "bad argument");
    *local_ptr = xc->local + size_t(which) * xc->n;
    return GLB_OK;
}

int glb_xchg_barrier(glb_ctx_t ctx, glb_xchg_t xc) {
    GLB_REQUIRE(ctx && xc && xc->connected && xc->ctx == ctx, "exchange is not connected to this context");
    return glb_xchg_signal(ctx, xc, true);
}

int glb_xchg_allgather(glb_ctx_t ctx, glb_xchg_t xc, int which, size_t offset, size_t count) {
    GLB_REQUIRE(ctx && xc && xc->connected && xc->ctx == ctx, "exchange is not connected to this context");
    GLB_REQUIRE(which >= 0 && which < xc->n_vectors && offset + count <= xc->n, "slice outside the vector");
    int rc = glb_xchg_push(ctx, xc, which, offset, count);
    if (rc) return rc;
    return glb_xchg_wait(ctx, xc);  // callers of the public entry point read the vector next: acquire here
}

int glb_xchg_status(glb_xchg_t xc, int *timed_out) {
    GLB_REQUIRE(xc && timed_out, "NULL argument");
    uint32_t e = 0;
    GLB_CUDA(cudaStreamSynchronize(xc->ctx->stream));
    GLB_CUDA(cudaMemcpy(&e, xc->d_state + 2, sizeof(e), cudaMemcpyDeviceToHost));
    *timed_out = int(e);
    return GLB_OK;
}

int glb_xchg_destroy(glb_xchg_t xc) {
    if (!xc) return GLB_OK;
    cudaSetDevice(xc->ctx->device);
    cudaStreamSynchronize(xc->ctx->stream);
    if (!xc->adopted) {
        for (int r = 0; r < xc->nranks; ++r)
            if (xc->connected && r != xc->rank && xc->peer[r]) cudaIpcCloseMemHandle(xc->peer[r]);
        cudaFree(xc->local);
    }
    cudaFree(xc->d_peer_flags);
    cudaFree(xc->d_state);
    glb_ctx_release(xc->ctx);
    delete xc;
    return GLB_OK;
}

}  // extern "C"
