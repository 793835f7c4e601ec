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
#include <cuda.h>  // driver API TYPES only: the entry points are resolved at run time (cudaGetDriverEntryPoint)
#include <string.h>
#include <unistd.h>

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

// ---- multicast exchange created through the C ABI ------------------------------------------------
// NVSwitch multicast object + this rank's block + both mappings, made with the driver's virtual-memory API
// (cuMulticastCreate / cuMulticastAddDevice / cuMemCreate / cuMulticastBindMem / cuMemMap) -- what
// graphlily_b200/exchange.py otherwise asks torch's symmetric memory for.  The driver entry points are looked up
// at run time, so the library still loads (and links) on a host without libcuda.
//   rank 0:   glb_xchg_mc_open(..., fd_in = -1, &xc, &fd_out)  creates the object; fd_out is its POSIX file descriptor,
//             which the host passes to the other ranks' processes (SCM_RIGHTS over a Unix socket, pidfd_getfd, ...)
//   rank r:   glb_xchg_mc_open(..., fd_in = <that descriptor in THIS process>, &xc, NULL)
//   -- host barrier: every rank has added its device --
//   all:      glb_xchg_mc_bind(xc)   allocates + binds + maps; the exchange is then connected (multicast only: peers'
//             blocks are not mapped one by one, every transfer goes through the switch)
//   -- host barrier before the first use --
namespace {

struct DriverApi {
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    bool ok = false;
};

const DriverApi &driver_api() {
    static const DriverApi api = [] {
        DriverApi a;
        bool ok = true;
        auto get = [&](const char *name, void *slot) {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
                cudaGetLastError();
                ok = false;
            }
            memcpy(slot, &fn, sizeof(fn));
        };
        get("cuMulticastCreate", &a.MulticastCreate);
        get("cuMulticastAddDevice", &a.MulticastAddDevice);
        get("cuMulticastBindMem", &a.MulticastBindMem);
        get("cuMulticastUnbind", &a.MulticastUnbind);
        get("cuMulticastGetGranularity", &a.MulticastGetGranularity);
        get("cuMemCreate", &a.MemCreate);
        get("cuMemRelease", &a.MemRelease);
        get("cuMemGetAllocationGranularity", &a.MemGetAllocationGranularity);
        get("cuMemAddressReserve", &a.MemAddressReserve);
        get("cuMemAddressFree", &a.MemAddressFree);
        get("cuMemMap", &a.MemMap);
        get("cuMemUnmap", &a.MemUnmap);
        get("cuMemSetAccess", &a.MemSetAccess);
        get("cuMemExportToShareableHandle", &a.MemExportToShareableHandle);
        get("cuMemImportFromShareableHandle", &a.MemImportFromShareableHandle);
        get("cuDeviceGet", &a.DeviceGet);
        get("cuDeviceGetAttribute", &a.DeviceGetAttribute);
        get("cuGetErrorString", &a.GetErrorString);
        a.ok = ok;
        return a;
    }();
    return api;
}

#define GLB_CU(call)                                                                  \
    do {                                                                              \
        const CUresult glb_cu_ = (call);                                              \
        if (glb_cu_ != CUDA_SUCCESS) {                                                \
            const char *msg_ = nullptr;                                               \
            if (D.GetErrorString) D.GetErrorString(glb_cu_, &msg_);                   \
            glb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, msg_ ? msg_ : "driver error"); \
            return GLB_ECUDA;                                                         \
        }                                                                             \
    } while (0)

CUmulticastObjectProp mc_prop(int nranks, size_t size) {
    CUmulticastObjectProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.numDevices = unsigned(nranks);
    prop.size = size;
    prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

CUmemAllocationProp mc_alloc_prop(int device) {
    CUmemAllocationProp ap;
    memset(&ap, 0, sizeof(ap));
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return ap;
}

void mc_release(glb_xchg_t xc) {
    const DriverApi &D = driver_api();
    if (!D.ok) return;
    if (xc->mc) { D.MemUnmap(CUdeviceptr(reinterpret_cast<uintptr_t>(xc->mc)), xc->mc_size); D.MemAddressFree(CUdeviceptr(reinterpret_cast<uintptr_t>(xc->mc)), xc->mc_size); }
    if (xc->local) { D.MemUnmap(CUdeviceptr(reinterpret_cast<uintptr_t>(xc->local)), xc->mc_size); D.MemAddressFree(CUdeviceptr(reinterpret_cast<uintptr_t>(xc->local)), xc->mc_size); }
    if (xc->mc_bound) {
        CUdevice dev;
        if (D.DeviceGet(&dev, xc->ctx->device) == CUDA_SUCCESS) D.MulticastUnbind(xc->mc_handle, dev, 0, xc->mc_size);
    }
    if (xc->mem_handle) D.MemRelease(xc->mem_handle);
    if (xc->mc_handle) D.MemRelease(xc->mc_handle);
    xc->mc = nullptr;
    xc->local = nullptr;
}

}  // namespace

extern "C" {

int glb_xchg_mc_supported(glb_ctx_t ctx, int *supported) {
    GLB_REQUIRE(ctx && supported, "NULL argument");
    *supported = 0;
    GLB_CUDA(cudaSetDevice(ctx->device));
    GLB_CUDA(cudaFree(nullptr));  // the primary context exists
    const DriverApi &D = driver_api();
    if (!D.ok) return GLB_OK;
    CUdevice dev;
    int v = 0;
    if (D.DeviceGet(&dev, ctx->device) == CUDA_SUCCESS &&
        D.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS)
        *supported = v;
    return GLB_OK;
}

int glb_xchg_mc_open(glb_ctx_t ctx, uint32_t n_floats, int n_vectors, int rank, int nranks, int fd_in, glb_xchg_t *out, int *fd_out) {
    GLB_REQUIRE(ctx && out && n_floats > 0 && n_vectors >= 2 && n_vectors <= 4, "bad argument");
    GLB_REQUIRE(nranks >= 2 && nranks <= GLB_MAX_PEERS + 1 && rank >= 0 && rank < nranks, "bad rank");
    GLB_REQUIRE((rank == 0) == (fd_in < 0), "rank 0 creates the multicast object (fd_in = -1); every other rank passes its descriptor");
    GLB_REQUIRE(rank != 0 || fd_out, "rank 0 needs fd_out");
    *out = nullptr;
    GLB_CUDA(cudaSetDevice(ctx->device));
    GLB_CUDA(cudaFree(nullptr));
    const DriverApi &D = driver_api();
    GLB_REQUIRE(D.ok, "this driver has no multicast entry points");
    CUdevice dev;
    GLB_CU(D.DeviceGet(&dev, ctx->device));
    int supported = 0;
    GLB_CU(D.DeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
    GLB_REQUIRE(supported, "the device does not support multicast objects (no NVSwitch?)");
    // same size on every rank: the block rounded to the multicast granularity
    const size_t bytes = glb_xchg_block_bytes(n_floats, n_vectors);
    CUmulticastObjectProp prop = mc_prop(nranks, bytes);
    size_t gran = 0, gran_alloc = 0;
    GLB_CU(D.MulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_MINIMUM));
    const CUmemAllocationProp ap = mc_alloc_prop(ctx->device);
    GLB_CU(D.MemGetAllocationGranularity(&gran_alloc, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (gran_alloc > gran) gran = gran_alloc;  // (powers of two)
    const size_t size = (bytes + gran - 1) / gran * gran;
    prop.size = size;
    CUmemGenericAllocationHandle mch = 0;
    if (rank == 0) {
        GLB_CU(D.MulticastCreate(&mch, &prop));
        int fd = -1;
        const CUresult r = D.MemExportToShareableHandle(&fd, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
        if (r != CUDA_SUCCESS) {
            D.MemRelease(mch);
            glb_set_error("glb_xchg_mc_open: exporting the multicast object failed");
            return GLB_ECUDA;
        }
        *fd_out = fd;
    } else {
        GLB_CU(D.MemImportFromShareableHandle(&mch, reinterpret_cast<void *>(static_cast<uintptr_t>(fd_in)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
        if (fd_out) *fd_out = -1;
    }
    const CUresult added = D.MulticastAddDevice(mch, dev);
    if (added != CUDA_SUCCESS) {
        D.MemRelease(mch);
        glb_set_error("glb_xchg_mc_open: cuMulticastAddDevice failed (%d)", int(added));
        return GLB_ECUDA;
    }
    glb_xchg_t xc = new glb_xchg_s();
    xc->ctx = ctx;
    glb_ctx_retain(ctx);
    xc->n = n_floats;
    xc->n_vectors = n_vectors;
    xc->rank = rank;
    xc->nranks = nranks;
    xc->adopted = true;  // (nothing here is cudaMalloc / CUDA-IPC memory)
    xc->mc_native = true;
    xc->mc_handle = mch;
    xc->mc_size = size;
    *out = xc;
    return GLB_OK;
}

int glb_xchg_mc_bind(glb_xchg_t xc) {
    GLB_REQUIRE(xc && xc->mc_native && !xc->connected, "not an exchange opened with glb_xchg_mc_open (or bound already)");
    glb_ctx_t ctx = xc->ctx;
    GLB_CUDA(cudaSetDevice(ctx->device));
    const DriverApi &D = driver_api();
    CUdevice dev;
    GLB_CU(D.DeviceGet(&dev, ctx->device));
    const CUmemAllocationProp ap = mc_alloc_prop(ctx->device);
    size_t gran = 0;
    GLB_CU(D.MemGetAllocationGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    GLB_REQUIRE(xc->mc_size % gran == 0, "multicast size is not a multiple of the allocation granularity");
    CUmemGenericAllocationHandle mem = 0;
    GLB_CU(D.MemCreate(&mem, xc->mc_size, &ap, 0));
    xc->mem_handle = mem;
    GLB_CU(D.MulticastBindMem(xc->mc_handle, 0, mem, 0, xc->mc_size, 0));
    xc->mc_bound = true;
    CUmemAccessDesc access;
    memset(&access, 0, sizeof(access));
    access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    access.location.id = ctx->device;
    access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CUdeviceptr uva = 0, mva = 0;
    GLB_CU(D.MemAddressReserve(&uva, xc->mc_size, gran, 0, 0));
    xc->local = reinterpret_cast<float *>(static_cast<uintptr_t>(uva));
    GLB_CU(D.MemMap(uva, xc->mc_size, 0, mem, 0));
    GLB_CU(D.MemSetAccess(uva, xc->mc_size, &access, 1));
    GLB_CU(D.MemAddressReserve(&mva, xc->mc_size, gran, 0, 0));
    xc->mc = reinterpret_cast<float *>(static_cast<uintptr_t>(mva));
    GLB_CU(D.MemMap(mva, xc->mc_size, 0, xc->mc_handle, 0));
    GLB_CU(D.MemSetAccess(mva, xc->mc_size, &access, 1));
    GLB_CUDA(cudaMemset(xc->local, 0, xc->mc_size));
    GLB_CUDA(cudaDeviceSynchronize());
    const size_t vec_bytes = glb_xchg_block_bytes(xc->n, xc->n_vectors) - 256;
    xc->local_flags = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(xc->local) + vec_bytes);
    xc->mc_flags = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(xc->mc) + vec_bytes);
    xc->peer[xc->rank] = xc->local;  // the other ranks' blocks are reached through the switch only
    xc->peer_flags[xc->rank] = xc->local_flags;
    int rc = alloc_device_state(xc);
    if (!rc && cudaMemcpy(xc->d_peer_flags, xc->peer_flags, sizeof(uint32_t *) * (GLB_MAX_PEERS + 1), cudaMemcpyHostToDevice) != cudaSuccess) {
        glb_set_error("glb_xchg_mc_bind: %s", cudaGetErrorString(cudaGetLastError()));
        rc = GLB_ECUDA;
    }
    if (rc) return rc;
    xc->connected = true;
    return GLB_OK;
}

}  // extern "C"

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
    if (xc->mc_native) {
        mc_release(xc);
    } else if (!xc->adopted) {
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
