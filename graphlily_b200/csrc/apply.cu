// Element-wise "apply" operators (overlay modes 3-6).
//
//   mode 3  kernel_add_scalar_vector_dense          /root/reference/graphlily/hw/kernel_add_scalar_vector_dense_impl.h:6-27
//   mode 4  kernel_assign_vector_dense              kernel_assign_vector_dense_impl.h:8-47
//   mode 5  kernel_assign_vector_sparse_no_new_frontier   kernel_assign_vector_sparse_no_new_frontier_impl.h:4-55
//   mode 6  kernel_assign_vector_sparse_new_frontier      kernel_assign_vector_sparse_new_frontier_impl.h:4-78
// with the semantics of the modules' compute_reference_results
// (add_scalar_vector_dense_module.h:195-204, assign_vector_dense_module.h:223-246,
//  assign_vector_sparse_module.h:306-335).
//
// All four are pure HBM streams / scatters: 128-bit vector accesses where the layout allows,
// grids sized as multiples of the SM count, list lengths read on the device from slot 0.
#include <string.h>

#include "glb_internal.h"
#include "semiring.cuh"

namespace {

constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

template <int VT>
__global__ void __launch_bounds__(kThreads) ewise_add_kernel(const float *in, float *out, uint32_t len, float val) {
    const uint32_t stride = gridDim.x * kThreads;
    const uint32_t tid = blockIdx.x * kThreads + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    uint32_t done = 0;
    if (aligned) {
        const uint32_t n4 = len >> 2;
        const float4 *in4 = reinterpret_cast<const float4 *>(in);
        float4 *out4 = reinterpret_cast<float4 *>(out);
        for (uint32_t i = tid; i < n4; i += stride) {
            float4 v = in4[i];
            v.x = Val<VT>::plus(v.x, val);
            v.y = Val<VT>::plus(v.y, val);
            v.z = Val<VT>::plus(v.z, val);
            v.w = Val<VT>::plus(v.w, val);
            out4[i] = v;
        }
        done = n4 << 2;
    }
    for (uint32_t i = done + tid; i < len; i += stride) out[i] = Val<VT>::plus(in[i], val);
}

template <int VT>
__global__ void __launch_bounds__(kThreads) assign_dense_kernel(const float *mask, float *inout, uint32_t len, float val,
                                                              int write_to_one) {
    const uint32_t stride = gridDim.x * kThreads;
    const uint32_t tid = blockIdx.x * kThreads + threadIdx.x;
    const bool want = write_to_one != 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(inout)) & 15u) == 0;
    uint32_t done = 0;
    if (aligned) {  // 128-bit mask loads; a group of four that is assigned entirely goes out as one 128-bit store
        const uint32_t n4 = len >> 2;
        const float4 *m4 = reinterpret_cast<const float4 *>(mask);
        float4 *io4 = reinterpret_cast<float4 *>(inout);
        for (uint32_t i = tid; i < n4; i += stride) {
            const float4 m = m4[i];
            const bool h0 = !Val<VT>::is_zero(m.x) == want, h1 = !Val<VT>::is_zero(m.y) == want,
                       h2 = !Val<VT>::is_zero(m.z) == want, h3 = !Val<VT>::is_zero(m.w) == want;
            if (h0 & h1 & h2 & h3) {
                io4[i] = make_float4(val, val, val, val);
            } else {
                float *p = inout + 4 * size_t(i);
                if (h0) p[0] = val;
                if (h1) p[1] = val;
                if (h2) p[2] = val;
                if (h3) p[3] = val;
            }
        }
        done = n4 << 2;
    }
    for (uint32_t i = done + tid; i < len; i += stride) {
        const bool nz = !Val<VT>::is_zero(mask[i]);
        if (nz == want) inout[i] = val;
    }
}

__global__ void __launch_bounds__(kThreads) assign_sparse_kernel(const glb_idx_val_t *__restrict__ list, float *inout,
                                                               float val) {
    const uint32_t n = list[0].index;
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) inout[list[i + 1].index] = val;
}

// inout[idx] = min(inout[idx], val) with the reference's test `inout[idx] > val`
// (assign_vector_sparse_module.h:318-335); true when THIS entry lowered the value.  A compare-and-
// swap loop instead of a plain compare-then-store so that lists naming an index more than once
// end with the minimum whatever the interleaving (NaN never wins the comparison, as in the host loop).
template <int VT>
__device__ __forceinline__ bool relax_min(float *addr, float val) {
    float old = *addr;
    while (Val<VT>::less(val, old)) {
        const int seen = atomicCAS(reinterpret_cast<int *>(addr), __float_as_int(old), __float_as_int(val));
        if (seen == __float_as_int(old)) return true;
        old = __int_as_float(seen);
    }
    return false;
}

// Overlay mode 6.  With distinct indices (what SpMSpV emits and the apps feed) the result equals
// the reference's sequential loop: same inout, same set of new-frontier entries (order
// unspecified).  With repeated indices inout is still the reference's (the minimum); the new
// frontier then holds every entry that lowered the value at the moment it was applied -- always
// including the one carrying the final minimum -- where the reference lists the record lows in
// list order.
template <int VT>
__global__ void __launch_bounds__(kThreads) assign_sparse_relax_kernel(const glb_idx_val_t *__restrict__ list,
                                                                     float *inout, glb_idx_val_t *new_frontier) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t n = list[0].index;
    const uint32_t n_round = (n + 31u) & ~31u;
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_round; i += stride) {
        bool emit = false;
        glb_idx_val_t e = {0u, 0.0f};
        if (i < n) {
            e = list[i + 1];
            emit = relax_min<VT>(inout + e.index, e.val);
        }
        const unsigned b = __ballot_sync(kFull, emit);
        if (b) {
            uint32_t base = 0;
            const int leader = __ffs(int(b)) - 1;
            if (int(lane) == leader) base = atomicAdd(&new_frontier[0].index, uint32_t(__popc(b)));
            base = __shfl_sync(kFull, base, leader);
            if (emit) new_frontier[1 + base + __popc(b & ((1u << lane) - 1u))] = e;
        }
    }
}

inline unsigned grid_for(glb_ctx_t ctx, uint64_t work_items, int per_sm) {
    uint64_t blocks = (work_items + kThreads - 1) / kThreads;
    const uint64_t cap = uint64_t(ctx->num_sms) * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return unsigned(blocks);
}

}  // namespace

static inline float from_bits(uint32_t bits) {
    float f;
    memcpy(&f, &bits, sizeof(f));
    return f;
}

#define GLB_VT_DISPATCH(vt, CALL)                                      \
    switch (vt) {                                                      \
        case GLB_VAL_F32: { constexpr int VT = GLB_VAL_F32; CALL; break; }       \
        case GLB_VAL_U32: { constexpr int VT = GLB_VAL_U32; CALL; break; }       \
        case GLB_VAL_UFIXED: { constexpr int VT = GLB_VAL_UFIXED; CALL; break; } \
        default: glb_set_error("%s: invalid value type %d", __func__, vt); return GLB_EINVAL; \
    }

extern "C" {

int glb_ewise_add_vt(glb_ctx_t ctx, int val_type, const void *in, void *out, uint32_t len, uint32_t val_bits) {
    GLB_REQUIRE(ctx && (len == 0 || (in && out)), "NULL argument");
    if (len == 0) return GLB_OK;
    GLB_VT_DISPATCH(val_type, (ewise_add_kernel<VT><<<grid_for(ctx, (uint64_t(len) + 3) / 4, 8), kThreads, 0, ctx->stream>>>(
                                   static_cast<const float *>(in), static_cast<float *>(out), len, from_bits(val_bits))));
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_ewise_add(glb_ctx_t ctx, const float *in, float *out, uint32_t len, float val) {
    uint32_t bits;
    memcpy(&bits, &val, sizeof(bits));
    return glb_ewise_add_vt(ctx, GLB_VAL_F32, in, out, len, bits);
}

int glb_assign_dense_vt(glb_ctx_t ctx, int val_type, const void *mask, void *inout, uint32_t len, uint32_t val_bits,
                        int mask_type) {
    GLB_REQUIRE(ctx && (len == 0 || (mask && inout)), "NULL argument");
    // the reference prints "Please set the mask type" and exits (assign_vector_dense_module.h:88-95)
    GLB_REQUIRE(mask_type == GLB_MASK_WRITE_TO_ZERO || mask_type == GLB_MASK_WRITE_TO_ONE,
                "dense assign needs kMaskWriteToZero or kMaskWriteToOne");
    if (len == 0) return GLB_OK;
    GLB_VT_DISPATCH(val_type, (assign_dense_kernel<VT><<<grid_for(ctx, (uint64_t(len) + 3) / 4, 8), kThreads, 0, ctx->stream>>>(
                                   static_cast<const float *>(mask), static_cast<float *>(inout), len, from_bits(val_bits),
                                   mask_type == GLB_MASK_WRITE_TO_ONE)));
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_assign_dense(glb_ctx_t ctx, const float *mask, float *inout, uint32_t len, float val, int mask_type) {
    uint32_t bits;
    memcpy(&bits, &val, sizeof(bits));
    return glb_assign_dense_vt(ctx, GLB_VAL_F32, mask, inout, len, bits, mask_type);
}

int glb_assign_sparse(glb_ctx_t ctx, const glb_idx_val_t *list, float *inout, float val) {
    GLB_REQUIRE(ctx && list && inout, "NULL argument");
    assign_sparse_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(list, inout, val);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_assign_sparse_vt(glb_ctx_t ctx, int val_type, const glb_idx_val_t *list, void *inout, uint32_t val_bits) {
    GLB_REQUIRE(val_type >= GLB_VAL_F32 && val_type <= GLB_VAL_UFIXED, "invalid value type");
    return glb_assign_sparse(ctx, list, static_cast<float *>(inout), from_bits(val_bits));  // a store of the word: the same for every type
}

int glb_assign_sparse_relax_vt(glb_ctx_t ctx, int val_type, const glb_idx_val_t *list, void *inout, glb_idx_val_t *new_frontier) {
    GLB_REQUIRE(ctx && list && inout && new_frontier, "NULL argument");
    GLB_REQUIRE(static_cast<const void *>(list) != static_cast<const void *>(new_frontier),
                "new_frontier must not alias list");
    GLB_CUDA(cudaMemsetAsync(new_frontier, 0, sizeof(glb_idx_val_t), ctx->stream));  // head = {0, 0}
    GLB_VT_DISPATCH(val_type, (assign_sparse_relax_kernel<VT><<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(
                                   list, static_cast<float *>(inout), new_frontier)));
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_assign_sparse_relax(glb_ctx_t ctx, const glb_idx_val_t *list, float *inout, glb_idx_val_t *new_frontier) {
    return glb_assign_sparse_relax_vt(ctx, GLB_VAL_F32, list, inout, new_frontier);
}

}  // extern "C"
