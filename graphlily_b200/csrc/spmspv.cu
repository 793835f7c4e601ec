// SpMSpV: y = A[:, idx(x)] (+).(x) x, CSC, sparse in -> sparse out (overlay mode 2).
//
// Replaces kernel_spmspv (/root/reference/graphlily/hw/kernel_spmspv_impl.h:448-562) and
// formatCSC (graphlily/io/data_formatter.h:608-721).  Semantics: SpMSpVModule::
// compute_reference_results (graphlily/module/spmspv_module.h:446-520) followed by the device
// write-back rule (kernel_spmspv_impl.h:200-229,263-281,551-555): only entries != zero that
// pass the mask are listed, the head slot is {count, zero}, order is unspecified.
//
// Three stream-ordered launches, none sized by device data (persistent grids read the
// frontier length from x[0].index):
//   1. scatter (column heads): one warp per active column combines the first kSeg non-zeros of
//      the column, a (x) v, into a dense accumulator that rests at the (+)-identity
//      (atomicAdd / benign store of 1.0f / ordered-int atomicMin) and queues the rest of a
//      longer column as segments of kSeg non-zeros;
//   2. scatter (queued segments): one warp per segment, so the 10^5-long columns of a power-law
//      graph are spread over the whole grid instead of serialising one warp (or one grid-wide
//      pass per column, as the first version did: 29 us for ~80 long columns on C3);
//   3. compact: scan the accumulator, fold `zero`, apply the mask, reset touched entries,
//      emit with warp-aggregated atomics on y[0].index.
#include <math.h>
#include <string.h>

#include <vector>

#include "glb_internal.h"
#include "semiring.cuh"

namespace {

constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kSeg = 512;             // non-zeros per unit of work (one warp: 16 strides of 32)
constexpr float kFloatInf = 999999999.0f;  // FLOAT_INF, global.h:80 (spmspv_module.h:482-491)

template <int OP>
__device__ __forceinline__ float spmspv_mul(float a, float v) {
    if (OP == GLB_OP_ADD_MIN) {
        if (a > kFloatInf || v > kFloatInf) return kFloatInf;
        const float s = __fadd_rn(a, v);
        return s > kFloatInf ? kFloatInf : s;
    }
    return Semi<OP>::mul(a, v);
}

template <int OP>
__device__ __forceinline__ void combine(float *acc, float p) {
    if (OP == GLB_OP_MUL_ADD) {
        atomicAdd(acc, p);
    } else if (OP == GLB_OP_LOGICAL_AND_OR) {
        if (p != 0.0f) *acc = 1.0f;  // idempotent: racing writers store the same word
    } else {
        // float min through integer atomics: non-negative floats order like ints,
        // negative floats order inversely as unsigned
        if (p >= 0.0f) atomicMin(reinterpret_cast<int *>(acc), __float_as_int(p));
        else atomicMax(reinterpret_cast<unsigned *>(acc), __float_as_uint(p));
    }
}

struct SpmspvParams {
    const uint32_t *__restrict__ indptr;
    const uint32_t *__restrict__ indices;
    const float *__restrict__ vals;
    const glb_idx_val_t *__restrict__ x;
    const float *__restrict__ mask;
    glb_idx_val_t *y;
    float *acc;
    uint32_t *heavy;  // [0] = number of queued segments, [2 + 2i], [3 + 2i] = {frontier slot k, segment number}
    uint32_t row_begin, row_end;  // output rows of this shard: the compaction scans only these
    uint32_t heavy_cap;  // segments the queue holds (nnz / kSeg + 1: enough unless x repeats columns)
    uint32_t num_rows, num_cols;
    float zero;
    int mask_type;
};

// one warp, non-zeros [s, t) of one column, t - s <= kSeg
template <int OP>
__device__ __forceinline__ void scatter_span(const SpmspvParams &P, uint32_t s, uint32_t t, float v, unsigned lane) {
#pragma unroll 4
    for (uint32_t i = s + lane; i < t; i += 32)
        combine<OP>(P.acc + __ldg(P.indices + i), spmspv_mul<OP>(__ldg(P.vals + i), v));
}

template <int OP>
__global__ void __launch_bounds__(kThreads) spmspv_scatter_light(const SpmspvParams P) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * kThreads) >> 5;
    const uint32_t nnz_x = P.x[0].index;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.y[0].index = 0;
        P.y[0].val = P.zero;
    }
    for (uint32_t k = warp; k < nnz_x; k += n_warps) {
        const glb_idx_val_t e = P.x[k + 1];
        if (e.index >= P.num_cols) continue;  // not a column of this matrix: ignored (warp-uniform)
        const uint32_t s = __ldg(P.indptr + e.index), t = __ldg(P.indptr + e.index + 1);
        const uint32_t len = t - s;
        bool queued = false;
        if (len > kSeg) {  // queue segments 1 .. n_extra of this column: {frontier slot, segment}
            const uint32_t n_extra = (len - 1) / kSeg;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(P.heavy, n_extra);
            base = __shfl_sync(kFull, base, 0);
            // a frontier that lists columns more than once can outgrow the queue: then this warp walks
            // the whole column itself and voids what it reserved
            queued = base + n_extra <= P.heavy_cap;
            for (uint32_t j = lane; j < n_extra && base + j < P.heavy_cap; j += 32) {
                P.heavy[2 + 2 * (base + j)] = k;
                P.heavy[3 + 2 * (base + j)] = queued ? j + 1 : 0xffffffffu;
            }
        }
        if (queued) {
            scatter_span<OP>(P, s, s + kSeg, e.val, lane);
        } else {
            for (uint32_t b = s; b < t; b += kSeg) scatter_span<OP>(P, b, t - b > kSeg ? b + kSeg : t, e.val, lane);
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) spmspv_scatter_heavy(const SpmspvParams P) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * kThreads) >> 5;
    const uint32_t n_seg = P.heavy[0] < P.heavy_cap ? P.heavy[0] : P.heavy_cap;
    for (uint32_t h = warp; h < n_seg; h += n_warps) {
        const glb_idx_val_t e = P.x[P.heavy[2 + 2 * h] + 1];
        const uint32_t seg = P.heavy[3 + 2 * h];
        if (seg == 0xffffffffu) continue;  // voided reservation
        const uint32_t t = __ldg(P.indptr + e.index + 1);
        const uint32_t s = __ldg(P.indptr + e.index) + seg * kSeg;
        scatter_span<OP>(P, s, (t - s > kSeg) ? s + kSeg : t, e.val, lane);
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) spmspv_compact(const SpmspvParams P) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t n_threads = gridDim.x * kThreads;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.heavy[0] = 0;  // ready for the next run
    const uint32_t n_round = (P.row_end - P.row_begin + 31u) & ~31u;  // keep warps converged for the ballots
    for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q < n_round; q += n_threads) {
        const uint32_t r = P.row_begin + q;
        bool emit = false;
        float val = 0.0f;
        if (r < P.row_end) {
            const float a = P.acc[r];
            if (a != Semi<OP>::ident()) {
                P.acc[r] = Semi<OP>::ident();
                val = Semi<OP>::with_zero(P.zero, a);
                bool off = false;
                if (P.mask_type == GLB_MASK_WRITE_TO_ONE) off = (P.mask[r] == P.zero);
                else if (P.mask_type == GLB_MASK_WRITE_TO_ZERO) off = (P.mask[r] != P.zero);
                emit = !off && (val != P.zero);
            }
        }
        const unsigned b = __ballot_sync(kFull, emit);
        if (b) {
            uint32_t base = 0;
            const int leader = __ffs(int(b)) - 1;
            if (int(lane) == leader) base = atomicAdd(&P.y[0].index, uint32_t(__popc(b)));
            base = __shfl_sync(kFull, base, leader);
            if (emit) {
                glb_idx_val_t o;
                o.index = r;
                o.val = val;
                P.y[1 + base + __popc(b & ((1u << lane) - 1u))] = o;
            }
        }
    }
}

__global__ void fill_f32_kernel(float *dst, float val, size_t n) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = val;
}

// start vector of an app: constant but for one element (bfs.h:108-112, sssp.h:153-156)
__global__ void fill_one_f32_kernel(float *dst, float val, size_t n, size_t index, float index_val) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (i == index) ? index_val : val;
}

__global__ void sparse_scatter_kernel(const glb_idx_val_t *__restrict__ list, float *dense, uint32_t begin, uint32_t end) {
    const uint32_t n = list[0].index;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const glb_idx_val_t e = list[i + 1];
        if (e.index >= begin && e.index < end) dense[e.index] = e.val;
    }
}

// dense -> sparse list {count, zero}, entries != zero (the inverse of sparse_scatter_kernel): the
// row-sharded push direction exchanges its frontier as a dense vector and lists it again on every rank
__global__ void __launch_bounds__(kThreads) dense_to_sparse_kernel(const float *__restrict__ dense, uint32_t len, float zero,
                                                                 glb_idx_val_t *list) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t n_threads = gridDim.x * kThreads;
    const uint32_t n_round = (len + 31u) & ~31u;
    for (uint32_t r = blockIdx.x * kThreads + threadIdx.x; r < n_round; r += n_threads) {
        const float v = r < len ? dense[r] : zero;
        const bool emit = r < len && v != zero;
        const unsigned b = __ballot_sync(kFull, emit);
        if (b) {
            uint32_t base = 0;
            const int leader = __ffs(int(b)) - 1;
            if (int(lane) == leader) base = atomicAdd(&list[0].index, uint32_t(__popc(b)));
            base = __shfl_sync(kFull, base, leader);
            if (emit) {
                glb_idx_val_t o;
                o.index = r;
                o.val = v;
                list[1 + base + __popc(b & ((1u << lane) - 1u))] = o;
            }
        }
    }
}

__global__ void sparse_head_kernel(glb_idx_val_t *list, float zero) {
    list[0].index = 0;
    list[0].val = zero;
}

template <int OP>
int run_spmspv(glb_ctx_t ctx, glb_csc_t m, const SpmspvParams &P) {
    const int grid = ctx->num_sms * 8;
    spmspv_scatter_light<OP><<<grid, kThreads, 0, ctx->stream>>>(P);
    spmspv_scatter_heavy<OP><<<grid, kThreads, 0, ctx->stream>>>(P);
    spmspv_compact<OP><<<grid, kThreads, 0, ctx->stream>>>(P);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

}  // namespace

extern "C" {

int glb_buffer_fill_f32(glb_ctx_t ctx, float *dst_dev, float val, size_t n) {
    GLB_REQUIRE(ctx && (n == 0 || dst_dev), "NULL argument");
    if (n == 0) return GLB_OK;
    size_t blocks = (n + kThreads - 1) / kThreads;
    const size_t cap = size_t(ctx->num_sms) * 16;
    if (blocks > cap) blocks = cap;
    fill_f32_kernel<<<unsigned(blocks), kThreads, 0, ctx->stream>>>(dst_dev, val, n);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_buffer_fill_one_f32(glb_ctx_t ctx, float *dst_dev, float val, size_t n, size_t index, float index_val) {
    GLB_REQUIRE(ctx && (n == 0 || dst_dev), "NULL argument");
    GLB_REQUIRE(index < n, "index outside the vector");
    size_t blocks = (n + kThreads - 1) / kThreads;
    const size_t cap = size_t(ctx->num_sms) * 16;
    if (blocks > cap) blocks = cap;
    fill_one_f32_kernel<<<unsigned(blocks), kThreads, 0, ctx->stream>>>(dst_dev, val, n, index, index_val);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_csc_create(glb_ctx_t ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                   const float *data, glb_csc_t *out) {
    return glb_csc_create_rows(ctx, num_rows, num_cols, indptr, indices, data, 0, num_rows, out);
}

int glb_csc_create_rows(glb_ctx_t ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                        const float *data, uint32_t row_begin, uint32_t row_end, glb_csc_t *out) {
    GLB_REQUIRE(ctx && out && indptr, "NULL argument");
    GLB_REQUIRE(row_begin <= row_end && row_end <= num_rows, "bad row range");
    *out = nullptr;
    const uint64_t nnz_all = indptr[num_cols];
    GLB_REQUIRE(nnz_all == 0 || (indices && data), "NULL indices/data");
    for (uint32_t c = 0; c < num_cols; ++c) GLB_REQUIRE(indptr[c + 1] >= indptr[c], "indptr not monotone");
    for (uint64_t i = 0; i < nnz_all; ++i) GLB_REQUIRE(indices[i] < num_rows, "row index out of range");
    // a row shard keeps, of every column, the entries whose row it owns (the reference's row tiles,
    // data_formatter.h:621,665, with the tile = the GPU)
    std::vector<uint32_t> s_indptr, s_indices;
    std::vector<float> s_data;
    if (row_begin != 0 || row_end != num_rows) {
        s_indptr.assign(size_t(num_cols) + 1, 0);
        for (uint32_t c = 0; c < num_cols; ++c) {
            uint32_t kept = 0;
            for (uint32_t i = indptr[c]; i < indptr[c + 1]; ++i) kept += (indices[i] >= row_begin && indices[i] < row_end);
            s_indptr[c + 1] = s_indptr[c] + kept;
        }
        s_indices.resize(s_indptr[num_cols]);
        s_data.resize(s_indptr[num_cols]);
        for (uint32_t c = 0; c < num_cols; ++c) {
            uint32_t w = s_indptr[c];
            for (uint32_t i = indptr[c]; i < indptr[c + 1]; ++i)
                if (indices[i] >= row_begin && indices[i] < row_end) { s_indices[w] = indices[i]; s_data[w] = data[i]; ++w; }
        }
        indptr = s_indptr.data();
        indices = s_indices.data();
        data = s_data.data();
    }
    const uint64_t nnz = indptr[num_cols];
    GLB_CUDA(cudaSetDevice(ctx->device));
    glb_csc_t m = new glb_csc_s();
    m->ctx = ctx;
    glb_ctx_retain(ctx);
    m->num_rows = num_rows;
    m->num_cols = num_cols;
    m->row_begin = row_begin;
    m->row_end = row_end;
    m->nnz = nnz;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 4);
    };
    alloc(reinterpret_cast<void **>(&m->indptr), sizeof(uint32_t) * (size_t(num_cols) + 1));
    alloc(reinterpret_cast<void **>(&m->indices), sizeof(uint32_t) * nnz);
    alloc(reinterpret_cast<void **>(&m->vals), sizeof(float) * nnz);
    alloc(reinterpret_cast<void **>(&m->acc), sizeof(float) * num_rows);
    alloc(reinterpret_cast<void **>(&m->acc_inf), sizeof(float) * num_rows);
    alloc(reinterpret_cast<void **>(&m->counter), sizeof(uint32_t) * (2 * (size_t(nnz) / kSeg + 1) + 4));  // segment queue
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(m->indptr, indptr, sizeof(uint32_t) * (size_t(num_cols) + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz)
        e = cudaMemcpyAsync(m->indices, indices, sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz)
        e = cudaMemcpyAsync(m->vals, data, sizeof(float) * nnz, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->counter, 0, sizeof(uint32_t), ctx->stream);
    if (e == cudaSuccess && num_rows) {
        fill_f32_kernel<<<ctx->num_sms * 8, kThreads, 0, ctx->stream>>>(m->acc, 0.0f, num_rows);
        fill_f32_kernel<<<ctx->num_sms * 8, kThreads, 0, ctx->stream>>>(m->acc_inf, HUGE_VALF, num_rows);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        glb_set_error("glb_csc_create: %s", cudaGetErrorString(e));
        glb_csc_destroy(m);
        return e == cudaErrorMemoryAllocation ? GLB_ENOMEM : GLB_ECUDA;
    }
    *out = m;
    return GLB_OK;
}

int glb_csc_destroy(glb_csc_t m) {
    if (!m) return GLB_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->indptr); cudaFree(m->indices); cudaFree(m->vals); cudaFree(m->acc); cudaFree(m->acc_inf); cudaFree(m->counter);
    glb_ctx_release(m->ctx);
    delete m;
    return GLB_OK;
}

int glb_spmspv(glb_ctx_t ctx, glb_csc_t m, int op, float zero, int mask_type, const glb_idx_val_t *x, const float *mask,
               glb_idx_val_t *y) {
    GLB_REQUIRE(ctx && m && x && y, "NULL argument");
    GLB_REQUIRE(m->ctx == ctx, "matrix belongs to another context");
    GLB_REQUIRE(mask_type >= GLB_MASK_NONE && mask_type <= GLB_MASK_WRITE_TO_ONE, "invalid mask type");
    GLB_REQUIRE(mask_type == GLB_MASK_NONE || mask, "mask is NULL but mask_type != kNoMask");
    GLB_REQUIRE(static_cast<const void *>(x) != static_cast<const void *>(y), "y must not alias x");
    SpmspvParams P;
    P.indptr = m->indptr;
    P.indices = m->indices;
    P.vals = m->vals;
    P.x = x;
    P.mask = mask;
    P.y = y;
    // one accumulator per (+)-identity, both filled when the matrix was created and left at rest by
    // every run: no launch depends on host-side state (launch sequences can be recorded at any time)
    P.acc = (op == GLB_OP_ADD_MIN) ? m->acc_inf : m->acc;
    P.heavy = m->counter;
    P.heavy_cap = uint32_t(m->nnz / kSeg + 1);
    P.row_begin = m->row_begin;
    P.row_end = m->row_end;
    P.num_rows = m->num_rows;
    P.num_cols = m->num_cols;
    P.zero = zero;
    P.mask_type = mask_type;
    switch (op) {
        case GLB_OP_MUL_ADD: return run_spmspv<GLB_OP_MUL_ADD>(ctx, m, P);
        case GLB_OP_LOGICAL_AND_OR: return run_spmspv<GLB_OP_LOGICAL_AND_OR>(ctx, m, P);
        case GLB_OP_ADD_MIN: return run_spmspv<GLB_OP_ADD_MIN>(ctx, m, P);
    }
    glb_set_error("glb_spmspv: invalid semiring op %d", op);
    return GLB_EINVAL;
}

int glb_sparse_count(glb_ctx_t ctx, const glb_idx_val_t *list, uint32_t *count) {
    GLB_REQUIRE(ctx && list && count, "NULL argument");
    glb_idx_val_t head;
    GLB_CUDA(cudaMemcpyAsync(&head, list, sizeof(head), cudaMemcpyDeviceToHost, ctx->stream));
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = head.index;
    return GLB_OK;
}

int glb_sparse_to_dense(glb_ctx_t ctx, const glb_idx_val_t *list, float *dense, uint32_t len, float zero) {
    GLB_REQUIRE(ctx && list && (len == 0 || dense), "NULL argument");
    if (len == 0) return GLB_OK;
    int rc = glb_buffer_fill_f32(ctx, dense, zero, len);
    if (rc) return rc;
    sparse_scatter_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(list, dense, 0u, len);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

// Row-sharded push: the frontier travels as a dense vector.  glb_sparse_to_dense_rows resets rows
// [row_begin, row_end) of `dense` to `zero` and scatters the list entries that fall inside;
// glb_dense_to_sparse lists the entries != zero of a dense vector ({count, zero} head, order unspecified).
int glb_sparse_to_dense_rows(glb_ctx_t ctx, const glb_idx_val_t *list, float *dense, uint32_t row_begin, uint32_t row_end,
                             float zero) {
    GLB_REQUIRE(ctx && list && dense && row_begin <= row_end, "bad argument");
    if (row_begin == row_end) return GLB_OK;
    int rc = glb_buffer_fill_f32(ctx, dense + row_begin, zero, row_end - row_begin);
    if (rc) return rc;
    sparse_scatter_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(list, dense, row_begin, row_end);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_dense_to_sparse(glb_ctx_t ctx, const float *dense, uint32_t len, float zero, glb_idx_val_t *list) {
    GLB_REQUIRE(ctx && list && (len == 0 || dense), "NULL argument");
    sparse_head_kernel<<<1, 1, 0, ctx->stream>>>(list, zero);
    if (len) dense_to_sparse_kernel<<<ctx->num_sms * 4, kThreads, 0, ctx->stream>>>(dense, len, zero, list);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

}  // extern "C"
