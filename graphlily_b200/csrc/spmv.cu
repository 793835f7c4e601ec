// SpMV: y = A (+).(x) x over a row shard in warp-segment layout (overlay mode 1).
//
// Replaces kernel_spmv (/root/reference/graphlily/hw/kernel_spmv_impl.h:392-819) and the
// CPSR formatter that feeds it (graphlily/io/data_formatter.h:457-534).  Semantics are those of
// SpMVModule::compute_reference_results (graphlily/module/spmv_module.h:478-532).
//
// Layout and schedule (see glb_internal.h for the arrays):
//   * the shard's nnz stream is cut into chunks of GLB_CHUNK = 1024 non-zeros, one warp each --
//     perfect nnz balance whatever the row-length distribution (power-law rows of 1 .. 2^20);
//   * per step a warp streams 128 non-zeros: lane l loads cols/vals [4l, 4l+4) with one
//     128-bit streaming load each (fully coalesced 512-B rows), gathers x, multiplies;
//   * row boundaries travel in bit 31 of the column word; a ballot/popc prefix gives every
//     row end its ordinal, a 5-step shuffle segmented scan gives the carry into each lane;
//   * finished rows are staged in shared memory by ordinal and written back by consecutive
//     lanes (coalesced mask reads / y writes), with the mask and the fused eWiseAdd / dense
//     assign epilogues applied in the same pass;
//   * rows that cross or touch a chunk boundary leave per-chunk head / tail carries that the
//     fix-up kernel combines (deterministic -- no float atomics); it also writes empty rows.
//
// Algorithmic HBM bytes per launch: 8*nnz (cols+vals) + 4*nonempty_rows (nz_rows) + 4*ncols (x)
// + 4*rows (y) [+ 4*rows mask], i.e. the CSR figure of SURVEY.md section 8d.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "glb_internal.h"
#include "semiring.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kStep = 128;  // non-zeros per warp step
constexpr unsigned kFull = 0xffffffffu;

struct SpmvParams {
    const uint32_t *__restrict__ cols;
    const float *__restrict__ vals;
    const uint32_t *__restrict__ nz_rows;
    const uint32_t *__restrict__ chunk_first;
    const float *__restrict__ x;   // gather source: the caller's x, or its relabelled copy xp
    const float *mask;  // may alias assign_inout
    float *y;
    float *head_carry;
    float *tail_carry;
    uint64_t nnz;
    uint32_t n_chunks;
    float zero;
    int mask_type;
    int add_enable;
    float add_val;
    float *assign_inout;
    float assign_val;
    int assign_mask_type;
    // fix-up lists
    const glb_fixup_t *fix_short;
    const glb_fixup_t *fix_long;
    const uint32_t *empty_rows;
    uint32_t n_fix_short, n_fix_long, n_empty;
    uint32_t tile_k;  // columns [0, tile_k) of the relabelled x live in shared memory (0 = none)
};

// Row write-back: fold `zero`, apply the mask (literal 0 compare / literal 0 write,
// spmv_module.h:513-532), then the optional fused eWiseAdd and dense assign.
template <int OP>
__device__ __forceinline__ void finish_row(const SpmvParams &P, uint32_t row, float total) {
    float v = Semi<OP>::with_zero(P.zero, total);
    if (P.mask_type == GLB_MASK_WRITE_TO_ZERO) {
        if (P.mask[row] != 0.0f) v = 0.0f;
    } else if (P.mask_type == GLB_MASK_WRITE_TO_ONE) {
        if (P.mask[row] == 0.0f) v = 0.0f;
    }
    if (P.add_enable) v = __fadd_rn(v, P.add_val);
    P.y[row] = v;
    if (P.assign_inout) {
        bool hit = (P.assign_mask_type == GLB_MASK_WRITE_TO_ONE) ? (v != 0.0f) : (v == 0.0f);
        if (hit) P.assign_inout[row] = P.assign_val;
    }
}

// One chunk (GLB_CHUNK non-zeros) by one warp.  TILE: columns below P.tile_k are read from the
// shared-memory copy of the hot end of the relabelled vector, the rest through L1/L2.
template <int OP, bool TILE>
__device__ __forceinline__ void process_chunk(const SpmvParams &P, const uint32_t chunk, const unsigned lane,
                                              float *my_stage, const float *tile) {
    const uint32_t cf = P.chunk_first[chunk];
    const uint32_t ord0 = cf & ~GLB_FLAG;
    const bool fresh = (cf & GLB_FLAG) != 0;
    const uint64_t base = uint64_t(chunk) * GLB_CHUNK;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned le_mask = lt_mask | (1u << lane);

    uint32_t ord_base = ord0;  // ordinal of the row open at the start of the step
    float wcarry = Semi<OP>::ident();

#pragma unroll 2
    for (int it = 0; it < int(GLB_CHUNK / kStep); ++it) {
        const uint64_t p = base + uint64_t(it) * kStep + lane * 4u;
        const uint4 c4 = __ldcs(reinterpret_cast<const uint4 *>(P.cols + p));
        const float4 a4 = __ldcs(reinterpret_cast<const float4 *>(P.vals + p));
        const uint32_t c[4] = {c4.x, c4.y, c4.z, c4.w};
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        // padding beyond nnz (last chunk only) contributes the identity
        const uint64_t left = (P.nnz > p) ? (P.nnz - p) : 0;
        const int rem = left > 4 ? 4 : int(left);

        bool f[4];
        float prod[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[j] = (c[j] & GLB_FLAG) != 0;
            const uint32_t col = c[j] & ~GLB_FLAG;
            float xv;
            if (TILE) xv = (col < P.tile_k) ? tile[col] : __ldg(P.x + col);
            else xv = __ldg(P.x + col);
            prod[j] = (j < rem) ? Semi<OP>::mul(a[j], xv) : Semi<OP>::ident();
        }

        // pass 1: value of the lane's open tail segment, flag census
        float tail = Semi<OP>::ident();
        bool hasf = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (f[j]) { tail = Semi<OP>::ident(); hasf = true; }
            tail = Semi<OP>::add(tail, prod[j]);
        }
        unsigned excl = 0, total = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned b = __ballot_sync(kFull, f[j]);
            excl += __popc(b & lt_mask);
            total += __popc(b);
        }
        const unsigned hb = __ballot_sync(kFull, hasf);

        // segmented inclusive scan of the tails across lanes (segments start at flagged lanes)
        if (lane == 0 && !hasf) tail = Semi<OP>::add(wcarry, tail);
        int start = 31 - __clz(int(hb & le_mask));
        if (start < 0) start = 0;
        float v = tail;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float tv = __shfl_up_sync(kFull, v, d);
            if (int(lane) - d >= start) v = Semi<OP>::add(tv, v);
        }
        float carry_in = __shfl_up_sync(kFull, v, 1);
        if (lane == 0) carry_in = wcarry;
        wcarry = __shfl_sync(kFull, v, 31);

        // pass 2: close rows; the k-th row end of this step goes to stage[k]
        float acc = carry_in;
        unsigned k = excl;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (f[j]) {
                my_stage[k] = acc;
                ++k;
                acc = Semi<OP>::ident();
            }
            acc = Semi<OP>::add(acc, prod[j]);
        }
        __syncwarp();
        for (unsigned k2 = lane; k2 < total; k2 += 32) {
            const float val = my_stage[k2];
            const uint32_t ord = ord_base + k2;
            if (ord == ord0 && !fresh) {
                P.head_carry[chunk] = val;  // row began in an earlier chunk
            } else {
                finish_row<OP>(P, P.nz_rows[ord], val);
            }
        }
        __syncwarp();
        ord_base += total;
    }
    if (lane == 0) P.tail_carry[chunk] = wcarry;
}

// Variant A: one chunk per warp, x gathered through L1/L2 only.
template <int OP>
__global__ void __launch_bounds__(kThreads) spmv_ws_kernel(const SpmvParams P) {
    __shared__ float stage[kWarpsPerBlock][kStep];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wib = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kWarpsPerBlock + wib;
    if (chunk >= P.n_chunks) return;  // warp-uniform; no block-wide barrier below
    process_chunk<OP, false>(P, chunk, lane, stage[wib], nullptr);
}

// ---- TMA (bulk async copy) + mbarrier helpers -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!done);
}

// Variant B: persistent CTAs, one per SM.  Each CTA pulls the hot end of the relabelled vector
// (columns [0, tile_k), the most referenced ones) into shared memory with TMA bulk copies
// once, then its warps walk the chunk list; hot gathers become bank-parallel LDS instead of
// one L1 wavefront + one 32-byte L2 sector per lane.
template <int OP>
__global__ void __launch_bounds__(1024, 1) spmv_ws_tile_kernel(const SpmvParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);
    const unsigned n_warps = blockDim.x >> 5;
    float *stage = tile + P.tile_k;
    uint64_t *bar = reinterpret_cast<uint64_t *>(stage + n_warps * kStep);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wib = threadIdx.x >> 5;

    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = P.tile_k * 4u;
        mbar_expect_tx(bar, bytes);
        constexpr uint32_t kPiece = 32768;
        for (uint32_t off = 0; off < bytes; off += kPiece) {
            const uint32_t nb = (bytes - off < kPiece) ? (bytes - off) : kPiece;
            bulk_g2s(reinterpret_cast<unsigned char *>(tile) + off, reinterpret_cast<const unsigned char *>(P.x) + off,
                     nb, bar);
        }
    }
    mbar_wait(bar, 0);

    const uint32_t total_warps = gridDim.x * n_warps;
    for (uint32_t chunk = blockIdx.x * n_warps + wib; chunk < P.n_chunks; chunk += total_warps)
        process_chunk<OP, true>(P, chunk, lane, stage + wib * kStep, tile);
}

// xp[i] = x[perm[i]]: the caller's vector in the matrix's relabelled column order.
__global__ void __launch_bounds__(kThreads) permute_x_kernel(const float *__restrict__ x,
                                                           const uint32_t *__restrict__ perm, float *__restrict__ xp,
                                                           uint32_t n4) {
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(perm) + i);
        float4 o;
        o.x = __ldg(x + q.x);
        o.y = __ldg(x + q.y);
        o.z = __ldg(x + q.z);
        o.w = __ldg(x + q.w);
        reinterpret_cast<float4 *>(xp)[i] = o;
    }
}

// Rows touching a chunk boundary: total = tail[c_begin .. c_last] (+) head[c_end].
template <int OP>
__global__ void __launch_bounds__(kThreads) spmv_fixup_kernel(const SpmvParams P, uint32_t nb_short, uint32_t nb_long) {
    const unsigned lane = threadIdx.x & 31u;
    if (blockIdx.x < nb_short) {
        const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
        if (i >= P.n_fix_short) return;
        const glb_fixup_t e = P.fix_short[i];
        const uint32_t c_end = e.c_end & ~GLB_FLAG;
        const bool has_head = (e.c_end & GLB_FLAG) != 0;
        const uint32_t c_stop = has_head ? c_end : c_end + 1;  // exclusive end of the tail range
        float t = Semi<OP>::ident();
        for (uint32_t c = e.c_begin; c < c_stop; ++c) t = Semi<OP>::add(t, P.tail_carry[c]);
        if (has_head) t = Semi<OP>::add(t, P.head_carry[c_end]);
        finish_row<OP>(P, e.row, t);
    } else if (blockIdx.x < nb_short + nb_long) {
        const uint32_t i = (blockIdx.x - nb_short) * kWarpsPerBlock + (threadIdx.x >> 5);
        if (i >= P.n_fix_long) return;
        const glb_fixup_t e = P.fix_long[i];
        const uint32_t c_end = e.c_end & ~GLB_FLAG;
        const bool has_head = (e.c_end & GLB_FLAG) != 0;
        const uint32_t c_stop = has_head ? c_end : c_end + 1;
        float t = Semi<OP>::ident();
        for (uint32_t c = e.c_begin + lane; c < c_stop; c += 32) t = Semi<OP>::add(t, P.tail_carry[c]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) t = Semi<OP>::add(t, __shfl_xor_sync(kFull, t, d));
        if (lane == 0) {
            if (has_head) t = Semi<OP>::add(t, P.head_carry[c_end]);
            finish_row<OP>(P, e.row, t);
        }
    } else {
        const uint32_t i = (blockIdx.x - nb_short - nb_long) * kThreads + threadIdx.x;
        if (i >= P.n_empty) return;
        finish_row<OP>(P, P.empty_rows[i], Semi<OP>::ident());
    }
}

template <int OP>
int launch_op(glb_ctx_t ctx, glb_csr_t m, const SpmvParams &P) {
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (ctx->timing) {
        for (auto &e : ev) GLB_CUDA(cudaEventCreate(&e));
        GLB_CUDA(cudaEventRecord(ev[0], ctx->stream));
    }
    if (P.n_chunks && P.tile_k) {
        const uint32_t n_warps = m->tile_threads / 32;
        const size_t smem = size_t(P.tile_k) * 4 + size_t(n_warps) * kStep * 4 + 16;
        static thread_local size_t attr_set[3] = {0, 0, 0};
        if (attr_set[OP] < smem) {
            GLB_CUDA(cudaFuncSetAttribute(spmv_ws_tile_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            attr_set[OP] = smem;
        }
        uint32_t grid = (P.n_chunks + n_warps - 1) / n_warps;
        if (grid > uint32_t(ctx->num_sms)) grid = uint32_t(ctx->num_sms);
        spmv_ws_tile_kernel<OP><<<grid, m->tile_threads, smem, ctx->stream>>>(P);
    } else if (P.n_chunks) {
        const uint32_t grid = (P.n_chunks + kWarpsPerBlock - 1) / kWarpsPerBlock;
        spmv_ws_kernel<OP><<<grid, kThreads, 0, ctx->stream>>>(P);
    }
    if (ctx->timing) GLB_CUDA(cudaEventRecord(ev[1], ctx->stream));
    const uint32_t nb_short = (P.n_fix_short + kThreads - 1) / kThreads;
    const uint32_t nb_long = (P.n_fix_long + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const uint32_t nb_empty = (P.n_empty + kThreads - 1) / kThreads;
    if (nb_short + nb_long + nb_empty)
        spmv_fixup_kernel<OP><<<nb_short + nb_long + nb_empty, kThreads, 0, ctx->stream>>>(P, nb_short, nb_long);
    if (ctx->timing) {
        GLB_CUDA(cudaEventRecord(ev[2], ctx->stream));
        for (auto e : ev) ctx->timing_events.push_back(e);
    }
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

template <typename T>
int upload(glb_ctx_t ctx, T **dptr, const T *host, size_t n, size_t n_alloc, size_t *bytes) {
    *dptr = nullptr;
    if (n_alloc == 0) n_alloc = 1;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(dptr), n_alloc * sizeof(T));
    if (e != cudaSuccess) {
        glb_set_error("cudaMalloc(%zu): %s", n_alloc * sizeof(T), cudaGetErrorString(e));
        return GLB_ENOMEM;
    }
    *bytes += n_alloc * sizeof(T);
    if (n_alloc > n) GLB_CUDA(cudaMemsetAsync(*dptr + n, 0, (n_alloc - n) * sizeof(T), ctx->stream));
    if (n) GLB_CUDA(cudaMemcpyAsync(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return GLB_OK;
}

}  // namespace

int glb_launch_spmv(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x, const float *mask,
                    float *y, const glb_spmv_epilogue_t *ep) {
    SpmvParams P;
    memset(&P, 0, sizeof(P));
    P.cols = m->cols;
    P.vals = m->vals;
    P.nz_rows = m->nz_rows;
    P.chunk_first = m->chunk_first;
    P.x = x;
    if (m->col_perm) {  // relabelled columns: gather from the permuted copy of x
        const uint32_t n4 = (m->num_cols + 3) / 4;
        uint32_t blocks = (n4 + kThreads - 1) / kThreads;
        const uint32_t cap = uint32_t(ctx->num_sms) * 8;
        if (blocks > cap) blocks = cap;
        permute_x_kernel<<<blocks, kThreads, 0, ctx->stream>>>(x, m->col_perm, m->xp, n4);
        P.x = m->xp;
    }
    P.tile_k = m->tile_k;
    P.mask = mask;
    P.y = y;
    P.head_carry = m->head_carry;
    P.tail_carry = m->tail_carry;
    P.nnz = m->nnz;
    P.n_chunks = m->n_chunks;
    P.zero = zero;
    P.mask_type = mask_type;
    if (ep) {
        P.add_enable = ep->add_enable;
        P.add_val = ep->add_val;
        P.assign_inout = ep->assign_inout;
        P.assign_val = ep->assign_val;
        P.assign_mask_type = ep->assign_mask_type;
    }
    P.fix_short = m->fix_short;
    P.fix_long = m->fix_long;
    P.empty_rows = m->empty_rows;
    P.n_fix_short = m->n_fix_short;
    P.n_fix_long = m->n_fix_long;
    P.n_empty = m->n_empty;
    switch (op) {
        case GLB_OP_MUL_ADD: return launch_op<GLB_OP_MUL_ADD>(ctx, m, P);
        case GLB_OP_LOGICAL_AND_OR: return launch_op<GLB_OP_LOGICAL_AND_OR>(ctx, m, P);
        case GLB_OP_ADD_MIN: return launch_op<GLB_OP_ADD_MIN>(ctx, m, P);
    }
    glb_set_error("glb_spmv: invalid semiring op %d", op);
    return GLB_EINVAL;
}

// Host side of the formatter: O(nnz) copy + O(rows) flagging.  Shared by glb_csr_create and
// the host-only glb_csr_format_host (which lets the layout be checked without a GPU).
struct HostLayout {
    uint32_t *cols = nullptr;  // malloc'd, nnz entries (padding is added on the device)
    std::vector<uint32_t> col_perm;  // relabelled column -> original column (empty = identity)
    std::vector<uint32_t> nz_rows, empty_rows, chunk_first;
    std::vector<glb_fixup_t> fix_short, fix_long;
    uint64_t nnz = 0, sb = 0;
    uint32_t n_chunks = 0;
    ~HostLayout() { free(cols); }
};

static int format_host(uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                       uint32_t row_begin, uint32_t row_end, HostLayout &L, bool relabel = false) {
    GLB_REQUIRE(indptr, "NULL indptr");
    GLB_REQUIRE(row_begin <= row_end && row_end <= num_rows, "bad row range");
    GLB_REQUIRE(num_cols >= 1 && num_cols < GLB_FLAG, "num_cols must be in [1, 2^31)");
    const uint64_t sb = indptr[row_begin];
    GLB_REQUIRE(indptr[row_end] >= sb, "indptr not monotone");
    const uint64_t nnz = uint64_t(indptr[row_end]) - sb;
    GLB_REQUIRE(nnz == 0 || indices, "NULL indices");
    L.nnz = nnz;
    L.sb = sb;
    L.n_chunks = uint32_t((nnz + GLB_CHUNK - 1) / GLB_CHUNK);
    L.cols = static_cast<uint32_t *>(malloc(sizeof(uint32_t) * (nnz ? nnz : 1)));
    if (!L.cols) { glb_set_error("glb_csr_create: host allocation failed"); return GLB_ENOMEM; }
    if (nnz) memcpy(L.cols, indices + sb, sizeof(uint32_t) * nnz);
    for (uint64_t i = 0; i < nnz; ++i) {
        if (L.cols[i] >= num_cols) {
            glb_set_error("glb_csr_create: column index %u out of range at nnz %llu", L.cols[i],
                          (unsigned long long)(sb + i));
            return GLB_EINVAL;
        }
    }
    if (relabel && nnz) {
        // Columns renumbered by descending reference count inside this shard (ties by id): the hot
        // end of the vector becomes one contiguous block -- a TMA-loadable shared-memory tile -- and
        // equally warm columns share 32-byte sectors.  col_perm[new] = old.
        std::vector<uint32_t> count(num_cols, 0);
        for (uint64_t i = 0; i < nnz; ++i) count[L.cols[i]]++;
        L.col_perm.resize(num_cols);
        for (uint32_t c = 0; c < num_cols; ++c) L.col_perm[c] = c;
        std::stable_sort(L.col_perm.begin(), L.col_perm.end(),
                         [&](uint32_t a, uint32_t b) { return count[a] > count[b]; });
        std::vector<uint32_t> &rank = count;  // reuse: rank[old] = new
        for (uint32_t n = 0; n < num_cols; ++n) rank[L.col_perm[n]] = n;
        for (uint64_t i = 0; i < nnz; ++i) L.cols[i] = rank[L.cols[i]];
    }
    L.chunk_first.assign(L.n_chunks, 0);
    L.nz_rows.reserve(row_end - row_begin);
    for (uint32_t r = row_begin; r < row_end; ++r) {
        if (indptr[r + 1] < indptr[r]) {
            glb_set_error("glb_csr_create: indptr not monotone at row %u", r);
            return GLB_EINVAL;
        }
        const uint64_t p0 = uint64_t(indptr[r]) - sb, p1 = uint64_t(indptr[r + 1]) - sb;
        if (p0 == p1) { L.empty_rows.push_back(r); continue; }
        const uint32_t ord = uint32_t(L.nz_rows.size());
        L.nz_rows.push_back(r);
        const uint32_t c_s = uint32_t(p0 / GLB_CHUNK), c_e = uint32_t((p1 - 1) / GLB_CHUNK);
        if (p0 % GLB_CHUNK == 0) L.chunk_first[c_s] = ord | GLB_FLAG;  // row starts with the chunk
        else L.cols[p0] |= GLB_FLAG;
        for (uint32_t c = c_s + 1; c <= c_e; ++c) L.chunk_first[c] = ord;  // chunks starting inside this row
        const bool ends_on_boundary = (p1 % GLB_CHUNK == 0) || (p1 == nnz);
        if (c_s == c_e && !ends_on_boundary) continue;  // interior row: the main kernel finishes it
        glb_fixup_t e;
        e.row = r;
        e.c_begin = c_s;
        e.c_end = ends_on_boundary ? c_e : (c_e | GLB_FLAG);
        ((c_e - c_s) <= 32 ? L.fix_short : L.fix_long).push_back(e);
    }
    return GLB_OK;
}

extern "C" {

// Host-only view of the layout for tests (no CUDA call).  Arrays are malloc'd copies the
// caller releases with glb_host_layout_free.
int glb_csr_format_host(uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                        uint32_t row_begin, uint32_t row_end, glb_host_layout_t *out) {
    GLB_REQUIRE(out, "out is NULL");
    memset(out, 0, sizeof(*out));
    HostLayout L;
    int rc = format_host(num_rows, num_cols, indptr, indices, row_begin, row_end, L);
    if (rc) return rc;
    auto dup = [](const void *src, size_t bytes) {
        void *p = malloc(bytes ? bytes : 1);
        if (p && bytes) memcpy(p, src, bytes);
        return p;
    };
    std::vector<glb_fixup_t> fix(L.fix_short);
    fix.insert(fix.end(), L.fix_long.begin(), L.fix_long.end());
    out->chunk = GLB_CHUNK;
    out->nnz = L.nnz;
    out->n_chunks = L.n_chunks;
    out->n_nz_rows = uint32_t(L.nz_rows.size());
    out->n_empty = uint32_t(L.empty_rows.size());
    out->n_fixups = uint32_t(fix.size());
    out->cols = static_cast<uint32_t *>(dup(L.cols, sizeof(uint32_t) * L.nnz));
    out->nz_rows = static_cast<uint32_t *>(dup(L.nz_rows.data(), sizeof(uint32_t) * L.nz_rows.size()));
    out->empty_rows = static_cast<uint32_t *>(dup(L.empty_rows.data(), sizeof(uint32_t) * L.empty_rows.size()));
    out->chunk_first = static_cast<uint32_t *>(dup(L.chunk_first.data(), sizeof(uint32_t) * L.chunk_first.size()));
    out->fixups = static_cast<uint32_t *>(dup(fix.data(), sizeof(glb_fixup_t) * fix.size()));
    return GLB_OK;
}

int glb_host_layout_free(glb_host_layout_t *l) {
    if (!l) return GLB_OK;
    free(l->cols); free(l->nz_rows); free(l->empty_rows); free(l->chunk_first); free(l->fixups);
    memset(l, 0, sizeof(*l));
    return GLB_OK;
}

int glb_csr_create(glb_ctx_t ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                   const float *data, uint32_t row_begin, uint32_t row_end, glb_csr_t *out) {
    GLB_REQUIRE(ctx && out, "NULL argument");
    *out = nullptr;
    // Tuning knobs (defaults chosen by measurement on B200):
    //   GLB_SPMV_RELABEL=0|1       renumber columns by popularity (needs one x-permute pass per SpMV)
    //   GLB_SPMV_TILE_K=<floats>   hot columns kept in shared memory (0 = gather everything via L1/L2)
    //   GLB_SPMV_TILE_THREADS=<n>  threads of the persistent tile kernel
    auto env_u32 = [](const char *name, uint32_t dflt) {
        const char *v = getenv(name);
        return v ? uint32_t(strtoul(v, nullptr, 10)) : dflt;
    };
    const bool relabel = env_u32("GLB_SPMV_RELABEL", 1) != 0 && num_cols >= 4096;
    uint32_t tile_k = relabel ? env_u32("GLB_SPMV_TILE_K", 0) : 0;  // measured: the 1-CTA/SM tile kernel loses (profiles/r1_sweep_v1.txt)
    uint32_t tile_threads = env_u32("GLB_SPMV_TILE_THREADS", 1024);
    if (tile_threads < 32 || tile_threads > 1024 || tile_threads % 32) tile_threads = 1024;
    const uint32_t cols_pad = (num_cols + 3u) & ~3u;
    if (tile_k > cols_pad) tile_k = cols_pad;
    tile_k &= ~3u;
    if (size_t(tile_k) * 4 + size_t(tile_threads / 32) * kStep * 4 + 16 > 227 * 1024)
        tile_k = uint32_t((227 * 1024 - size_t(tile_threads / 32) * kStep * 4 - 16) / 4) & ~3u;

    HostLayout L;
    int rc = format_host(num_rows, num_cols, indptr, indices, row_begin, row_end, L, relabel);
    if (rc) return rc;
    const uint64_t nnz = L.nnz, sb = L.sb;
    GLB_REQUIRE(nnz == 0 || data, "NULL data");
    GLB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t n_chunks = L.n_chunks;
    const uint64_t nnz_pad = uint64_t(n_chunks) * GLB_CHUNK;

    glb_csr_t m = new glb_csr_s();
    m->ctx = ctx;
    m->num_rows = num_rows;
    m->num_cols = num_cols;
    m->row_begin = row_begin;
    m->row_end = row_end;
    m->nnz = nnz;
    m->n_chunks = n_chunks;
    m->n_nz_rows = uint32_t(L.nz_rows.size());
    m->n_fix_short = uint32_t(L.fix_short.size());
    m->n_fix_long = uint32_t(L.fix_long.size());
    m->n_empty = uint32_t(L.empty_rows.size());

    size_t bytes = 0;
    if (!rc) rc = upload(ctx, &m->cols, L.cols, size_t(nnz), size_t(nnz_pad), &bytes);
    if (!rc) rc = upload(ctx, &m->vals, nnz ? data + sb : nullptr, size_t(nnz), size_t(nnz_pad), &bytes);
    if (!rc) rc = upload(ctx, &m->nz_rows, L.nz_rows.data(), L.nz_rows.size(), L.nz_rows.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->chunk_first, L.chunk_first.data(), L.chunk_first.size(), L.chunk_first.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->fix_short, L.fix_short.data(), L.fix_short.size(), L.fix_short.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->fix_long, L.fix_long.data(), L.fix_long.size(), L.fix_long.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->empty_rows, L.empty_rows.data(), L.empty_rows.size(), L.empty_rows.size(), &bytes);
    if (!rc && !L.col_perm.empty()) {
        rc = upload(ctx, &m->col_perm, L.col_perm.data(), L.col_perm.size(), size_t(cols_pad), &bytes);
        if (!rc) rc = upload<float>(ctx, &m->xp, nullptr, 0, size_t(cols_pad), &bytes);
        m->tile_k = tile_k;
        m->tile_threads = tile_threads;
    }
    if (!rc) rc = upload<float>(ctx, &m->head_carry, nullptr, 0, n_chunks, &bytes);
    if (!rc) rc = upload<float>(ctx, &m->tail_carry, nullptr, 0, n_chunks, &bytes);
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        glb_set_error("glb_csr_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = GLB_ECUDA;
    }
    m->device_bytes = bytes;
    if (rc) { glb_csr_destroy(m); return rc; }
    *out = m;
    return GLB_OK;
}

int glb_csr_destroy(glb_csr_t m) {
    if (!m) return GLB_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->cols); cudaFree(m->vals); cudaFree(m->nz_rows); cudaFree(m->chunk_first);
    cudaFree(m->fix_short); cudaFree(m->fix_long); cudaFree(m->empty_rows);
    cudaFree(m->head_carry); cudaFree(m->tail_carry); cudaFree(m->col_perm); cudaFree(m->xp);
    cudaFree(m->dx); cudaFree(m->dmask); cudaFree(m->dy);
    delete m;
    return GLB_OK;
}

int glb_csr_info(glb_csr_t m, uint64_t info[8]) {
    GLB_REQUIRE(m && info, "NULL argument");
    info[0] = m->row_end - m->row_begin;
    info[1] = m->num_cols;
    info[2] = m->nnz;
    info[3] = m->n_chunks;
    info[4] = uint64_t(m->n_fix_short) + m->n_fix_long;
    info[5] = m->n_empty;
    info[6] = m->device_bytes;
    info[7] = GLB_CHUNK;
    return GLB_OK;
}

static int check_spmv_args(glb_ctx_t ctx, glb_csr_t m, int op, int mask_type, const float *x, const float *mask,
                           float *y, const glb_spmv_epilogue_t *ep) {
    GLB_REQUIRE(ctx && m && x && y, "NULL argument");
    GLB_REQUIRE(m->ctx == ctx, "matrix belongs to another context");
    GLB_REQUIRE(op >= GLB_OP_MUL_ADD && op <= GLB_OP_ADD_MIN, "invalid semiring op");
    GLB_REQUIRE(mask_type >= GLB_MASK_NONE && mask_type <= GLB_MASK_WRITE_TO_ONE, "invalid mask type");
    GLB_REQUIRE(mask_type == GLB_MASK_NONE || mask, "mask is NULL but mask_type != kNoMask");
    GLB_REQUIRE(x != y, "y must not alias x");
    if (ep && ep->assign_inout)
        GLB_REQUIRE(ep->assign_mask_type == GLB_MASK_WRITE_TO_ZERO || ep->assign_mask_type == GLB_MASK_WRITE_TO_ONE,
                    "fused dense assign needs a mask type");
    return GLB_OK;
}

int glb_spmv_fused(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x, const float *mask,
                   float *y, const glb_spmv_epilogue_t *ep) {
    int rc = check_spmv_args(ctx, m, op, mask_type, x, mask, y, ep);
    if (rc) return rc;
    return glb_launch_spmv(ctx, m, op, zero, mask_type, x, mask, y, ep);
}

int glb_spmv(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x, const float *mask,
             float *y) {
    return glb_spmv_fused(ctx, m, op, zero, mask_type, x, mask, y, nullptr);
}

int glb_spmv_host(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x_host,
                  const float *mask_host, float *y_host) {
    GLB_REQUIRE(ctx && m && x_host && y_host, "NULL argument");
    GLB_REQUIRE(m->ctx == ctx, "matrix belongs to another context");
    GLB_REQUIRE(mask_type == GLB_MASK_NONE || mask_host, "mask is NULL but mask_type != kNoMask");
    GLB_CUDA(cudaSetDevice(ctx->device));
    if (!m->dx) GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->dx), sizeof(float) * m->num_cols));
    if (!m->dy) GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->dy), sizeof(float) * m->num_rows));
    if (mask_type != GLB_MASK_NONE && !m->dmask)
        GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->dmask), sizeof(float) * m->num_rows));
    GLB_CUDA(cudaMemcpyAsync(m->dx, x_host, sizeof(float) * m->num_cols, cudaMemcpyHostToDevice, ctx->stream));
    if (mask_type != GLB_MASK_NONE)
        GLB_CUDA(cudaMemcpyAsync(m->dmask, mask_host, sizeof(float) * m->num_rows, cudaMemcpyHostToDevice, ctx->stream));
    int rc = glb_spmv(ctx, m, op, zero, mask_type, m->dx, m->dmask, m->dy);
    if (rc) return rc;
    const size_t nr = size_t(m->row_end - m->row_begin);
    GLB_CUDA(cudaMemcpyAsync(y_host + m->row_begin, m->dy + m->row_begin, sizeof(float) * nr, cudaMemcpyDeviceToHost,
                             ctx->stream));
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GLB_OK;
}

}  // extern "C"
