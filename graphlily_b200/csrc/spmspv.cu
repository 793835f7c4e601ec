// SpMSpV: y = A[:, idx(x)] (+).(x) x, CSC, sparse in -> sparse out (overlay mode 2).
//
// Replaces kernel_spmspv (/root/reference/graphlily/hw/kernel_spmspv_impl.h:448-562) and
// formatCSC (graphlily/io/data_formatter.h:608-721).  Semantics: SpMSpVModule::
// compute_reference_results (graphlily/module/spmspv_module.h:446-520) followed by the device
// write-back rule (kernel_spmspv_impl.h:200-229,263-281,551-555): only entries != zero that
// pass the mask are listed, the head slot is {count, zero}, order is unspecified.
//
// ONE persistent launch (4 CTAs per SM, all co-resident) with a grid-wide barrier between its two
// phases; nothing is sized by host-side knowledge of the frontier (its length is read from
// x[0].index), so the launch can be recorded and replayed:
//   phase 1  scatter: a warp takes every (number of warps)-th frontier entry; columns longer than kSeg
//            non-zeros are queued in the CTA's shared memory and cut into segments its 8 warps share
//            (a 10^4-long hub column is spread over the CTA instead of serialising one warp).
//            a (x) v is combined into a dense accumulator that rests at the (+)-identity with
//            fire-and-forget operations (reduction atomics, an idempotent store for or-and: nothing
//            waits for a result).  A SMALL frontier also marks the rows it touches in a bitmap
//            (one more reduction atomic per non-zero, 1/32 of the accumulator's size);
//   phase 2  compact: a small frontier walks the bitmap and visits only the touched rows, a large one
//            scans the accumulator of the shard: fold `zero`, apply the mask, reset accumulator and
//            bitmap, emit with warp-aggregated appends -- and apply the
//            fused epilogue on the emitted entries: the sparse assign of a BFS push level
//            (inout[row] = val, bfs.h:147-151) or the relax + new-frontier of an SSSP push level
//            (assign_vector_sparse_module.h:318-335, sssp.h:178-190), so a push level is one launch.
// Direction switch on the device (bfs.h:160-219, sssp.h:197-243): the last CTA out of phase 2
// evaluates the reference's "keep pushing" test on the result count, writes it where the recorded
// IF / ELSE node of the next level reads it (cudaGraphSetConditional), and when pushing stops every
// CTA helps to turn the frontier into the dense input of the first pull level.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "glb_internal.h"
#include "semiring.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kSeg = 512;             // non-zeros per unit of work (one warp: 16 strides of 32)
constexpr float kFloatInf = 999999999.0f;  // FLOAT_INF, global.h:80 (spmspv_module.h:482-491)

// Order-preserving map float -> signed int (and back: it is an involution): negative floats, whose bit patterns
// order inversely, get their magnitude bits flipped.  -0.0f sorts just below +0.0f; NaN keys are never produced
// by min-plus on ordered inputs.
__device__ __forceinline__ int f32_key(float v) {
    const int b = __float_as_int(v);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float f32_unkey(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

// (x) of the scatter.  fp32 min-plus clamps at FLOAT_INF like the reference's CPU path
// (spmspv_module.h:482-491); the integer value types follow the processing element's ALU
// (ufixed_pe_fwd.h:23-44): plain (wrapping / saturating) addition.
template <int OP, int VT>
__device__ __forceinline__ float spmspv_mul(float a, float v) {
    if (OP == GLB_OP_ADD_MIN && VT == GLB_VAL_F32) {
        if (a > kFloatInf || v > kFloatInf) return kFloatInf;
        const float s = __fadd_rn(a, v);
        return s > kFloatInf ? kFloatInf : s;
    }
    return Semi<OP, VT>::mul(a, v);
}

// device-resident state of one CSC matrix: two sets of counters used alternately, so that a launch
// resets the set of the NEXT launch while nobody reads it (no host-side reset, replayable)
struct SpmspvCounters {
    uint32_t q_tail, q_head, producers_done, bar[3], done;
    uint32_t pad;
};
struct SpmspvState {
    uint32_t parity;
    uint32_t keep_pushing;  // decision of the last launch that carried a `next` block (1 = push again)
    uint32_t push_levels;   // launches with a `next` block since glb_spmspv_reset_levels (push_iterations_ of the apps)
    uint32_t pad;
    SpmspvCounters c[2];
};

struct SpmspvParams {
    const uint32_t *__restrict__ indptr;
    const uint32_t *__restrict__ indices;
    const float *__restrict__ vals;
    const glb_idx_val_t *x;
    const float *mask;
    glb_idx_val_t *y;
    float *acc;
    uint32_t *bitmap;   // bit r = row r was touched by this launch (small frontiers only); at rest all zero
    unsigned long long *queue;  // segments of long columns: {frontier slot, segment}
    uint32_t queue_cap;
    SpmspvState *state;
    uint32_t row_begin, row_end;
    uint32_t num_rows, num_cols;
    float zero;
    int mask_type;
    // fused epilogue on the emitted entries
    int ep_mode;        // GLB_SPMSPV_EP_NONE / _ASSIGN / _RELAX
    float *ep_inout;
    float ep_val;
    glb_idx_val_t *ep_new_frontier;
    // direction decision after this push level (glb_spmspv_next_t)
    int has_next;
    int force_stop;
    float threshold;
    float n_vertices;
    unsigned long long cond_next;  // cudaGraphConditionalHandle of the next level's IF / ELSE node (0: none)
    int dense_mode;                // on stop: GLB_SPMSPV_DENSE_NONE / _SCATTER (y list -> dense) / _COPY (dense_src -> dense)
    float *dense;
    const float *dense_src;
    uint32_t dense_len;
};

// all CTAs of the launch are co-resident (4 per SM): a counter barrier
__device__ __forceinline__ void grid_barrier(uint32_t *counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        uint32_t seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < gridDim.x);
    }
    __syncthreads();
}

// Non-zeros [s, t) of one column by a group of STRIDE lanes (`sub` = lane inside the group): all combines
// are fire-and-forget (reduction atomics / idempotent stores), so the index / value loads of a lane are in
// flight together and nothing waits for an accumulator.  MARK: also set the row's bit in the touched bitmap.
template <int OP, bool MARK, int STRIDE, int VT>
__device__ __forceinline__ void scatter_span(const SpmspvParams &P, uint32_t s, uint32_t t, float v, unsigned sub) {
#pragma unroll 8
    for (uint32_t i = s + sub; i < t; i += STRIDE) {
        const uint32_t row = __ldg(P.indices + i);
        const float p = spmspv_mul<OP, VT>(__ldg(P.vals + i), v);
        float *a = P.acc + row;
        if (OP == GLB_OP_MUL_ADD) {
            if (VT == GLB_VAL_F32) {
                atomicAdd(a, p);
            } else if (VT == GLB_VAL_U32) {
                atomicAdd(reinterpret_cast<unsigned *>(a), __float_as_uint(p));  // modulo 2^32
            } else {  // saturating sum: no such atomic, compare-and-swap (sums of non-negative words commute)
                unsigned *w = reinterpret_cast<unsigned *>(a);
                unsigned old = *w, seen;
                do {
                    seen = old;
                    old = atomicCAS(w, seen, __float_as_uint(Val<VT>::plus(__uint_as_float(seen), p)));
                } while (old != seen);
            }
        } else if (OP == GLB_OP_LOGICAL_AND_OR) {
            if (!Val<VT>::is_zero(p)) *a = Val<VT>::one();  // idempotent: racing writers store the same word
        } else if (VT == GLB_VAL_F32) {
            // float min as ONE integer reduction: the accumulator holds order-preserving keys (f32_key) of the
            // floats, so signed-int min is float min for every sign; +inf is its own key, so the rest state is +inf
            atomicMin(reinterpret_cast<int *>(a), f32_key(p));
        } else {
            atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(p));
        }
        if (MARK) atomicOr(P.bitmap + (row >> 5), 1u << (row & 31u));
    }
}

constexpr int kSub = 8;                 // lanes per light column: a warp works on 32 / kSub columns at a time
constexpr int kColsPerWarp = 32 / kSub;
  // heavy columns a CTA can queue for its warps to share

template <int OP, int VT>
__global__ void __launch_bounds__(kThreads, 4) spmspv_kernel(const SpmspvParams P) {
    __shared__ uint32_t s_last;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    SpmspvState *S = P.state;
    const uint32_t par = S->parity;  // flipped by CTA 0 after the first barrier: every CTA has read it by then
    SpmspvCounters *C = &S->c[par];
    const uint32_t nnz_x = P.x[0].index;
    // Small frontier: mark the touched rows in the bitmap, so that phase 2 visits only those instead of
    // scanning the shard's accumulator; a large frontier touches most rows anyway and saves the marks.
    const bool tracked = uint64_t(nnz_x) * 64u <= uint64_t(P.row_end - P.row_begin);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.y[0].index = 0;
        P.y[0].val = P.zero;
        if (P.ep_mode == GLB_SPMSPV_EP_RELAX) {
            P.ep_new_frontier[0].index = 0;
            P.ep_new_frontier[0].val = 0.0f;
        }
        SpmspvCounters *N = &S->c[par ^ 1u];  // the next launch's counters
        N->q_tail = 0; N->q_head = 0; N->producers_done = 0; N->bar[0] = 0; N->bar[1] = 0; N->bar[2] = 0; N->done = 0;
    }

    // ---- phase 1a: short columns, and the long ones cut into queued segments ------------------------
    // What bounds this phase on short columns is the dependent chain frontier entry -> column bounds ->
    // indices -> atomics, one chain per column: so a warp works on 4 columns at a time (8 lanes each).
    // Consecutive frontier entries go to different CTAs (a 6-vertex frontier must not land on one SM).
    // A column longer than kSeg non-zeros is cut into kSeg-long segments pushed to a device-wide queue
    // that all warps of the grid work off in phase 1b, after a grid barrier: a 10^4-long hub column is
    // spread over many SMs -- the reduction atomics of ONE SM retire at ~1 per clock, 27 000 of them
    // would take 20 us.  A launch without long columns skips phase 1b and its barrier.
    const unsigned sub = lane & (kSub - 1), grp = lane / kSub;
    const unsigned grp_mask = ((1u << kSub) - 1u) << (grp * kSub);  // control flow is uniform per lane group
    const uint32_t n_lane_groups = gridDim.x * kWarps * kColsPerWarp;
    for (uint32_t k = (warp * kColsPerWarp + grp) * gridDim.x + blockIdx.x; k < nnz_x; k += n_lane_groups) {
        const glb_idx_val_t e = P.x[k + 1];
        if (e.index >= P.num_cols) continue;  // not a column of this matrix: ignored
        const uint32_t s = __ldg(P.indptr + e.index), t = __ldg(P.indptr + e.index + 1);
        if (t - s > kSeg) {
            const uint32_t n_seg = (t - s + kSeg - 1) / kSeg;
            uint32_t base = 0;
            if (sub == 0) base = atomicAdd(&C->q_tail, n_seg);
            base = __shfl_sync(grp_mask, base, int(grp * kSub));
            const bool fits = base + n_seg <= P.queue_cap;  // (a frontier naming columns repeatedly can outgrow the queue)
            for (uint32_t j = sub; j < n_seg && base + j < P.queue_cap; j += kSub) {
                const unsigned long long item = fits ? ((unsigned long long)(k) << 32 | j) : ((unsigned long long)(k) << 32 | 0xfffffffeull);
                P.queue[base + j] = item;
            }
            if (fits) continue;
        }
        if (tracked) scatter_span<OP, true, kSub, VT>(P, s, t, e.val, sub);
        else scatter_span<OP, false, kSub, VT>(P, s, t, e.val, sub);
    }
    grid_barrier(&C->bar[0]);
    if (blockIdx.x == 0 && threadIdx.x == 0) S->parity = par ^ 1u;

    // ---- phase 1b: the queued segments, one warp each (the queue is complete: every CTA passed the barrier) ----
    uint32_t n_queued;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(n_queued) : "l"(&C->q_tail) : "memory");
    if (n_queued > P.queue_cap) n_queued = P.queue_cap;
    if (n_queued) {  // grid-uniform
        const uint32_t n_warps = gridDim.x * kWarps;
        for (uint32_t q = warp * gridDim.x + blockIdx.x; q < n_queued; q += n_warps) {  // consecutive segments on different SMs
            const unsigned long long item = __ldcg(P.queue + q);
            const uint32_t seg = uint32_t(item);
            if (seg == 0xfffffffeu) continue;  // voided reservation (queue overflow): its producer walked the column itself
            const glb_idx_val_t e = P.x[uint32_t(item >> 32) + 1];
            const uint32_t t = __ldg(P.indptr + e.index + 1);
            const uint32_t b = __ldg(P.indptr + e.index) + seg * kSeg;
            const uint32_t b_end = t - b > kSeg ? b + kSeg : t;
            if (tracked) scatter_span<OP, true, 32, VT>(P, b, b_end, e.val, lane);
            else scatter_span<OP, false, 32, VT>(P, b, b_end, e.val, lane);
        }
        grid_barrier(&C->bar[2]);
    }

    // ---- phase 2: compact -----------------------------------------------------------------------------
    // A warp step covers kBatch groups of 32 consecutive rows (lane = row inside the group: coalesced).
    // Tracked: the step first reads the kBatch bitmap words and is skipped when all are zero.  All loads of
    // a step (accumulator, mask, relax target) are issued together, the emitted entries are counted with
    // shuffles and the output cursors are advanced once per step.
    constexpr int kBatch = 8;
    const uint32_t base_row = P.row_begin & ~31u;
    const uint32_t n_groups = (P.row_end - base_row + 31u) >> 5;
    const uint32_t n_warps_total = gridDim.x * kWarps, gwarp = blockIdx.x * kWarps + warp;
    const bool relax = P.ep_mode == GLB_SPMSPV_EP_RELAX;
    for (uint32_t g0 = gwarp * kBatch; g0 < n_groups; g0 += n_warps_total * kBatch) {
        uint32_t my_word = 0xffffffffu;
        if (tracked) {
            my_word = (lane < kBatch && g0 + lane < n_groups) ? __ldcg(P.bitmap + (base_row >> 5) + g0 + lane) : 0u;
            if (my_word) P.bitmap[(base_row >> 5) + g0 + lane] = 0u;  // back to rest
            if (!__any_sync(kFull, my_word != 0u)) continue;
        }
        float a[kBatch], mk[kBatch], dv[kBatch];
        uint32_t live = 0;  // bit j: row j of this lane is a candidate
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const uint32_t word = __shfl_sync(kFull, my_word, j);
            const uint32_t r = base_row + ((g0 + j) << 5) + lane;
            const bool ok = g0 + j < n_groups && r >= P.row_begin && r < P.row_end && ((word >> lane) & 1u);
            a[j] = ok ? __ldcg(P.acc + r) : Semi<OP, VT>::ident();
            if (OP == GLB_OP_ADD_MIN && VT == GLB_VAL_F32) a[j] = f32_unkey(__float_as_int(a[j]));  // keys -> floats (+inf is its own key)
            mk[j] = (ok && P.mask_type != GLB_MASK_NONE) ? P.mask[r] : 0.0f;
            dv[j] = (ok && relax) ? P.ep_inout[r] : 0.0f;
            if (ok) live |= 1u << j;
        }
        uint32_t emit = 0, improved = 0;
        float val[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            val[j] = 0.0f;
            if (!((live >> j) & 1u) || Val<VT>::equal(a[j], Semi<OP, VT>::ident())) continue;
            const uint32_t r = base_row + ((g0 + j) << 5) + lane;
            P.acc[r] = Semi<OP, VT>::ident();
            val[j] = Semi<OP, VT>::with_zero(P.zero, a[j]);
            bool off = false;
            if (P.mask_type == GLB_MASK_WRITE_TO_ONE) off = Val<VT>::equal(mk[j], P.zero);
            else if (P.mask_type == GLB_MASK_WRITE_TO_ZERO) off = !Val<VT>::equal(mk[j], P.zero);
            if (off || Val<VT>::equal(val[j], P.zero)) continue;
            emit |= 1u << j;
            if (P.ep_mode == GLB_SPMSPV_EP_ASSIGN) P.ep_inout[r] = P.ep_val;
            if (relax && Val<VT>::less(val[j], dv[j])) {
                P.ep_inout[r] = val[j];  // rows are unique: no other thread touches this element
                improved |= 1u << j;
            }
        }
        // positions: exclusive prefix of the per-lane counts, one cursor update per list and step
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const uint32_t bits = which ? improved : emit;
            if (which && !relax) break;  // uniform
            const uint32_t cnt = uint32_t(__popc(bits));
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t tq = __shfl_up_sync(kFull, incl, d);
                if (int(lane) >= d) incl += tq;
            }
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            if (total == 0) continue;  // uniform
            glb_idx_val_t *list = which ? P.ep_new_frontier : P.y;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&list[0].index, total);
            base = __shfl_sync(kFull, base, 0) + (incl - cnt);
#pragma unroll
            for (int j = 0; j < kBatch; ++j)
                if ((bits >> j) & 1u) {
                    glb_idx_val_t o;
                    o.index = base_row + ((g0 + j) << 5) + lane;
                    o.val = val[j];
                    list[1 + base++] = o;
                }
        }
    }
    if (!P.has_next) return;

    // ---- direction decision: the last CTA out of phase 2 sees the final count ------------------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&C->done, 1u) == gridDim.x - 1;
        if (s_last) {
            __threadfence();
            uint32_t count;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(count) : "l"(&P.y[0].index) : "memory");
            // bfs.h:190 / sssp.h:214: keep pushing while iter < num_iterations (force_stop, from the host) and
            // float(nnz) / num_vertices < threshold
            const uint32_t keep = (!P.force_stop && (float(count) / P.n_vertices < P.threshold)) ? 1u : 0u;
            S->keep_pushing = keep;
            S->push_levels += 1;
            if (P.cond_next) cudaGraphSetConditional(cudaGraphConditionalHandle(P.cond_next), keep);
        }
    }
    if (P.dense_mode == GLB_SPMSPV_DENSE_NONE) return;
    grid_barrier(&C->bar[1]);  // every CTA learns the decision
    uint32_t keep;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(keep) : "l"(&S->keep_pushing) : "memory");
    if (keep) return;
    // pushing stops here: build the dense input of the first pull level, all CTAs together
    const uint32_t n_threads = gridDim.x * kThreads;
    if (P.dense_mode == GLB_SPMSPV_DENSE_SCATTER) {  // `dense` was filled with the semiring zero when the run was set up
        const uint32_t count = P.y[0].index;
        for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < count; i += n_threads) {
            const glb_idx_val_t e = P.y[i + 1];
            if (e.index < P.dense_len) P.dense[e.index] = e.val;
        }
    } else {
        for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < P.dense_len; i += n_threads) P.dense[i] = P.dense_src[i];
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

__global__ void sparse_one_kernel(glb_idx_val_t *list, uint32_t index, float val) {
    list[0].index = 1;
    list[0].val = 0.0f;
    list[1].index = index;
    list[1].val = val;
}

__global__ void sparse_head_kernel(glb_idx_val_t *list, float zero) {
    list[0].index = 0;
    list[0].val = zero;
}

template <int OP, int VT>
int run_spmspv(glb_ctx_t ctx, glb_csc_t m, const SpmspvParams &P) {
    // The grid barriers need every CTA of the launch resident at once: ask the runtime how many CTAs of THIS
    // instantiation fit on an SM (once per instantiation) instead of trusting the launch bounds.
    static int per_sm = 0;
    if (per_sm == 0) {
        int n = 0;
        GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spmspv_kernel<OP, VT>, kThreads, 0));
        if (n < 1) {
            glb_set_error("glb_spmspv: the kernel does not fit on an SM");
            return GLB_ECUDA;
        }
        per_sm = n < 4 ? n : 4;
    }
    spmspv_kernel<OP, VT><<<ctx->num_sms * per_sm, kThreads, 0, ctx->stream>>>(P);
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
    alloc(reinterpret_cast<void **>(&m->bitmap), sizeof(uint32_t) * ((size_t(num_rows) + 31) / 32 + 1));
    alloc(reinterpret_cast<void **>(&m->state), sizeof(SpmspvState));
    m->queue_cap = uint32_t(std::min<uint64_t>(nnz / kSeg + num_cols + 64, 0x7fffffffull));
    alloc(reinterpret_cast<void **>(&m->queue), sizeof(unsigned long long) * m->queue_cap);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(m->indptr, indptr, sizeof(uint32_t) * (size_t(num_cols) + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz)
        e = cudaMemcpyAsync(m->indices, indices, sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz)
        e = cudaMemcpyAsync(m->vals, data, sizeof(float) * nnz, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->bitmap, 0, sizeof(uint32_t) * ((size_t(num_rows) + 31) / 32 + 1), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->state, 0, sizeof(SpmspvState), ctx->stream);
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
    cudaFree(m->indptr); cudaFree(m->indices); cudaFree(m->vals); cudaFree(m->acc); cudaFree(m->acc_inf); cudaFree(m->acc_max);
    cudaFree(m->bitmap); cudaFree(m->state); cudaFree(m->queue);
    glb_ctx_release(m->ctx);
    delete m;
    return GLB_OK;
}

static int spmspv_any(glb_ctx_t ctx, glb_csc_t m, int val_type, int op, float zero, int mask_type, const glb_idx_val_t *x,
                      const float *mask, glb_idx_val_t *y, const glb_spmspv_epilogue_t *ep, const glb_spmspv_next_t *next) {
    GLB_REQUIRE(ctx && m && x && y, "NULL argument");
    GLB_REQUIRE(val_type >= GLB_VAL_F32 && val_type <= GLB_VAL_UFIXED, "invalid value type");
    GLB_REQUIRE(m->ctx == ctx, "matrix belongs to another context");
    GLB_REQUIRE(mask_type >= GLB_MASK_NONE && mask_type <= GLB_MASK_WRITE_TO_ONE, "invalid mask type");
    GLB_REQUIRE(mask_type == GLB_MASK_NONE || mask, "mask is NULL but mask_type != kNoMask");
    GLB_REQUIRE(static_cast<const void *>(x) != static_cast<const void *>(y), "y must not alias x");
    SpmspvParams P;
    memset(&P, 0, sizeof(P));
    P.indptr = m->indptr;
    P.indices = m->indices;
    P.vals = m->vals;
    P.x = x;
    P.mask = mask;
    P.y = y;
    // one accumulator per (+)-identity, both filled when the matrix was created and left at rest by
    // every run: no launch depends on host-side state (launch sequences can be recorded at any time)
    P.acc = (op == GLB_OP_ADD_MIN) ? m->acc_inf : m->acc;
    if (op == GLB_OP_ADD_MIN && val_type != GLB_VAL_F32) {  // the integer infinity is all ones: a third accumulator, on first use
        if (!m->acc_max) {
            GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->acc_max), sizeof(float) * (m->num_rows ? m->num_rows : 1)));
            GLB_CUDA(cudaMemsetAsync(m->acc_max, 0xff, sizeof(float) * m->num_rows, ctx->stream));
        }
        P.acc = m->acc_max;
    }
    P.bitmap = m->bitmap;
    P.queue = m->queue;
    P.queue_cap = m->queue_cap;
    P.state = static_cast<SpmspvState *>(m->state);
    P.row_begin = m->row_begin;
    P.row_end = m->row_end;
    P.num_rows = m->num_rows;
    P.num_cols = m->num_cols;
    P.zero = zero;
    P.mask_type = mask_type;
    if (ep && ep->mode != GLB_SPMSPV_EP_NONE) {
        GLB_REQUIRE(ep->mode == GLB_SPMSPV_EP_ASSIGN || ep->mode == GLB_SPMSPV_EP_RELAX, "invalid epilogue mode");
        GLB_REQUIRE(ep->inout, "epilogue needs an inout vector");
        GLB_REQUIRE(ep->mode != GLB_SPMSPV_EP_RELAX || ep->new_frontier, "relax epilogue needs a new-frontier list");
        GLB_REQUIRE(static_cast<const void *>(ep->new_frontier) != static_cast<const void *>(x) &&
                        static_cast<const void *>(ep->new_frontier) != static_cast<const void *>(y),
                    "new_frontier must not alias x or y");
        P.ep_mode = ep->mode;
        P.ep_inout = ep->inout;
        P.ep_val = ep->val;
        P.ep_new_frontier = ep->new_frontier;
    }
    if (next) {
        GLB_REQUIRE(next->num_vertices > 0, "num_vertices must be positive");
        GLB_REQUIRE(next->dense_mode >= GLB_SPMSPV_DENSE_NONE && next->dense_mode <= GLB_SPMSPV_DENSE_COPY, "invalid dense mode");
        GLB_REQUIRE(next->dense_mode == GLB_SPMSPV_DENSE_NONE || next->dense, "dense vector is NULL");
        GLB_REQUIRE(next->dense_mode != GLB_SPMSPV_DENSE_COPY || next->dense_src, "dense_src is NULL");
        P.has_next = 1;
        P.force_stop = next->force_stop;
        P.threshold = next->threshold;
        P.n_vertices = float(next->num_vertices);
        P.cond_next = next->cond_next;
        P.dense_mode = next->dense_mode;
        P.dense = next->dense;
        P.dense_src = next->dense_src;
        P.dense_len = next->dense_len;
    }
    switch (op * 3 + val_type) {
        case GLB_OP_MUL_ADD * 3 + GLB_VAL_F32: return run_spmspv<GLB_OP_MUL_ADD, GLB_VAL_F32>(ctx, m, P);
        case GLB_OP_LOGICAL_AND_OR * 3 + GLB_VAL_F32: return run_spmspv<GLB_OP_LOGICAL_AND_OR, GLB_VAL_F32>(ctx, m, P);
        case GLB_OP_ADD_MIN * 3 + GLB_VAL_F32: return run_spmspv<GLB_OP_ADD_MIN, GLB_VAL_F32>(ctx, m, P);
        case GLB_OP_MUL_ADD * 3 + GLB_VAL_U32: return run_spmspv<GLB_OP_MUL_ADD, GLB_VAL_U32>(ctx, m, P);
        case GLB_OP_LOGICAL_AND_OR * 3 + GLB_VAL_U32: return run_spmspv<GLB_OP_LOGICAL_AND_OR, GLB_VAL_U32>(ctx, m, P);
        case GLB_OP_ADD_MIN * 3 + GLB_VAL_U32: return run_spmspv<GLB_OP_ADD_MIN, GLB_VAL_U32>(ctx, m, P);
        case GLB_OP_MUL_ADD * 3 + GLB_VAL_UFIXED: return run_spmspv<GLB_OP_MUL_ADD, GLB_VAL_UFIXED>(ctx, m, P);
        case GLB_OP_LOGICAL_AND_OR * 3 + GLB_VAL_UFIXED: return run_spmspv<GLB_OP_LOGICAL_AND_OR, GLB_VAL_UFIXED>(ctx, m, P);
        case GLB_OP_ADD_MIN * 3 + GLB_VAL_UFIXED: return run_spmspv<GLB_OP_ADD_MIN, GLB_VAL_UFIXED>(ctx, m, P);
    }
    glb_set_error("glb_spmspv: invalid semiring op %d", op);
    return GLB_EINVAL;
}

int glb_spmspv_fused(glb_ctx_t ctx, glb_csc_t m, int op, float zero, int mask_type, const glb_idx_val_t *x, const float *mask,
                     glb_idx_val_t *y, const glb_spmspv_epilogue_t *ep, const glb_spmspv_next_t *next) {
    return spmspv_any(ctx, m, GLB_VAL_F32, op, zero, mask_type, x, mask, y, ep, next);
}

int glb_spmspv_vt(glb_ctx_t ctx, glb_csc_t m, int val_type, int op, uint32_t zero_bits, int mask_type, const glb_idx_val_t *x,
                  const void *mask, glb_idx_val_t *y) {
    float zero;
    memcpy(&zero, &zero_bits, sizeof(zero));
    return spmspv_any(ctx, m, val_type, op, zero, mask_type, x, static_cast<const float *>(mask), y, nullptr, nullptr);
}

int glb_spmspv(glb_ctx_t ctx, glb_csc_t m, int op, float zero, int mask_type, const glb_idx_val_t *x, const float *mask,
               glb_idx_val_t *y) {
    return glb_spmspv_fused(ctx, m, op, zero, mask_type, x, mask, y, nullptr, nullptr);
}

int glb_spmspv_push_state(glb_ctx_t ctx, glb_csc_t m, uint32_t *keep_pushing, uint32_t *push_levels) {
    GLB_REQUIRE(ctx && m && m->ctx == ctx, "bad argument");
    SpmspvState h;
    GLB_CUDA(cudaMemcpyAsync(&h, m->state, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (keep_pushing) *keep_pushing = h.keep_pushing;
    if (push_levels) *push_levels = h.push_levels;
    return GLB_OK;
}

int glb_spmspv_reset_levels(glb_ctx_t ctx, glb_csc_t m) {
    GLB_REQUIRE(ctx && m && m->ctx == ctx, "bad argument");
    // keep_pushing, push_levels (parity and the counter sets are left alone)
    GLB_CUDA(cudaMemsetAsync(reinterpret_cast<char *>(m->state) + sizeof(uint32_t), 0, 2 * sizeof(uint32_t), ctx->stream));
    return GLB_OK;
}

int glb_sparse_fill_one(glb_ctx_t ctx, glb_idx_val_t *list, uint32_t index, float val) {
    GLB_REQUIRE(ctx && list, "NULL argument");
    sparse_one_kernel<<<1, 1, 0, ctx->stream>>>(list, index, val);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
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
