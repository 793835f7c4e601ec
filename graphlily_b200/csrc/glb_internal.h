// Internal declarations shared by the translation units of libgraphlily_b200.so.
#ifndef GLB_INTERNAL_H_
#define GLB_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "graphlily_b200.h"

// ---------------------------------------------------------------- error plumbing
void glb_set_error(const char *fmt, ...);

#define GLB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            glb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return GLB_ECUDA;                                                                   \
        }                                                                                       \
    } while (0)

#define GLB_REQUIRE(cond, msg)                                      \
    do {                                                            \
        if (!(cond)) {                                              \
            glb_set_error("%s: %s", __func__, msg);                 \
            return GLB_EINVAL;                                      \
        }                                                           \
    } while (0)

// ---------------------------------------------------------------- context
struct glb_ctx_s {
    // Matrices, exchanges and recorded sequences keep their context alive: glb_ctx_destroy with
    // children outstanding only marks it, the last child's destroy releases it (hosts with garbage
    // collection destroy objects in any order).
    int children = 0;
    bool destroyed = false;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    void *nccl_comm = nullptr;  // ncclComm_t when built with GLB_WITH_NCCL
    int nccl_rank = 0, nccl_nranks = 1;
    // optional per-kernel timing (glb_ctx_kernel_timing): event triples of timed launches
    bool timing = false;
    std::vector<cudaEvent_t> timing_events;  // 3 per launch: before main, after main, after fix-up
    // page-locked bounce buffer for blocking copies from / to pageable host memory (two slots)
    void *staging = nullptr;
    size_t staging_bytes = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    // kernel attributes already applied on this device (cudaFuncSetAttribute is per device)
    int carveout_set[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};  // [value type][semiring]
    int bits_carveout_set = -1;
    size_t tile_smem_set[3] = {0, 0, 0};
    // recording the two arms of a branch (glb_graph_branch_*): launches go to a side stream that
    // captures into the body graph of the IF / ELSE node
    cudaStream_t branch_stream = nullptr, outer_stream = nullptr;
    cudaGraph_t branch_body[2] = {nullptr, nullptr};
    int in_branch = 0;  // 1: IF body, 2: ELSE body
    // split steps of a row-sharded run (spmv.cu: launch_split): one stream per sub-block of the shard
#define GLB_MAX_SPLIT 8
    cudaStream_t split_stream[GLB_MAX_SPLIT] = {};
    cudaEvent_t split_ev_head = nullptr, split_ev_main[GLB_MAX_SPLIT] = {}, split_ev_done[GLB_MAX_SPLIT] = {};
    // pusher CTAs of a row-sharded step (GLB_XCHG_MC=pusher): a kernel on a stream of its own beside the SpMV kernels
    cudaStream_t pusher_stream = nullptr;
    cudaEvent_t pusher_ev_fork = nullptr, pusher_ev_done = nullptr;
    int pusher_smem_set[3] = {-1, -1, -1};
    // copy streams + events of the pipelined host-buffer path (glb_spmv_host_batch), created on first use
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t pipe_ev[3][2] = {};  // [uploaded | computed | downloaded][slot]
    cudaEvent_t pipe_pushed = nullptr;  // host batch over an exchange: "the slice push of the current vector has left"
};

// ---------------------------------------------------------------- lane-segment CSR (SpMV)
//
// The nnz stream of the row shard is cut into chunks of up to GLB_MAX_GROUPS groups of GLB_GROUP
// non-zeros; one warp owns one chunk.  In a chunk of n groups lane l owns the 4n consecutive
// non-zeros [4n*l, 4n*(l+1)) (positions past the chunk's real length are padding); storage is
// transposed so that group g holds elements 4g..4g+3 of every lane:
//   stream[256*G + 4*l + e]        encoded column of element 4g+e of lane l   (G = chunk_goff[c] + g)
//   stream[256*G + 128 + 4*l + e]  its fp32 value
//   encoded column                 rank (< tile_k) of a hot column in hot_cols, else tile_k + column
//   flags[32*c + l]                bit r = "element r of lane l starts a row" (never set for the chunk's
//                                  first non-zero; the first padding element, if any, is flagged)
//   chunk_goff[c]                  first group of chunk c (n_chunks + 1 entries)
//   chunk_first[c]                 ordinal k into nz_rows of the row open at the chunk start,
//                                  bit31 = that row starts exactly there
//   nz_rows[k]                     row id of the k-th non-empty row of the shard
// A chunk holds at most GLB_ROW_CAP flags.  Rows that cross a chunk boundary (or end exactly on
// one) are finished by the fix-up kernel from per-chunk carries: see spmv.cu.
#define GLB_GROUP 128u
#define GLB_MAX_GROUPS 8
#ifndef GLB_ROW_CAP
#define GLB_ROW_CAP 128u
#endif
#define GLB_FLAG 0x80000000u
#define GLB_DEFAULT_TILE_K 40960u
#define GLB_DEFAULT_CARVEOUT_PCT 12u
#define GLB_DEFAULT_TILE_THREADS 0u

struct glb_fixup_t {
    uint32_t row;    // global row id
    uint32_t c_begin;  // first chunk contributing its tail carry
    uint32_t c_end;    // last chunk; bit31 = it contributes its HEAD carry (row ends mid-chunk)
};

struct glb_csr_s {
    glb_ctx_t ctx = nullptr;
    uint32_t num_rows = 0, num_cols = 0;  // global dims
    uint32_t row_begin = 0, row_end = 0;  // shard
    uint64_t nnz = 0;                     // in shard
    uint32_t n_chunks = 0, n_groups = 0;
    uint32_t n_nz_rows = 0;
    uint32_t n_fix_short = 0, n_fix_long = 0, n_empty = 0;
    // device arrays
    uint32_t *stream = nullptr;
    uint32_t *flags = nullptr;
    uint32_t *chunk_goff = nullptr;
    uint32_t *chunk_first = nullptr;
    uint32_t *nz_rows = nullptr;
    glb_fixup_t *fix_short = nullptr;  // span <= 32 chunks: one thread each
    glb_fixup_t *fix_long = nullptr;   // longer spans: one warp each
    uint32_t *empty_rows = nullptr;
    float *head_carry = nullptr, *tail_carry = nullptr;  // per chunk
    // hot columns: the tile_k most referenced columns of the shard, renumbered 0 .. tile_k-1
    uint32_t tile_k = 0;           // encoded columns below it are hot
    uint32_t n_hot = 0;            // entries of hot_cols (0 when tile_k == 0 or tile_k == num_cols: identity)
    uint32_t *hot_cols = nullptr;  // rank -> column
    float *hot_x = nullptr;        // x[hot_cols[.]], rebuilt by every SpMV launch
    uint32_t *xbits = nullptr;     // or-and: one bit per stored column word, rebuilt by every or-and launch
    uint32_t *mbits = nullptr;     // masked or-and: one bit per row of the shard (mask != 0), rebuilt with xbits
    uint32_t uniform_groups = 0, last_groups = 0;  // all chunks but the last hold uniform_groups groups (0: they vary)
    bool all_nonzero = false;      // no stored value is 0.0f (or-and skips the value stream)
    int smem_carveout_pct = 20;
    uint32_t tile_threads = 0;     // > 0: persistent shared-memory-tile kernel with that many threads per CTA
    // sub-blocks of a split step (host side): chunk / row boundaries and the ranges of the fix-up and empty-row
    // lists whose rows lie in each sub-block
    std::vector<uint32_t> sub_chunk, sub_row, sub_fs, sub_fl, sub_em;
    // progressive push of a row-sharded run over a multicast exchange (spmv.cu: push_block_when_complete)
    uint32_t *push_bits = nullptr, *push_lo = nullptr, *push_count = nullptr;  // push_count[n_push_blocks] counts fix-up CTAs
    uint32_t n_push_blocks = 0;
    // pusher CTAs (spmv.cu: xchg_pusher_kernel): fix-up rows inside one push block / across blocks, per-block ranges
    glb_fixup_t *fix_in = nullptr, *fix_def = nullptr;
    uint32_t *blk_fs = nullptr, *blk_em = nullptr, *pusher_bits = nullptr;
    uint32_t n_fix_def = 0;
    // scratch vectors for glb_spmv_host
    float *dx = nullptr, *dmask = nullptr, *dy = nullptr;
    float *dx2 = nullptr, *dmask2 = nullptr, *dy2 = nullptr;  // second slot of glb_spmv_host_batch
    size_t device_bytes = 0;
};

// ---------------------------------------------------------------- CSC (SpMSpV)
struct glb_csc_s {
    glb_ctx_t ctx = nullptr;
    uint32_t num_rows = 0, num_cols = 0;
    uint32_t row_begin = 0, row_end = 0;  // output rows this shard owns (all rows when not sharded)
    uint64_t nnz = 0;                     // non-zeros kept (rows inside the shard)
    uint32_t *indptr = nullptr;   // num_cols + 1
    uint32_t *indices = nullptr;  // row ids
    float *vals = nullptr;
    float *acc = nullptr;         // dense accumulator of plus-times / or-and, num_rows, at rest 0.0f between runs
    float *acc_inf = nullptr;     // dense accumulator of min-plus, at rest +inf
    float *acc_max = nullptr;     // ... of min-plus on the integer value types, at rest 0xffffffff (allocated on first use)
    uint32_t *bitmap = nullptr;   // rows touched by the running launch (small frontiers), all zero at rest
    unsigned long long *queue = nullptr;  // segment queue of the long columns (spmspv.cu), all ones at rest
    uint32_t queue_cap = 0;
    void *state = nullptr;        // SpmspvState (spmspv.cu): counters, direction decision, push levels
};

// ---------------------------------------------------------------- peer exchange (multi-GPU)
// Every rank owns one cudaMalloc'd block [n_vectors x n floats][flags]; the blocks of all ranks
// are mapped into every process with CUDA IPC, so a kernel can store rows straight into the
// peers' copies.  flags[s] of rank r = last epoch rank s has finished writing into r's block.
#define GLB_MAX_PEERS 7
struct glb_xchg_s {
    glb_ctx_t ctx = nullptr;
    uint32_t n = 0;      // floats per vector
    int n_vectors = 0;
    float *local = nullptr;               // this rank's block
    uint32_t *local_flags = nullptr;      // GLB_MAX_PEERS + 1 slots, inside the block
    bool connected = false;
    int rank = 0, nranks = 1;
    float *peer[GLB_MAX_PEERS + 1] = {};        // blocks of all ranks (peer[rank] == local)
    uint32_t *peer_flags[GLB_MAX_PEERS + 1] = {};
    uint32_t **d_peer_flags = nullptr;    // device copy of peer_flags
    // device-resident state: [0] epoch this rank published last (advanced by the publishing kernel, so
    // recorded launch sequences replay), [1] CTA ticket of the push kernel, [2] set when a wait timed out
    uint32_t *d_state = nullptr;
    // adopted blocks (glb_xchg_adopt): memory mapped by the host (symmetric memory), optionally with
    // a multicast mapping -- one store to `mc` lands in every rank's block (NVSwitch multicast)
    bool adopted = false;
    float *mc = nullptr;
    uint32_t *mc_flags = nullptr;  // multicast mapping of the flag words
    // multicast object created through the C ABI (glb_xchg_mc_open / _bind, exchange.cu): driver handles to release
    bool mc_native = false;
    unsigned long long mc_handle = 0, mem_handle = 0;  // CUmemGenericAllocationHandle
    size_t mc_size = 0;
    bool mc_bound = false;
};
extern "C" int glb_xchg_signal_wait(glb_ctx_t ctx, glb_xchg_t xc);  // internal (not in the public header)
struct GlbXchgWait;                                        // exchange.cuh
GlbXchgWait glb_xchg_wait_desc(glb_xchg_t xc);
int glb_xchg_signal(glb_ctx_t ctx, glb_xchg_t xc, bool wait);
int glb_xchg_wait(glb_ctx_t ctx, glb_xchg_t xc);
int glb_xchg_wait_launch(glb_ctx_t ctx, const GlbXchgWait &w);
int glb_xchg_push(glb_ctx_t ctx, glb_xchg_t xc, int which, size_t offset, size_t count);

struct glb_graph_s {
    cudaGraphExec_t exec = nullptr;
    glb_ctx_t ctx = nullptr;
    int device = 0;
};

void glb_ctx_retain(glb_ctx_t ctx);
void glb_ctx_release(glb_ctx_t ctx);  // child destroyed; frees the context if it was destroyed meanwhile

// launchers (defined in the .cu files)
// Multicast destination of a row-sharded launch: y_mc = multicast mapping of y.  progressive: finished
// blocks of rows are pushed from inside the main kernel, the fix-up kernel sends its rows and publishes
// the epoch (*published tells whether it did -- it is not launched when it has no rows); otherwise every
// row is stored to y_mc by the write-back that produces it and the caller publishes.
struct GlbSpmvMc {
    float *y_mc;
    bool progressive;
    bool pusher;  // the kernels only COUNT finished CTAs per push block / fix-up CTAs; glb_launch_pusher's CTAs send the rows
    uint32_t *pub_flags_mc, *pub_state;
    int rank;
};
// Split step: the finished rows of every sub-block go to these peer copies of y with copy-engine copies while
// later sub-blocks still compute (launch_split); the caller publishes the epoch afterwards.
struct GlbSpmvSplit {
    float *y_peers[GLB_MAX_PEERS];
    int n_peers;
    bool last_by_caller;  // the caller sends the last sub-block's rows itself (multicast push kernel)
};
// `wait`: acquire of the previous step's exchange, folded into the head of the launch's first kernel (or NULL)
int glb_launch_spmv(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x, const float *mask,
                    float *y, const glb_spmv_epilogue_t *ep, float *const *y_peers, int n_peers, const GlbSpmvMc *mc,
                    const GlbXchgWait *wait, bool *published, int val_type, const GlbSpmvSplit *split);

// The pusher kernel of one row-sharded step (spmv.cu): launched on ctx->pusher_stream BEFORE the step's SpMV kernels.
int glb_xchg_preload();  // loads the exchange kernels (see glb_launch_pusher)
bool glb_pusher_applies(glb_csr_t m, const float *y, const float *y_mc);
int glb_launch_pusher(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *mask, float *y,
                      const glb_spmv_epilogue_t *ep, float *y_mc, uint32_t *mc_flags, uint32_t *state, int rank);

#endif  // GLB_INTERNAL_H_
