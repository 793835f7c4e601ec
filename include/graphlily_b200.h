/* graphlily_b200.h -- the C ABI of libgraphlily_b200.so (hand-written sm_100a CUDA kernels).
 *
 * This is the drop-in boundary for GraphLily's hot path.  In the reference the seam is
 * "module class -> OpenCL kernel `overlay`":
 *     extern "C" void overlay(..., mode)        /root/reference/graphlily/hw/overlay.cpp:14-81,308-414
 * whose `mode` selects one of six device functions, driven by kernel_.setArg(i, ...) +
 * enqueueTask + finish (spmv_module.h:104-128,471-475).  Here every overlay mode is one
 * entry point taking plain device pointers and sizes (no torch / C++ types):
 *
 *     overlay mode 1  kernel_spmv                          -> glb_spmv / glb_spmv_fused
 *     overlay mode 2  kernel_spmspv                        -> glb_spmspv
 *     overlay mode 3  kernel_add_scalar_vector_dense       -> glb_ewise_add
 *     overlay mode 4  kernel_assign_vector_dense           -> glb_assign_dense
 *     overlay mode 5  kernel_assign_vector_sparse_no_new_frontier -> glb_assign_sparse
 *     overlay mode 6  kernel_assign_vector_sparse_new_frontier    -> glb_assign_sparse_relax
 *
 * plus what the OpenCL runtime did around it: context / queue (BaseModule::set_up_runtime,
 * base_module.h:106-133), buffers (cl::Buffer + enqueueMigrateMemObjects), device-to-device
 * copies (base_module.h:82-85) and the matrix upload (send_matrix_host_to_device,
 * spmv_module.h:374-420, spmspv_module.h:290-370) -- which is where the device layout is built
 * (CPSR / formatCSC in the reference, the lane-segment layout here).
 *
 * Conventions
 *  - every function returns 0 on success, a GLB_E* code otherwise; glb_last_error() gives the
 *    text (thread-local).  There is NO CPU fallback: without a CUDA device every compute
 *    entry point fails with GLB_ECUDA.
 *  - all kernels are enqueued on the context's stream and return without synchronising;
 *    glb_ctx_sync, glb_buffer_d2h / glb_buffer_h2d and glb_sparse_count synchronise, which are
 *    the points where the reference calls command_queue_.finish() and the host reads data.
 *  - values are fp32 (`val_t = float`, global.h:64 variant), indices uint32 (idx_t, global.h:65).
 *  - sparse vectors follow global.h:69-70,153-164: element 0 is {index = nnz, val = unused},
 *    elements 1..nnz are {index, val}; order unspecified.
 */
#ifndef GRAPHLILY_B200_H_
#define GRAPHLILY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLB_VERSION 100

/* status codes */
#define GLB_OK 0
#define GLB_EINVAL 1  /* bad argument (the reference prints and exit()s, e.g. assign_vector_dense_module.h:88-95) */
#define GLB_ECUDA 2   /* CUDA runtime / launch failure, or no device */
#define GLB_ENOMEM 3  /* host or device allocation failure */
#define GLB_ENCCL 4   /* NCCL failure or library built without NCCL */

/* OperationType, global.h:83-87 */
#define GLB_OP_MUL_ADD 0
#define GLB_OP_LOGICAL_AND_OR 1
#define GLB_OP_ADD_MIN 2

/* MaskType, global.h:103-107 */
/* Value types (global.h:60-64 of the reference offers three; all are 32-bit words).  Every entry point
 * without a _vt suffix computes in GLB_VAL_F32, the type of the reference's CPU path and the parity target. */
#define GLB_VAL_F32 0     /* float */
#define GLB_VAL_U32 1     /* unsigned, arithmetic modulo 2^32, UINT_INF = 0xffffffff */
#define GLB_VAL_UFIXED 2  /* ap_ufixed<32, 8, AP_RND, AP_SAT>: Q8.24, rounded products, saturating sums */

#define GLB_MASK_NONE 0
#define GLB_MASK_WRITE_TO_ZERO 1 /* write where mask is zero     */
#define GLB_MASK_WRITE_TO_ONE 2  /* write where mask is non-zero */

/* idx_val_t with val_t = float, global.h:69 */
typedef struct glb_idx_val {
    uint32_t index;
    float val;
} glb_idx_val_t;

typedef struct glb_ctx_s *glb_ctx_t; /* one CUDA device + one stream (cl::Context + cl::CommandQueue)  */
typedef struct glb_csr_s *glb_csr_t; /* device-resident CSR row shard in lane-segment layout (SpMV)    */
typedef struct glb_csc_s *glb_csc_t; /* device-resident CSC (SpMSpV)                                    */

/* ------------------------------------------------------------------ runtime ------- */
int glb_version(void);
const char *glb_last_error(void);
int glb_device_count(int *count);

/* BaseModule::set_up_runtime / ModuleCollection::set_up_runtime (base_module.h:106-133,
 * module_collection.h:69-114).  `cuda_stream` may be an existing cudaStream_t (e.g. the
 * caller's torch stream) or NULL to create a private non-blocking stream. */
int glb_ctx_create(int device, void *cuda_stream, glb_ctx_t *out);
int glb_ctx_destroy(glb_ctx_t ctx);
int glb_ctx_sync(glb_ctx_t ctx);
/* Waits for ALL work on the context's device (every stream): what command_queue.finish() means to a
 * caller that shares buffers between modules with runtimes of their own. */
int glb_device_sync(glb_ctx_t ctx); /* command_queue_.finish() */
int glb_ctx_stream(glb_ctx_t ctx, void **cuda_stream);
/* Per-kernel device timing for roofline accounting (the queues of the reference are created
 * with CL_QUEUE_PROFILING_ENABLE, base_module.h:127).  While enabled, glb_spmv / glb_spmv_fused
 * bracket the main kernel and the fix-up kernel with CUDA events on the context's stream.
 * glb_ctx_kernel_timing_read synchronises and returns, summed since the last read:
 * out[0] = ms in the SpMV main kernel, out[1] = ms in the fix-up kernel, out[2] = launches timed. */
int glb_ctx_kernel_timing(glb_ctx_t ctx, int enable);
int glb_ctx_kernel_timing_read(glb_ctx_t ctx, double out[3]);

/* ------------------------------------------------------------------ buffers -------
 * cl::Buffer + enqueueMigrateMemObjects (e.g. spmv_module.h:424-459, 229-253). */
int glb_buffer_alloc(glb_ctx_t ctx, size_t bytes, void **dptr);
int glb_buffer_free(glb_ctx_t ctx, void *dptr);
int glb_buffer_h2d(glb_ctx_t ctx, void *dst_dev, const void *src_host, size_t bytes); /* blocking */
int glb_buffer_d2h(glb_ctx_t ctx, void *dst_host, const void *src_dev, size_t bytes); /* blocking */
int glb_buffer_h2d_async(glb_ctx_t ctx, void *dst_dev, const void *src_host, size_t bytes);
int glb_buffer_d2h_async(glb_ctx_t ctx, void *dst_host, const void *src_dev, size_t bytes);
/* BaseModule::copy_buffer_device_to_device, base_module.h:82-85 (stream-ordered). */
int glb_buffer_d2d(glb_ctx_t ctx, void *dst_dev, const void *src_dev, size_t bytes);
int glb_buffer_fill_f32(glb_ctx_t ctx, float *dst_dev, float val, size_t n);
/* dst[i] = val except dst[index] = index_val: the start vectors of the apps (bfs.h:108-112,
 * sssp.h:153-156, sssp.h:172-176) built on the device in one launch. */
int glb_buffer_fill_one_f32(glb_ctx_t ctx, float *dst_dev, float val, size_t n, size_t index, float index_val);
/* page-locked host memory: the role of xcl2's aligned_allocator (xcl2.hpp:61-76). */
int glb_host_alloc(size_t bytes, void **hptr);
int glb_host_free(void *hptr);

/* ------------------------------------------------------------------ matrices ------
 * SpMVModule::load_and_format_matrix + send_matrix_host_to_device (spmv_module.h:282-420):
 * takes a host CSR (CSRMatrix<float>, data_loader.h:19-31), builds the lane-segment layout
 * for rows [row_begin, row_end) and uploads it.  Row ids stay GLOBAL: glb_spmv writes
 * y[r] for r in the shard only, so ranks of a row-sharded run fill disjoint slices of one
 * full-length vector.  Pass row_begin = 0, row_end = num_rows for the whole matrix. */
int glb_csr_create(glb_ctx_t ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr,
                   const uint32_t *indices, const float *data, uint32_t row_begin, uint32_t row_end,
                   glb_csr_t *out);
int glb_csr_destroy(glb_csr_t m);
/* info[0]=rows in shard, [1]=num_cols, [2]=nnz in shard, [3]=chunks, [4]=fix-up rows,
 * [5]=empty rows in shard, [6]=device bytes of the layout, [7]=hot columns (tile_k). */
int glb_csr_info(glb_csr_t m, uint64_t info[8]);

/* Host-only view of the lane-segment layout glb_csr_create builds (no CUDA call), so the
 * formatter -- the counterpart of csr2cpsr, data_formatter.h:457-534 -- can be checked on a
 * machine without a GPU.  The nnz stream is cut into chunks of up to `max_groups` groups of
 * `group` non-zeros (one warp per chunk); in a chunk of n groups lane l owns the 4n consecutive
 * non-zeros [4n*l, 4n*(l+1)), stored transposed:
 *   stream[256*G + 4*l + e]        encoded column of element 4g+e of lane l (G = chunk_goff[c] + g)
 *   stream[256*G + 128 + 4*l + e]  bit pattern of its fp32 value (0 when `data` was NULL)
 *   encoded column                 rank (< tile_k) in hot_cols of a hot column, else tile_k + column;
 *                                  with tile_k == num_cols (n_hot == 0) the numbering is the identity
 *   flags[32*c + l]                bit r = element r of lane l starts a row (never the chunk's first
 *                                  non-zero); the first padding element of a chunk is flagged too
 *   chunk_first[c]                 ordinal into nz_rows of the row open at the chunk start,
 *                                  bit31 = that row starts exactly there
 *   fixups                         n_fixups x {row, c_begin, c_end}: the row's value is
 *                                  tail_carry[c_begin .. c_end - 1] (+) head_carry[c_end] when bit31 of
 *                                  c_end is set, else tail_carry[c_begin .. c_end]
 * A chunk holds at most row_cap flags. */
typedef struct glb_host_layout {
    uint32_t group, max_groups, row_cap;
    uint64_t nnz;
    uint32_t n_chunks, n_groups, n_nz_rows, n_empty, n_fixups, tile_k, n_hot;
    uint32_t *stream;      /* 256 * n_groups */
    uint32_t *flags;       /* 32 * n_chunks */
    uint32_t *chunk_goff;  /* n_chunks + 1 */
    uint32_t *chunk_first; /* n_chunks */
    uint32_t *nz_rows;     /* n_nz_rows */
    uint32_t *empty_rows;  /* n_empty */
    uint32_t *fixups;      /* 3 * n_fixups */
    uint32_t *hot_cols;    /* n_hot */
} glb_host_layout_t;
int glb_csr_format_host(uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                        const float *data, uint32_t row_begin, uint32_t row_end, uint32_t tile_k,
                        glb_host_layout_t *out);
int glb_host_layout_free(glb_host_layout_t *layout);

/* SpMSpVModule::load_and_format_matrix + send_matrix_host_to_device (spmspv_module.h:264-370):
 * host CSC (CSCMatrix<float>, data_loader.h:92-104; indptr has num_cols+1 entries). */
int glb_csc_create(glb_ctx_t ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr,
                   const uint32_t *indices, const float *data, glb_csc_t *out);
/* Row shard of the push direction (multi-GPU): keeps, of every column, the entries whose row lies
 * in [row_begin, row_end) -- the reference's row tiles (data_formatter.h:621,665) with the tile =
 * the GPU.  glb_spmspv on it lists only rows of the shard; row ids stay global. */
int glb_csc_create_rows(glb_ctx_t ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr,
                        const uint32_t *indices, const float *data, uint32_t row_begin, uint32_t row_end, glb_csc_t *out);
int glb_csc_destroy(glb_csc_t m);

/* ------------------------------------------------------------------ overlay mode 1
 * kernel_spmv (kernel_spmv_impl.h:392-819) == SpMVModule::compute_reference_results
 * (spmv_module.h:478-532): y[r] = zero (+) SUM_{i in row r} data[i] (x) x[indices[i]], then
 *   WRITE_TO_ZERO: mask[r] != 0 -> y[r] = 0     WRITE_TO_ONE: mask[r] == 0 -> y[r] = 0
 * (literal 0, not `zero`).  x has num_cols entries, mask / y num_rows entries; y must not
 * alias x.  `mask` may be NULL iff mask_type == GLB_MASK_NONE. */
int glb_spmv(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x,
             const float *mask, float *y);

/* The per-iteration launch sequences of the apps, folded into the SpMV write-back:
 *   v = masked SpMV row result (as glb_spmv)
 *   if (add_enable)     v = v + add_val                 eWiseAddModule::run   (pagerank.h:86-88)
 *   y[r] = v
 *   if (assign_inout)   DenseAssign with mask := v       AssignVectorDenseModule::run (bfs.h:117-124)
 *        WRITE_TO_ONE : v != 0 -> assign_inout[r] = assign_val
 *        WRITE_TO_ZERO: v == 0 -> assign_inout[r] = assign_val
 * `assign_inout` may alias `mask` (BFS: both are the distance vector). */
typedef struct glb_spmv_epilogue {
    int add_enable;
    float add_val;
    float *assign_inout;
    float assign_val;
    int assign_mask_type;
} glb_spmv_epilogue_t;
int glb_spmv_fused(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x,
                   const float *mask, float *y, const glb_spmv_epilogue_t *ep);

/* End-to-end convenience with HOST vectors (what SpMVModule does around run():
 * send_vector_host_to_device, [send_mask_host_to_device], run, send_results_device_to_host).
 * x_host / mask_host / y_host should be page-locked (glb_host_alloc) for full PCIe speed. */
int glb_spmv_host(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x_host,
                  const float *mask_host, float *y_host);

/* The same sequence for n_vectors independent vectors (the loop of bench_spmv.cpp:96-104 with a
 * fresh host vector per run), pipelined: upload of vector k+1, kernels of vector k and download of
 * result k-1 run concurrently on three streams over two device slots.  y_hosts[k] receives rows
 * [row_begin, row_end) of A (+).(x) x_hosts[k]; mask_hosts may be NULL when mask_type is kNoMask.
 * Returns after the last result has landed. */
int glb_spmv_host_batch(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, int n_vectors,
                        const float *const *x_hosts, const float *const *mask_hosts, float *const *y_hosts);

/* ------------------------------------------------------------------ overlay mode 2
 * kernel_spmspv (kernel_spmspv_impl.h:448-562) == SpMSpVModule::compute_reference_results
 * (spmspv_module.h:446-520) followed by the device's sparse write-back: for each active
 * (c, v) of x and each (r, a) of CSC column c: acc[r] = acc[r] (+) a (x) v  (ADD_MIN clamps at
 * 999999999); rows failing the mask (compared against `zero`) are dropped; `y` receives
 * {count, zero} then every {r, acc[r]} with acc[r] != zero (kernel_spmspv_impl.h:200-229,
 * 263-281,551-555).  x: device list, capacity >= nnz+1; y: capacity num_rows+1. */
int glb_spmspv(glb_ctx_t ctx, glb_csc_t m, int op, float zero, int mask_type, const glb_idx_val_t *x,
               const float *mask, glb_idx_val_t *y);
/* A whole push level in one launch: glb_spmspv, then on the listed entries the sparse assign of the
 * apps, fused into the kernel's compaction phase:
 *   GLB_SPMSPV_EP_ASSIGN  inout[row] = val                          AssignVectorSparse (no new frontier), bfs.h:147-151
 *   GLB_SPMSPV_EP_RELAX   if inout[row] > v { inout[row] = v; append {row, v} to new_frontier }
 *                         AssignVectorSparse (new frontier), sssp.h:178-190; new_frontier[0] = {count, 0}
 * `inout` may alias `mask` (BFS: both are the distance vector).
 * `next` (optional) = the direction decision of pull_push, taken ON THE DEVICE after this level
 * (bfs.h:186-190, sssp.h:210-214): keep pushing iff !force_stop && float(count(y)) / num_vertices < threshold.
 * The decision goes to the IF / ELSE node `cond_next` of a recorded sequence (glb_graph_cond_create; 0: none)
 * and can be read back with glb_spmspv_push_state.  When pushing stops the same launch prepares the dense
 * input of the first pull level: GLB_SPMSPV_DENSE_SCATTER scatters y into `dense` (which the caller
 * filled with the semiring zero beforehand), GLB_SPMSPV_DENSE_COPY copies dense_src (the distance
 * vector, sssp.h:219-222) into it. */
#define GLB_SPMSPV_EP_NONE 0
#define GLB_SPMSPV_EP_ASSIGN 1
#define GLB_SPMSPV_EP_RELAX 2
#define GLB_SPMSPV_DENSE_NONE 0
#define GLB_SPMSPV_DENSE_SCATTER 1
#define GLB_SPMSPV_DENSE_COPY 2
typedef struct glb_spmspv_epilogue {
    int mode;
    float *inout;
    float val;
    glb_idx_val_t *new_frontier;
} glb_spmspv_epilogue_t;
typedef struct glb_spmspv_next {
    int force_stop;
    float threshold;
    uint32_t num_vertices;
    uint64_t cond_next;
    int dense_mode;
    float *dense;
    const float *dense_src;
    uint32_t dense_len;
} glb_spmspv_next_t;
int glb_spmspv_fused(glb_ctx_t ctx, glb_csc_t m, int op, float zero, int mask_type, const glb_idx_val_t *x,
                     const float *mask, glb_idx_val_t *y, const glb_spmspv_epilogue_t *ep,
                     const glb_spmspv_next_t *next);
/* Blocking read of the last decision (1 = keep pushing) and of the number of levels that carried a
 * `next` block since glb_spmspv_reset_levels (stream-ordered) -- push_iterations of a pull_push run. */
int glb_spmspv_push_state(glb_ctx_t ctx, glb_csc_t m, uint32_t *keep_pushing, uint32_t *push_levels);
int glb_spmspv_reset_levels(glb_ctx_t ctx, glb_csc_t m);
/* list = {1, 0}, {index, val}: the one-entry start frontier of the push apps (bfs.h:131-135, sssp.h:169-171),
 * built on the device (stream-ordered) instead of uploaded with a blocking copy. */
int glb_sparse_fill_one(glb_ctx_t ctx, glb_idx_val_t *list, uint32_t index, float val);
/* SpMSpVModule::get_results_nnz (spmspv_module.h:239-242): blocking read of list[0].index. */
int glb_sparse_count(glb_ctx_t ctx, const glb_idx_val_t *list, uint32_t *count);
/* convert_sparse_vec_to_dense_vec (global.h:153-164) done on the device (the reference does it
 * on the host at the push->pull switch, bfs.h:195-201). dense[0..len) = zero, then scatter. */
int glb_sparse_to_dense(glb_ctx_t ctx, const glb_idx_val_t *list, float *dense, uint32_t len, float zero);
/* The row-sharded push direction exchanges its frontier as a dense vector (north_star: one dense
 * exchange per iteration): _rows resets rows [row_begin, row_end) of `dense` to `zero` and scatters the
 * list entries inside that range; glb_dense_to_sparse lists the entries != zero of a dense vector
 * ({count, zero} head, order unspecified). */
int glb_sparse_to_dense_rows(glb_ctx_t ctx, const glb_idx_val_t *list, float *dense, uint32_t row_begin,
                             uint32_t row_end, float zero);
int glb_dense_to_sparse(glb_ctx_t ctx, const float *dense, uint32_t len, float zero, glb_idx_val_t *list);

/* ------------------------------------------------------------------ overlay mode 3
 * kernel_add_scalar_vector_dense_impl.h:6-27: out[i] = in[i] + val. in may equal out. */
int glb_ewise_add(glb_ctx_t ctx, const float *in, float *out, uint32_t len, float val);

/* ------------------------------------------------------------------ overlay mode 4
 * kernel_assign_vector_dense_impl.h:8-47: WRITE_TO_ZERO: mask[i]==0 -> inout[i]=val;
 * WRITE_TO_ONE: mask[i]!=0 -> inout[i]=val; GLB_MASK_NONE -> GLB_EINVAL. */
int glb_assign_dense(glb_ctx_t ctx, const float *mask, float *inout, uint32_t len, float val, int mask_type);

/* ------------------------------------------------------------------ overlay mode 5
 * kernel_assign_vector_sparse_no_new_frontier_impl.h:4-55:
 * inout[list[i+1].index] = val for i < list[0].index. */
int glb_assign_sparse(glb_ctx_t ctx, const glb_idx_val_t *list, float *inout, float val);

/* ------------------------------------------------------------------ overlay mode 6
 * kernel_assign_vector_sparse_new_frontier_impl.h:4-78: for each list entry {i, v}:
 * if inout[i] > v { inout[i] = v; emit {i, v} }; new_frontier[0] = {count, 0}.
 * `capacity` = number of glb_idx_val_t slots in new_frontier (>= list count + 1).
 * Indices in `list` are unique (it is an SpMSpV result), so the emitted SET equals the
 * reference's; order is unspecified. new_frontier must not alias list. */
int glb_assign_sparse_relax(glb_ctx_t ctx, const glb_idx_val_t *list, float *inout,
                            glb_idx_val_t *new_frontier);

/* ------------------------------------------------------------------ value-type variants
 * The same operators on the other two `val_t` choices of the reference (global.h:60-64): GLB_VAL_U32 and
 * GLB_VAL_UFIXED (the Q8.24 of the shipped bitstream), with the arithmetic of its processing elements
 * (ufixed_pe_fwd.h:23-65; semiring.cuh).  Vectors, masks and matrix values are 32-bit WORDS of that type
 * (glb_csr_create / glb_csc_create take them through their `data` pointer unchanged); scalars travel as
 * words: zero_bits, val_bits, and the add_val / assign_val members of the epilogue hold the word's bits.
 * Mask tests compare against the word 0 (SpMV, dense assign) or against `zero` (SpMSpV), as in the fp32
 * entry points.  Exact integer arithmetic: results do not depend on the reduction order.
 * The reference ships no artefact that pins its device numerics (no bitstream, no emulator), so parity of
 * these variants is against a software model of the documented ap_ufixed semantics (oracle/valtype_model.h),
 * not against reference output: "parity unpinned". */
int glb_spmv_vt(glb_ctx_t ctx, glb_csr_t m, int val_type, int op, uint32_t zero_bits, int mask_type, const void *x,
                const void *mask, void *y, const glb_spmv_epilogue_t *ep);
int glb_spmspv_vt(glb_ctx_t ctx, glb_csc_t m, int val_type, int op, uint32_t zero_bits, int mask_type,
                  const glb_idx_val_t *x, const void *mask, glb_idx_val_t *y);
int glb_ewise_add_vt(glb_ctx_t ctx, int val_type, const void *in, void *out, uint32_t len, uint32_t val_bits);
int glb_assign_dense_vt(glb_ctx_t ctx, int val_type, const void *mask, void *inout, uint32_t len, uint32_t val_bits,
                        int mask_type);
int glb_assign_sparse_vt(glb_ctx_t ctx, int val_type, const glb_idx_val_t *list, void *inout, uint32_t val_bits);
int glb_assign_sparse_relax_vt(glb_ctx_t ctx, int val_type, const glb_idx_val_t *list, void *inout,
                               glb_idx_val_t *new_frontier);

/* ------------------------------------------------------------------ multi-GPU ------
 * Row-range sharding: one process per GPU, each holding glb_csr_create(..., row_begin, row_end)
 * and a full-length x.  After glb_spmv every rank owns y[row_begin:row_end); one allgather
 * makes the next x.  The host may do that exchange with its own communicator on the same
 * stream (torch.distributed in bench.py), or use the built-in NCCL binding below. */
#define GLB_NCCL_UNIQUE_ID_BYTES 128
int glb_nccl_available(void);
int glb_nccl_unique_id(void *id128);
int glb_comm_init(glb_ctx_t ctx, const void *id128, int rank, int nranks);
int glb_comm_destroy(glb_ctx_t ctx);
/* In-place allgather: every rank contributed buf[rank*count_per_rank .. +count_per_rank). */
int glb_allgather_f32(glb_ctx_t ctx, float *buf, size_t count_per_rank);

/* Fused exchange.  The allgather moves each rank's slice of y to every rank AFTER the SpMV; here
 * the SpMV write-back itself stores every row into all ranks' copies of the vector over NVLink
 * (peer-mapped memory, CUDA IPC across the per-GPU processes), so the transfer overlaps the
 * kernel and what remains of the collective is one signal / wait kernel.
 *   glb_xchg_create   allocates this rank's block: n_vectors (2..4) vectors of n_floats + flags
 *   glb_xchg_export   64-byte IPC handle of the block; the host exchanges the handles of all ranks
 *   glb_xchg_connect  maps the peers' blocks (handles = nranks x 64 bytes, in rank order)
 *   glb_xchg_vector   local device pointer of vector `which` (fill vector 0 before the first step)
 *   glb_spmv_exchange y = A (+).(x) x with x = local vector src_vec, y rows written into vector
 *                     dst_vec of EVERY rank; returns after enqueueing the signal / wait, so the
 *                     next call on this stream sees the complete dst_vec.  Alternate src / dst.
 *   glb_xchg_allgather every rank's slice [offset, offset + count) of vector `which` is copied into
 *                     all peers' copies (peer-to-peer copies on the stream) + signal / wait: what
 *                     ncclAllGather does for vectors the SpMV did not write (BFS distance at the end)
 *   glb_xchg_barrier  signal / wait alone: no rank's later work on its stream starts before every
 *                     rank's earlier work has finished (call between two runs that reuse the vectors)
 *   glb_xchg_status   synchronises; *timed_out != 0 if a peer never signalled (it died) */
#define GLB_IPC_HANDLE_BYTES 64
typedef struct glb_xchg_s *glb_xchg_t;
int glb_xchg_create(glb_ctx_t ctx, uint32_t n_floats, int n_vectors, glb_xchg_t *out);
int glb_xchg_export(glb_xchg_t xc, void *handle64);
int glb_xchg_connect(glb_xchg_t xc, int rank, int nranks, const void *handles);
/* Blocks mapped by the host instead of CUDA IPC (e.g. torch symmetric memory): blocks[r] = rank r's
 * block as mapped in THIS process (glb_xchg_block_bytes each, zero-filled), multicast_block = the
 * multicast mapping of the same blocks or NULL.  With a multicast mapping the slices travel as
 * multimem.st through the NVSwitch: one store lands on every rank. */
size_t glb_xchg_block_bytes(uint32_t n_floats, int n_vectors);
int glb_xchg_adopt(glb_ctx_t ctx, uint32_t n_floats, int n_vectors, int rank, int nranks, void *const *blocks,
                   void *multicast_block, glb_xchg_t *out);
int glb_xchg_has_multicast(glb_xchg_t xc);
/* The same exchange with the NVSwitch multicast object, this rank's block and both mappings made by the library
 * itself (CUDA driver virtual-memory API; no framework):
 *   rank 0:       glb_xchg_mc_open(ctx, n, v, 0, nranks, -1, &xc, &fd) creates the multicast object and returns its POSIX
 *                 file descriptor; the host hands it to the other ranks' processes (SCM_RIGHTS, pidfd_getfd, ...)
 *   other ranks:  glb_xchg_mc_open(ctx, n, v, rank, nranks, fd_in_this_process, &xc, NULL)
 *   -- host barrier (every rank has added its device to the object) --
 *   all ranks:    glb_xchg_mc_bind(xc): allocate, bind, map; the exchange is connected (multicast transfers only)
 *   -- host barrier, then use like any other exchange --
 * glb_xchg_mc_supported: *supported != 0 if the device and driver offer multicast objects. */
int glb_xchg_mc_supported(glb_ctx_t ctx, int *supported);
int glb_xchg_mc_open(glb_ctx_t ctx, uint32_t n_floats, int n_vectors, int rank, int nranks, int fd_in, glb_xchg_t *out, int *fd_out);
int glb_xchg_mc_bind(glb_xchg_t xc);
int glb_xchg_vector(glb_xchg_t xc, int which, float **local_ptr);
int glb_xchg_allgather(glb_ctx_t ctx, glb_xchg_t xc, int which, size_t offset, size_t count);
int glb_xchg_barrier(glb_ctx_t ctx, glb_xchg_t xc);
int glb_xchg_status(glb_xchg_t xc, int *timed_out);
int glb_xchg_destroy(glb_xchg_t xc);
int glb_spmv_exchange(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, glb_xchg_t xc, int src_vec,
                      int dst_vec, const float *mask, const glb_spmv_epilogue_t *ep);
/* n_steps iterations of a pull loop (pagerank.h:80-90, bfs.h:117-124, sssp.h:157-163) in one call: step k
 * reads vector src_vec (k even) / dst_vec (k odd) and writes the other one on every rank; eps (NULL or
 * n_steps entries) = the fused epilogue of each step.  Between steps the acquire of the exchange rides
 * in the head of the next SpMV's first kernel instead of a launch of its own; the call ends with the
 * acquire of the last step.  Graph-recordable (glb_graph_begin).  vec_plan (NULL, or 2 * n_steps vector
 * indices {read, written} per step) replaces the ping-pong, e.g. a stream of independent input vectors. */
int glb_spmv_exchange_iterate(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, glb_xchg_t xc, int src_vec,
                              int dst_vec, const float *mask, const glb_spmv_epilogue_t *eps, int n_steps,
                              const int *vec_plan);
/* glb_spmv_host_batch for a row-sharded run: called by every rank with the same full-length host
 * vectors; each rank uploads only its 1/nranks slice of x over PCIe, the slices meet over NVLink
 * (exchange vectors 0 and 1 are the two pipeline slots), and y_hosts[k] receives the rank's rows. */
int glb_spmv_host_batch_exchange(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, glb_xchg_t xc,
                                 int n_vectors, const float *const *x_hosts, const float *const *mask_hosts,
                                 float *const *y_hosts);

/* ------------------------------------------------------------------ launch replay ------
 * The reference drains its queue after every module run (spmv_module.h:471-475); an app
 * iteration is 2-3 such launches (bfs.h:117-124).  Here launches are stream-ordered, and a fixed
 * sequence (the iteration loop of an app: same buffers, same per-iteration scalars) can be
 * recorded once as a CUDA graph and replayed with one call:
 *   glb_graph_begin   every glb_* launch on this context that follows is recorded, not executed
 *                     (no blocking call -- copies to / from the host, glb_ctx_sync, glb_sparse_count,
 *                     buffer allocation -- may be made while recording; the exchange steps of a
 *                     row-sharded run ARE recordable: their epoch lives in device memory)
 *   glb_graph_end     stops recording and returns the executable sequence
 *   glb_graph_launch  enqueues the whole sequence on the context's stream */
typedef struct glb_graph_s *glb_graph_t;
int glb_graph_begin(glb_ctx_t ctx);
/* A branch inside the recorded sequence, decided on the device at replay time (the push / pull choice of
 * pull_push, bfs.h:186-190, without the host reading the frontier size):
 *   glb_graph_cond_create   a condition (0 at the start of every replay); pass it to the launch that
 *                           decides -- glb_spmspv_fused's next->cond_next -- recorded BEFORE the branch
 *   glb_graph_branch_begin  launches that follow are recorded into the IF arm (condition != 0) ...
 *   glb_graph_branch_else   ... from here on into the ELSE arm (condition == 0) ...
 *   glb_graph_branch_end    ... and from here on after the branch again.  Branches do not nest. */
int glb_graph_cond_create(glb_ctx_t ctx, uint64_t *cond);
int glb_graph_branch_begin(glb_ctx_t ctx, uint64_t cond);
int glb_graph_branch_else(glb_ctx_t ctx);
int glb_graph_branch_end(glb_ctx_t ctx);
int glb_graph_end(glb_ctx_t ctx, glb_graph_t *out);
int glb_graph_launch(glb_ctx_t ctx, glb_graph_t g);
int glb_graph_destroy(glb_graph_t g);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHLILY_B200_H_ */
