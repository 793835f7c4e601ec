// Device side of the exchange's acquire step, shared by exchange.cu and the kernels that open an
// SpMV launch (spmv.cu): "every rank's slice of the previous step has landed in my copy of x".
//
// Each rank publishes a monotonically increasing epoch into flag word [rank] of every rank's block
// after its slice has been stored there (exchange.cu); a consumer spins until all flag words of its
// own block have reached the epoch this rank itself published last (state[0], device resident so
// that recorded launch sequences replay with the right value).  The wait sits at the head of the
// kernel that first reads x in the next step, so it costs no launch of its own.
#ifndef GLB_EXCHANGE_CUH_
#define GLB_EXCHANGE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

struct GlbXchgWait {
    const uint32_t *flags;  // this rank's flag words, one per rank (NULL: nothing to wait for)
    const uint32_t *state;  // [0] = epoch published last by this rank
    uint32_t *err;          // set before the trap when a peer never signalled
    int n;                  // ranks
};

__device__ __forceinline__ void glb_xchg_spin(const uint32_t *flag, uint32_t epoch, uint32_t *err) {
    const long long t0 = clock64();
    uint32_t seen;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if (int32_t(seen - epoch) >= 0) break;
        if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer died or hangs.  Never continue as if the
            *err = 1;                          // slices had arrived: the trap fails this stream's next
            __threadfence_system();            // synchronisation (and every later call) loudly
            asm volatile("trap;");
        }
    }
}

// Call with all threads of the CTA; blockDim.x >= w.n (at most 8 ranks).
__device__ __forceinline__ void glb_xchg_wait_head(const GlbXchgWait &w) {
    if (w.flags == nullptr) return;  // uniform
    if (int(threadIdx.x) < w.n) {
        const uint32_t epoch = *reinterpret_cast<const volatile uint32_t *>(w.state);
        glb_xchg_spin(w.flags + threadIdx.x, epoch, w.err);
    }
    __syncthreads();
}

#endif  // GLB_EXCHANGE_CUH_
