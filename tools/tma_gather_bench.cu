// Microbenchmark: can the TMA unit (cp.async.bulk, 16-byte copies global -> shared) serve random
// gathers IN ADDITION to the LSU path (which tops out at ~1 L1 miss request per clock per SM)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_gather_bench tools/tma_gather_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

constexpr int K = 8;           // gathers per thread per round
constexpr int WARPS = 8;

// tma_k of the K gathers per round go through cp.async.bulk (16 B each), the rest through the LSU.
__global__ void __launch_bounds__(WARPS * 32) tma_kernel(const float *__restrict__ x, uint32_t n_mask, uint32_t rounds,
                                                        float *out, int tma_k) {
    __shared__ __align__(16) float4 slots[WARPS][K][32];
    __shared__ __align__(8) uint64_t bars[WARPS];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[w])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    float acc = 0.f;
    uint32_t phase = 0;
    for (uint32_t r = 0; r < rounds; ++r) {
        if (tma_k > 0) {
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[w])), "r"(uint32_t(tma_k * 32 * 16)) : "memory");
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t h = hash32(tid * 2654435761u + (r * K + k) * 40503u + 17u);
            const uint32_t c = h & n_mask;
            if (k < tma_k) {
                const float *src = x + (c & ~3u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(smem_u32(&slots[w][k][lane])), "l"(src), "r"(smem_u32(&bars[w])) : "memory");
            } else {
                float v;
                asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(x + c));
                acc += v;
            }
        }
        if (tma_k > 0) {
            uint32_t done;
            do {
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bars[w])), "r"(phase) : "memory");
            } while (!done);
            phase ^= 1;
#pragma unroll
            for (int k = 0; k < K; ++k) if (k < tma_k) acc += slots[w][k][lane].x;
            __syncwarp();
        }
    }
    if (acc == 123.456f) out[tid] = acc;
}

int main() {
    const uint32_t n = 4194304u;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float *x, *out;
    CK(cudaMalloc(&x, size_t(n) * 4));
    CK(cudaMemset(x, 0, size_t(n) * 4));
    CK(cudaMalloc(&out, size_t(1) << 26));
    for (int bps : {2, 4, 6}) {
        for (int tma_k : {0, 1, 2, 4, 8}) {
            const int grid = sms * bps, threads = WARPS * 32;
            const uint64_t gathers = 67108864ull;
            const uint32_t rounds = uint32_t(gathers / (uint64_t(grid) * threads * K));
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            for (int i = 0; i < 2; ++i) tma_kernel<<<grid, threads>>>(x, n - 1, rounds, out, tma_k);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 3; ++i) tma_kernel<<<grid, threads>>>(x, n - 1, rounds, out, tma_k);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
            const double g = double(rounds) * K * grid * threads;
            printf("CTAs/SM %d  tma %d/%d of gathers: %8.1f us  %7.1f Ggather/s  %.2f gathers/clk/SM\n", bps, tma_k, K, ms * 1e3,
                   g / ms / 1e6, g / ms / 1e6 / sms / 1.9);
        }
    }
    return 0;
}
