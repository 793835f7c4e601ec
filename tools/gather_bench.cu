// Microbenchmark: how fast can one B200 gather 4-byte words at random from a vector that lives
// in L2 (the x of SpMV)?  Decides the SpMV gather path: plain LDG, LDG with cache hints, the
// texture path, shared memory for a hot prefix, or a mix.  Build (in-tree, travels with gpurun):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/gather_bench tools/gather_bench.cu
// Run: tools/gather_bench [n_cols=4194304] [gathers=134217728]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

enum { M_LDG = 0, M_LDG_NC_NOALLOC, M_LDG_CG, M_TEX, M_SMEM, M_MIX, M_LDG_IDX, M_MIX_IDX, M_L1MIX_IDX, M_L1MIX2_IDX };

// Each thread performs `per_thread` gathers; the index is a hash (ALU only, no index stream), or
// read from a coalesced uint4 stream (IDX modes: the real SpMV shape, 4 per lane per step).
template <int MODE>
__global__ void __launch_bounds__(1024) gather_kernel(const float *__restrict__ x, cudaTextureObject_t tex, uint32_t n_mask,
                                                     uint32_t tile_k, const uint4 *__restrict__ idx, uint32_t per_thread,
                                                     float *out, uint32_t hot_permille) {
    extern __shared__ float tile[];
    if (MODE == M_SMEM || MODE == M_MIX || MODE == M_MIX_IDX) {
        for (uint32_t i = threadIdx.x; i < tile_k; i += blockDim.x) tile[i] = x[i];
        __syncthreads();
    }
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nthreads = gridDim.x * blockDim.x;
    float acc = 0.f;
    if (MODE == M_LDG_IDX || MODE == M_MIX_IDX || MODE == M_L1MIX_IDX || MODE == M_L1MIX2_IDX) {
        for (uint32_t it = 0; it < per_thread / 4; ++it) {
            const uint4 c = __ldcs(idx + size_t(it) * nthreads + tid);
            const uint32_t cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == M_MIX_IDX) acc += (cc[j] < tile_k) ? tile[cc[j]] : __ldg(x + cc[j]);
                else if (MODE == M_L1MIX_IDX || MODE == M_L1MIX2_IDX) {
                    // hot prefix through L1 (allocating, evict_last), cold around it (no_allocate)
                    float v;
                    if (cc[j] < tile_k) {
                        if (MODE == M_L1MIX2_IDX) asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(x + cc[j]));
                        else v = __ldg(x + cc[j]);
                    } else asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(x + cc[j]));
                    acc += v;
                }
                else acc += __ldg(x + cc[j]);
            }
        }
    } else {
#pragma unroll 4
        for (uint32_t it = 0; it < per_thread; ++it) {
            const uint32_t h = hash32(tid * 2654435761u + it * 40503u + 17u);
            uint32_t c = h & n_mask;
            float v;
            if (MODE == M_LDG) v = __ldg(x + c);
            else if (MODE == M_LDG_NC_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(x + c));
            else if (MODE == M_LDG_CG) v = __ldcg(x + c);
            else if (MODE == M_TEX) v = tex1Dfetch<float>(tex, int(c));
            else if (MODE == M_SMEM) v = tile[c % tile_k];
            else {  // M_MIX: hot_permille of the gathers go to the shared-memory tile
                const bool hot = ((h >> 22) % 1000u) < hot_permille;
                v = hot ? tile[c % tile_k] : __ldg(x + c);
            }
            acc += v;
        }
    }
    if (acc == 123.456f) out[tid] = acc;
}

template <int MODE>
float run(const char *name, const float *x, cudaTextureObject_t tex, uint32_t n, uint32_t tile_k, const uint4 *idx,
          uint64_t gathers, float *out, int threads, int blocks_per_sm, uint32_t hot_permille, int sms) {
    const size_t smem = (MODE == M_SMEM || MODE == M_MIX || MODE == M_MIX_IDX) ? size_t(tile_k) * 4 : 0;
    CK(cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int grid = sms * blocks_per_sm;
    const uint32_t per_thread = uint32_t(gathers / (uint64_t(grid) * threads)) & ~3u;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) gather_kernel<MODE><<<grid, threads, smem>>>(x, tex, n - 1, tile_k, idx, per_thread, out, hot_permille);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int i = 0; i < reps; ++i) gather_kernel<MODE><<<grid, threads, smem>>>(x, tex, n - 1, tile_k, idx, per_thread, out, hot_permille);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double g = double(per_thread) * grid * threads;
    printf("%-34s threads %4d x %d/SM hot %4u  %8.1f us  %7.1f Ggather/s  (%.2f gathers/clk/SM @1.9GHz)\n", name, threads,
           blocks_per_sm, hot_permille, ms * 1e3, g / ms / 1e6, g / ms / 1e6 / sms / 1.9);
    return ms;
}

__global__ void fill_idx(uint32_t *idx, size_t n, uint32_t n_cols, uint32_t tile_k, uint32_t hot_permille) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const uint32_t h = hash32(uint32_t(i) * 2654435761u + 99u), h2 = hash32(h + 0x9e3779b9u);
        const bool hot = (h2 % 1000u) < hot_permille;
        idx[i] = hot ? (h % tile_k) : (tile_k + h % (n_cols - tile_k));
    }
}

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? uint32_t(atol(argv[1])) : 4194304u;   // power of two
    const uint64_t gathers = argc > 2 ? uint64_t(atoll(argv[2])) : 134217728ull;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs; vector %u floats (%.1f MB), %llu gathers\n", prop.name, sms, n, n * 4 / 1e6,
           (unsigned long long)gathers);
    float *x, *out;
    CK(cudaMalloc(&x, size_t(n) * 4));
    CK(cudaMemset(x, 0, size_t(n) * 4));
    CK(cudaMalloc(&out, size_t(sms) * 16 * 1024 * 4));
    uint32_t *idx;
    CK(cudaMalloc(&idx, gathers * 4));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = x;
    rd.res.linear.desc = cudaCreateChannelDesc<float>();
    rd.res.linear.sizeInBytes = size_t(n) * 4;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const uint32_t tile_k = 49152;

    for (int bps : {1, 2}) {
        run<M_LDG>("ldg (ld.global.nc)", x, tex, n, tile_k, nullptr, gathers, out, 1024, bps, 0, sms);
        run<M_LDG_NC_NOALLOC>("ldg nc L1::no_allocate", x, tex, n, tile_k, nullptr, gathers, out, 1024, bps, 0, sms);
        run<M_LDG_CG>("ldcg (L2 only)", x, tex, n, tile_k, nullptr, gathers, out, 1024, bps, 0, sms);
        run<M_TEX>("tex1Dfetch", x, tex, n, tile_k, nullptr, gathers, out, 1024, bps, 0, sms);
    }
    run<M_LDG>("ldg, vector fits L1 (64 KB)", x, tex, 16384, tile_k, nullptr, gathers, out, 1024, 2, 0, sms);
    run<M_LDG>("ldg, vector 1 MB", x, tex, 262144, tile_k, nullptr, gathers, out, 1024, 2, 0, sms);
    run<M_SMEM>("smem tile 48K floats", x, tex, n, tile_k, nullptr, gathers, out, 1024, 1, 0, sms);
    for (uint32_t hot : {0u, 300u, 540u, 700u, 850u})
        run<M_MIX>("mix smem/ldg (hash idx)", x, tex, n, tile_k, nullptr, gathers, out, 1024, 1, hot, sms);
    for (uint32_t hot : {0u, 460u, 540u, 850u}) {
        fill_idx<<<sms * 8, 256>>>(idx, gathers, n, tile_k, hot);
        CK(cudaDeviceSynchronize());
        run<M_LDG_IDX>("ldg, idx stream (uint4/lane)", x, tex, n, tile_k, reinterpret_cast<const uint4 *>(idx), gathers, out, 1024, 2, hot, sms);
        run<M_MIX_IDX>("mix smem/ldg, idx stream", x, tex, n, tile_k, reinterpret_cast<const uint4 *>(idx), gathers, out, 1024, 1, hot, sms);
        run<M_MIX_IDX>("mix smem/ldg, idx stream", x, tex, n, tile_k, reinterpret_cast<const uint4 *>(idx), gathers, out, 512, 1, hot, sms);
        run<M_L1MIX_IDX>("L1 hot(ldg)/cold(no_alloc), idx", x, tex, n, tile_k, reinterpret_cast<const uint4 *>(idx), gathers, out, 1024, 2, hot, sms);
        run<M_L1MIX2_IDX>("L1 hot(evict_last)/cold(no_alloc)", x, tex, n, tile_k, reinterpret_cast<const uint4 *>(idx), gathers, out, 1024, 2, hot, sms);
    }
    return 0;
}
