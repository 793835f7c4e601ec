// Microbenchmark: random 4-byte gathers from DISTRIBUTED shared memory (a cluster-wide copy of the
// hot part of x), to see whether the SM-to-SM network can take gather traffic off the L2 fabric.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_bench tools/dsmem_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// remote_permille of the gathers go to a random CTA of the cluster (possibly self), cold_permille to
// global memory (L2-resident x), the rest to the local tile.
__global__ void __launch_bounds__(1024) dsmem_kernel(const float *__restrict__ x, uint32_t n_mask, uint32_t tile_k,
                                                    uint32_t per_thread, float *out, uint32_t remote_permille,
                                                    uint32_t cold_permille, uint32_t csize) {
    extern __shared__ float tile[];
    cg::cluster_group cluster = cg::this_cluster();
    for (uint32_t i = threadIdx.x; i < tile_k; i += blockDim.x) tile[i] = x[i];
    cluster.sync();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t tile_base = uint32_t(__cvta_generic_to_shared(tile));
    float acc = 0.f;
#pragma unroll 4
    for (uint32_t it = 0; it < per_thread; ++it) {
        const uint32_t h = hash32(tid * 2654435761u + it * 40503u + 17u);
        const uint32_t sel = (h >> 20) % 1000u;
        const uint32_t off = (h % tile_k) * 4u;
        float v;
        if (sel < remote_permille) {
            const uint32_t rank = (h >> 12) % csize;
            uint32_t raddr;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(tile_base + off), "r"(rank));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr));
        } else if (sel < remote_permille + cold_permille) {
            asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(x + (h & n_mask)));
        } else {
            v = tile[off >> 2];
        }
        acc += v;
    }
    if (acc == 123.456f) out[tid] = acc;
    cluster.sync();
}

int main(int argc, char **argv) {
    const uint32_t n = 4194304u;
    const uint64_t gathers = 134217728ull;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    float *x, *out;
    CK(cudaMalloc(&x, size_t(n) * 4));
    CK(cudaMemset(x, 0, size_t(n) * 4));
    CK(cudaMalloc(&out, size_t(1) << 24));
    const uint32_t tile_k = 49152;
    const size_t smem = size_t(tile_k) * 4;
    CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (uint32_t csize : {1u, 2u, 4u, 8u, 16u}) {
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(1024);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int max_clusters = 0;
        cfg.gridDim = dim3(csize);
        cudaError_t qe = cudaOccupancyMaxActiveClusters(&max_clusters, dsmem_kernel, &cfg);
        if (qe != cudaSuccess || max_clusters == 0) { printf("cluster %u: not launchable (%s)\n", csize, cudaGetErrorString(qe)); cudaGetLastError(); continue; }
        const uint32_t grid = uint32_t(max_clusters) * csize;
        cfg.gridDim = dim3(grid);
        const uint32_t per_thread = uint32_t(gathers / (uint64_t(grid) * 1024)) & ~3u;
        struct { uint32_t r, c; } mixes[] = {{1000, 0}, {500, 0}, {0, 1000}, {600, 320}, {300, 400}, {200, 350}, {0, 540}};
        for (auto mx : mixes) {
            if (csize == 1 && mx.r && mx.r != 1000) continue;
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            const uint32_t nm = n - 1;
            void *args[] = {(void *)&x, (void *)&nm, (void *)&tile_k, (void *)&per_thread, (void *)&out, (void *)&mx.r, (void *)&mx.c, (void *)&csize};
            for (int i = 0; i < 2; ++i) CK(cudaLaunchKernelExC(&cfg, (const void *)dsmem_kernel, args));
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 3; ++i) CK(cudaLaunchKernelExC(&cfg, (const void *)dsmem_kernel, args));
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
            const double g = double(per_thread) * grid * 1024;
            printf("cluster %2u  CTAs %3u  remote %4u cold %4u local %4u : %8.1f us  %7.1f Ggather/s  %.2f gathers/clk/SM(active)\n", csize, grid,
                   mx.r, mx.c, 1000 - mx.r - mx.c, ms * 1e3, g / ms / 1e6, g / ms / 1e6 / grid / 1.9);
        }
    }
    return 0;
}
