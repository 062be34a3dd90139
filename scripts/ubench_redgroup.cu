// ubench_redgroup.cu -- over how many lanes does B200 merge the accesses of ONE warp instruction?
// Pairs of lanes (l, l ^ dist) of a warp address the same random 32-byte sector (different 8-byte entries) of a 48 MB table; every
// other lane pair has its own sector.  If the hardware coalesces over the whole warp the rate is independent of `dist`; if it merges
// per group of 8 (or 16) lanes the rate drops to the no-sharing rate once dist >= 8 (16).  Measured for red.global.add.v2.f32 (the
// backward's scatter) and for 8-byte gathers (the forward), to size the sector / line model of DESIGN.md 3.1.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_redgroup ubench_redgroup.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int OP>  // 0: gather 8 B, 1: red.v2.f32
__global__ void __launch_bounds__(256) k(float* table, uint32_t n_sectors, uint32_t per_thread, uint32_t dist, float* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    // pair id inside the warp: the lane index with bit `dist` removed (dist = 0: no sharing, every lane its own sector)
    const uint32_t pair = dist ? (((lane >> 1) & ~(dist - 1)) | (lane & (dist - 1))) : lane;
    const uint32_t which = dist ? ((lane / dist) & 1u) : 0u;
    float acc = 0.f;
#pragma unroll 8
    for (uint32_t it = 0; it < per_thread; ++it) {
        const uint32_t sector = hash32((warp * 32u + pair) * 2654435761u + it * 40503u) % n_sectors;
        const uint32_t slot = sector * 4u + which;      // 8-byte entries, 4 per sector
        if (OP == 0) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(table) + slot);
            acc += v.x + v.y;
        } else {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2ull * slot), "f"(1.f), "f"(2.f) : "memory");
        }
    }
    if (acc == 123.456f) *sink = acc;
}

template <int OP>
float run(float* table, uint32_t n_sectors, uint32_t dist, float* sink, uint64_t total) {
    const uint32_t per_thread = 64, threads = (uint32_t)(total / per_thread);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) k<OP><<<threads / 256, 256>>>(table, n_sectors, per_thread, dist, sink);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) k<OP><<<threads / 256, 256>>>(table, n_sectors, per_thread, dist, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    const uint32_t mb = 48, n_sectors = mb * 1024u * 1024u / 32u;
    float *table, *sink;
    cudaMalloc(&table, (size_t)mb << 20);
    cudaMemset(table, 0, (size_t)mb << 20);
    cudaMalloc(&sink, 4);
    const uint64_t total = 1ull << 28;
    for (uint32_t dist : {0u, 1u, 2u, 4u, 8u, 16u}) {
        const float g = run<0>(table, n_sectors, dist, sink, total), r = run<1>(table, n_sectors, dist, sink, total);
        printf("partner lane distance %2u%s  gather 8B %8.1f G lanes/s   red.v2.f32 %8.1f G lanes/s\n", dist, dist ? "" : " (no sharing)",
               total / g / 1e6, total / r / 1e6);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
