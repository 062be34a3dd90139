// ubench_mem.cu -- micro-benchmarks that size the LoTD kernels' design space on B200:
//   random 8-byte gathers / red.v2 / red.v4 / scalar red into tables of 4 MB (one hash level) and 48 MB (all levels),
//   with and without intra-warp address locality.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mem ubench_mem.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// mode 0: random slot per (thread, k);  mode 1: slots of a warp fall in a window of `window` slots (sorted-points proxy)
template <int OP>  // 0 gather v2, 1 red v2, 2 red v4 (pairs of slots), 3 red scalar x2, 4 gather v4
__global__ void __launch_bounds__(256) k_bench(float* table, uint32_t n_slots, uint32_t per_thread, uint32_t window, float* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t warp = tid >> 5;
    float acc = 0.f;
#pragma unroll 8
    for (uint32_t k = 0; k < per_thread; ++k) {
        uint32_t slot;
        if (window) slot = (hash32(warp * 7919u + k) + (hash32(tid * 31u + k) % window)) % n_slots;
        else slot = hash32(tid * 2654435761u + k * 40503u) % n_slots;
        if (OP == 0) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(table) + slot);
            acc += v.x + v.y;
        } else if (OP == 1) {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2ull * slot), "f"(1.f), "f"(2.f) : "memory");
        } else if (OP == 2) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(table + 4ull * (slot >> 1)), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
        } else if (OP == 3) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(table + 2ull * slot), "f"(1.f) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(table + 2ull * slot + 1), "f"(2.f) : "memory");
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4*>(table) + (slot >> 1));
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

template <int OP>
float run(float* table, uint32_t n_slots, uint32_t window, float* sink, uint64_t total_ops, int occ_blocks) {
    const uint32_t per_thread = 64;
    const uint32_t threads = (uint32_t)(total_ops / per_thread);
    dim3 grid(threads / 256);
    (void)occ_blocks;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) k_bench<OP><<<grid, 256>>>(table, n_slots, per_thread, window, sink);
    cudaEventRecord(a);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) k_bench<OP><<<grid, 256>>>(table, n_slots, per_thread, window, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const uint64_t total = 1ull << 28;  // 268M ops ~ one "8 fine levels x 4M points x 8 corners" pass
    float* sink;
    cudaMalloc(&sink, 4);
    const char* names[5] = {"gather v2 (8B)", "red.v2.f32", "red.v4.f32", "2x red.f32", "gather v4 (16B)"};
    for (uint32_t mb : {4u, 48u, 512u}) {
        const uint32_t n_slots = mb * 1024u * 1024u / 8u;
        float* table;
        cudaMalloc(&table, (size_t)n_slots * 8);
        cudaMemset(table, 0, (size_t)n_slots * 8);
        for (uint32_t window : {0u, 64u, 8u, 1u}) {
            float ms[5];
            ms[0] = run<0>(table, n_slots, window, sink, total, 0);
            ms[1] = run<1>(table, n_slots, window, sink, total, 0);
            ms[2] = run<2>(table, n_slots, window, sink, total, 0);
            ms[3] = run<3>(table, n_slots, window, sink, total, 0);
            ms[4] = run<4>(table, n_slots, window, sink, total, 0);
            for (int o = 0; o < 5; ++o)
                printf("table %3u MB  warp-window %3u slots  %-16s %8.3f ms  %8.1f Gop/s\n", mb, window, names[o], ms[o], total / ms[o] / 1e6);
        }
        cudaFree(table);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
