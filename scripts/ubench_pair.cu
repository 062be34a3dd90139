// ubench_pair.cu -- what does one warp-wide gather / red instruction cost on B200 when GROUPS of lanes share a random
// 128-byte line or a random 32-byte sector?  Decides between the "cost = distinct lines" and "cost = distinct sectors" models
// for the LoTD kernels (x-neighbour corners of a hash level share a sector 75 % and a line 94 % of the time).
//   group g lanes share one random base; lane j of the group addresses  base + j * stride_slots   (slot = 8 bytes)
//   stride 1: neighbours inside a sector (g <= 4) / line;  stride 4: same line, different sectors (g <= 4)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pair ubench_pair.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int OP>  // 0 gather 8B, 1 red.v2, 2 gather 32B (v8.f32 per lane, group ignored), 3 gather 16B
__global__ void __launch_bounds__(256) k_pair(float* table, uint32_t n_lines, uint32_t per_thread, uint32_t g, uint32_t stride, float* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t grp = tid / g, j = tid % g;
    float acc = 0.f;
#pragma unroll 8
    for (uint32_t k = 0; k < per_thread; ++k) {
        const uint32_t line = hash32(grp * 2654435761u + k * 40503u) % n_lines;
        const uint32_t slot = line * 16u + ((j * stride) & 15u);
        if (OP == 0) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(table) + slot);
            acc += v.x + v.y;
        } else if (OP == 1) {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2ull * slot), "f"(1.f), "f"(2.f) : "memory");
        } else if (OP == 2) {
            float v0, v1, v2, v3, v4, v5, v6, v7;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3), "=f"(v4), "=f"(v5), "=f"(v6), "=f"(v7) : "l"(table + 2ull * (slot & ~3u)));
            acc += v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4*>(table) + (slot >> 1));
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

template <int OP>
float run(float* table, uint32_t n_lines, uint32_t g, uint32_t stride, float* sink, uint64_t total_ops) {
    const uint32_t per_thread = 64;
    const uint32_t threads = (uint32_t)(total_ops / per_thread);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) k_pair<OP><<<threads / 256, 256>>>(table, n_lines, per_thread, g, stride, sink);
    cudaEventRecord(a);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) k_pair<OP><<<threads / 256, 256>>>(table, n_lines, per_thread, g, stride, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const uint32_t mb = 48;
    const uint32_t n_lines = mb * 1024u * 1024u / 128u;
    float *table, *sink;
    cudaMalloc(&table, (size_t)mb << 20);
    cudaMemset(table, 0, (size_t)mb << 20);
    cudaMalloc(&sink, 4);
    const uint64_t total = 1ull << 28;
    const char* names[4] = {"gather 8B", "red.v2.f32", "gather 32B(v8)", "gather 16B"};
    const uint32_t cfg[][2] = {{1, 1}, {2, 1}, {2, 4}, {4, 1}, {4, 4}, {8, 1}, {8, 2}, {16, 1}, {32, 1}};
    for (auto& c : cfg) {
        float ms[4];
        ms[0] = run<0>(table, n_lines, c[0], c[1], sink, total);
        ms[1] = run<1>(table, n_lines, c[0], c[1], sink, total);
        ms[3] = run<3>(table, n_lines, c[0], c[1], sink, total);
        for (int o : {0, 1, 3})
            printf("group %2u lanes  stride %u slots  %-16s %8.3f ms  %8.1f G lanes/s\n", c[0], c[1], names[o], ms[o], total / ms[o] / 1e6);
    }
    float m = run<2>(table, n_lines, 1, 1, sink, total);
    printf("group  1 lanes  stride 1 slots  %-16s %8.3f ms  %8.1f G lanes/s\n", names[2], m, total / m / 1e6);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
