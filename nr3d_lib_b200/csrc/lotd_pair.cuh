// lotd_pair.cuh -- pieces of the "two lanes per point" layout shared by the fast LoTD kernels (lotd_fast.cu) and the fused
// encoder + decoder kernel (lotd_fused.cu).  See lotd_fast.cu for the measurements behind the layout.
#pragma once
#include "lotd_device.cuh"

namespace nr3d {

struct FastIn {
    uint64_t N;
    const float4* xs;        // sorted records (x, y, z, original index as bits) [N]
    const void* params;      // fp32 or fp16 table
    int32_t max_level;
    uint32_t base_aligned16;  // params pointer is 16-byte aligned
};

struct Geo2 {
    uint32_t key;   // cell key (10 bits per axis) for run detection
    float w[4];     // n-linear weights of this lane's four corners
    uint32_t e[4];  // element offsets (floats, from the start of the parameter array) of their feature pairs
};

__device__ __forceinline__ void pair_geo(const LevelDesc& L, uint32_t gfo, bool smooth, float x, float y, float z, uint32_t side, Geo2& g) {
    const uint32_t Ry = L.res[1], Rz = L.res[2];
    float p[3];
    uint32_t c[3];
    const float xv[3] = {x, y, z};
    const uint32_t R[3] = {L.res[0], Ry, Rz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float sc = (float)(R[d] - 2u);
        float v = xv[d] * sc + 0.5f;
        const float fl = floorf(v);
        c[d] = (uint32_t)fl;
        v -= fl;  // == (float)c[d] for the valid range x >= 0
        p[d] = smooth ? v * v * (3.0f - 2.0f * v) : v;
    }
    g.key = c[0] | (c[1] << 10) | (c[2] << 20);
    const float wx[2] = {1.0f - p[0], p[0]}, wy[2] = {1.0f - p[1], p[1]}, wz[2] = {1.0f - p[2], p[2]};
    const uint32_t nf = L.n_feat;
    const uint32_t base = L.offset + gfo;
    if (L.type == NR3D_LOD_DENSE) {
        const float wzs = side ? wz[1] : wz[0];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t dx = q & 1, dy = q >> 1;
            const uint32_t cell = ((c[0] + dx) * Ry + (c[1] + dy)) * Rz + c[2] + side;  // uint32 arithmetic as in the reference
            g.e[q] = base + cell * nf;
            g.w[q] = (wx[dx] * wy[dy]) * wzs;
        }
    } else {  // Hash
        const uint32_t size = L.size;
        const bool pow2 = (size & (size - 1u)) == 0;
        const uint32_t hx = c[0] + side;
        const float wxs = side ? wx[1] : wx[0];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t dy = q & 1, dz = q >> 1;
            const uint32_t hyz = ((c[1] + dy) * 2654435761u) ^ ((c[2] + dz) * 805459861u);
            const uint32_t h = pow2 ? ((hx ^ hyz) & (size - 1u)) : ((hx ^ hyz) % size);
            g.e[q] = base + h * nf;
            g.w[q] = (wxs * wy[dy]) * wz[dz];
        }
    }
}

}  // namespace nr3d
