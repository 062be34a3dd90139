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

// pair_geo plus the derivative weights of the lane's four corners:
//   dw[d][q] = scale_d * phi'_d * (+1 if corner q is the right neighbour along d else -1) * prod_{d' != d} w_{d'}(q)
// so that  dy/dx_d = sum_q dw[d][q] * value(q)  (summed over both lanes of the pair; reference linear_interpolate.cuh:122-150) and the
// second-order scatter weight of corner q is  sum_d dL_ddLdx[d] * dw[d][q]  (reference lotd_hash_only.h:472-695 walks the faces instead).
__device__ __forceinline__ void pair_geo_d(const LevelDesc& L, uint32_t gfo, bool smooth, float x, float y, float z, uint32_t side, Geo2& g,
                                           float (&dw)[3][4]) {
    const uint32_t Ry = L.res[1], Rz = L.res[2];
    float p[3], sd[3];   // sd = scale * phi'
    uint32_t c[3];
    const float xv[3] = {x, y, z};
    const uint32_t R[3] = {L.res[0], Ry, Rz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float sc = (float)(R[d] - 2u);
        float v = xv[d] * sc + 0.5f;
        const float fl = floorf(v);
        c[d] = (uint32_t)fl;
        v -= fl;
        p[d] = smooth ? v * v * (3.0f - 2.0f * v) : v;
        sd[d] = smooth ? sc * (6.0f * v * (1.0f - v)) : sc;
    }
    g.key = c[0] | (c[1] << 10) | (c[2] << 20);
    const float w3[3][2] = {{1.0f - p[0], p[0]}, {1.0f - p[1], p[1]}, {1.0f - p[2], p[2]}};
    const uint32_t nf = L.n_feat;
    const uint32_t base = L.offset + gfo;
    const bool dense = L.type == NR3D_LOD_DENSE;
    const uint32_t size = L.size;
    const bool pow2 = (size & (size - 1u)) == 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        // bits of this corner along (x, y, z): the lane's side bit sits on z for Dense levels and on x for Hash levels
        const uint32_t b[3] = {dense ? (uint32_t)(q & 1) : side, dense ? (uint32_t)(q >> 1) : (uint32_t)(q & 1), dense ? side : (uint32_t)(q >> 1)};
        float wsel[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) wsel[d] = b[d] ? w3[d][1] : w3[d][0];
        if (dense) {
            const uint32_t cell = ((c[0] + b[0]) * Ry + (c[1] + b[1])) * Rz + c[2] + b[2];
            g.e[q] = base + cell * nf;
            g.w[q] = (wsel[0] * wsel[1]) * wsel[2];
        } else {
            const uint32_t hyz = ((c[1] + b[1]) * 2654435761u) ^ ((c[2] + b[2]) * 805459861u);
            const uint32_t hx = c[0] + b[0];
            const uint32_t h = pow2 ? ((hx ^ hyz) & (size - 1u)) : ((hx ^ hyz) % size);
            g.e[q] = base + h * nf;
            g.w[q] = (wsel[0] * wsel[1]) * wsel[2];
        }
        dw[0][q] = (b[0] ? sd[0] : -sd[0]) * (wsel[1] * wsel[2]);
        dw[1][q] = (b[1] ? sd[1] : -sd[1]) * (wsel[0] * wsel[2]);
        dw[2][q] = (b[2] ? sd[2] : -sd[2]) * (wsel[0] * wsel[1]);
    }
}

}  // namespace nr3d
