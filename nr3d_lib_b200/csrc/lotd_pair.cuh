// lotd_pair.cuh -- pieces of the "two lanes per point" layout shared by the fast LoTD kernels (lotd_fast.cu), the point sort
// (lotd_sort.cu) and the fused encoder + decoder kernel (lotd_fused.cu).  See lotd_fast.cu for the measurements behind the layout.
#pragma once
#include "lotd_device.cuh"
#include <type_traits>

namespace nr3d {

#ifndef NR3D_BIN_ORDER      // order of the sort bins: 0 x fastest, 1 z fastest, 2 bricks of 8 x 4 x 2 bins (x fastest inside and between bricks)
#define NR3D_BIN_ORDER 0    // A/B on B200 (profiles/r2_ab_tiles.txt): x fastest 2.046 ms / step, bricks 2.085 ms (bricks only pay with CTA tiles)
#endif

// Position of bin (bx, by, bz) in the sorted order.  Bricks keep the 16 points of a warp on one x pencil (what the Hash levels like: the
// x neighbours of a cell share a 128-byte line) and make the 128 points of a CTA a compact 8 x 4 x 2 block of bins, so that the corner
// bounding box of a CTA is small on the coarse and middle levels (shared-memory tiles of the backward, lotd_fast.cu).  `res` is a multiple of 8.
__host__ __device__ __forceinline__ uint32_t bin_order(uint32_t bx, uint32_t by, uint32_t bz, uint32_t res) {
#if NR3D_BIN_ORDER == 0
    return (bz * res + by) * res + bx;
#elif NR3D_BIN_ORDER == 1
    return (bx * res + by) * res + bz;
#else
    const uint32_t nbx = res >> 3, nby = res >> 2;
    const uint32_t brick = ((bz >> 1) * nby + (by >> 2)) * nbx + (bx >> 3);
    return (brick << 6) | ((bz & 1u) << 5) | ((by & 3u) << 3) | (bx & 7u);
#endif
}

// Per-level constants derived on the host once per call, so that the kernels do not re-derive them per point and level (the fast kernels
// are issue-slot bound, profiles/r2_pair_ncu_summary.txt): float(res - 2) per axis and the hash mask.
struct FastLevel {
    float scale[3];   // (float)(res[d] - 2): the factor of cell_pos()
    uint32_t hmask;   // Hash levels with a power-of-two table (> 1 entry): size - 1; otherwise 0 (-> modulo)
};

struct FastIn {
    uint64_t N;
    const float4* xs;         // sorted records (x, y, z, original index as bits) [N]
    const uint16_t* scenes;   // scene of every sorted record (batched calls; 0xffff = skipped point) or NULL for a single scene
    const void* params;       // fp32 or fp16 tables, n_scenes * n_params elements
    int32_t max_level;
    uint32_t n_params;        // elements per scene
    uint32_t pl_begin, pl_end;  // pseudo levels [pl_begin, pl_end) to process (backward only: level groups whose all-reduce starts early)
    uint32_t merge_res;       // backward only: levels up to this resolution merge same-cell points inside a warp before the scatter (0: none)
    FastLevel fl[NR3D_MAX_LEVELS];   // filled by fast_levels()
};

inline void fast_levels(const nr3d_lotd_meta* m, FastIn& in) {
    for (uint32_t l = 0; l < NR3D_MAX_LEVELS; ++l) {
        FastLevel& f = in.fl[l];
        f.scale[0] = f.scale[1] = f.scale[2] = 0.f; f.hmask = 0;
        if (l >= m->n_levels) continue;
        for (int d = 0; d < 3; ++d) f.scale[d] = (float)(m->level_res[l][d] - 2u);
        const uint32_t size = m->level_sizes[l];
        if (size > 1u && (size & (size - 1u)) == 0u) f.hmask = size - 1u;
    }
}

struct Geo2 {
    uint32_t key;   // cell key (10 bits per axis) for run detection
    uint32_t c[3];  // cell of the point
    float w[4];     // n-linear weights of this lane's four corners
    uint32_t e[4];  // element offsets (from the start of the scene's parameter array) of their feature groups
};

// cell of coordinate v on a level with `res` cells along the axis: floor(v * (res - 2) + 0.5) with ONE rounding (fma), the reference's
// pos_fract (lotd_cuda.h:959-1077 compiled with -fmad=true).  Monotone in v, which the CTA bounding boxes of the backward rely on.
__device__ __forceinline__ float cell_pos(float v, uint32_t res) { return __fmaf_rn(v, (float)(res - 2u), 0.5f); }

// gfo: feature offset of the pseudo level inside the level's entries (map_cnt * F); pbase: first element of the point's scene (the element
// offsets g.e[] include it: 32-bit arithmetic throughout, check_fast guarantees n_scenes * n_params < 2^32)
__device__ __forceinline__ void pair_geo(const LevelDesc& L, const FastLevel& X, uint32_t gfo, uint32_t pbase, bool smooth, float x, float y, float z,
                                         uint32_t side, Geo2& g) {
    const uint32_t Ry = L.res[1], Rz = L.res[2];
    float p[3];
    const float xv[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float v = __fmaf_rn(xv[d], X.scale[d], 0.5f);   // == cell_pos(xv[d], L.res[d])
        const float fl = floorf(v);
        g.c[d] = (uint32_t)fl;
        v -= fl;  // == (float)c[d] for the valid range x >= 0
        p[d] = v;
    }
    if (smooth) {   // (uniform; kept out of the common path: the fast kernels are issue bound)
#pragma unroll
        for (int d = 0; d < 3; ++d) p[d] = p[d] * p[d] * (3.0f - 2.0f * p[d]);
    }
    g.key = g.c[0] | (g.c[1] << 10) | (g.c[2] << 20);
    const float wx[2] = {1.0f - p[0], p[0]}, wy[2] = {1.0f - p[1], p[1]}, wz[2] = {1.0f - p[2], p[2]};
    const uint32_t nf = L.n_feat;
    const uint32_t base = L.offset + gfo + pbase;
    if (L.type == NR3D_LOD_DENSE) {
        const float wzs = side ? wz[1] : wz[0];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t dx = q & 1, dy = q >> 1;
            const uint32_t cell = ((g.c[0] + dx) * Ry + (g.c[1] + dy)) * Rz + g.c[2] + side;  // uint32 arithmetic as in the reference
            g.e[q] = base + cell * nf;
            g.w[q] = (wx[dx] * wy[dy]) * wzs;
        }
    } else {  // Hash: (x + s) ^ (y + dy) * P1 ^ (z + dz) * P2, the products of the +1 corners by one addition (uint32 arithmetic wraps like the product)
        const uint32_t hx = g.c[0] + side;
        const float wxs = side ? wx[1] : wx[0];
        const uint32_t y0 = g.c[1] * 2654435761u, y1 = y0 + 2654435761u, z0 = g.c[2] * 805459861u, z1 = z0 + 805459861u;
        const uint32_t h4[4] = {hx ^ y0 ^ z0, hx ^ y1 ^ z0, hx ^ y0 ^ z1, hx ^ y1 ^ z1};   // q = dy + 2 dz
        if (X.hmask) {   // power-of-two table (uniform per level)
#pragma unroll
            for (int q = 0; q < 4; ++q) g.e[q] = base + (h4[q] & X.hmask) * nf;
        } else {
            const uint32_t size = L.size;
#pragma unroll
            for (int q = 0; q < 4; ++q) g.e[q] = base + (h4[q] % size) * nf;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) g.w[q] = (wxs * wy[q & 1]) * wz[q >> 1];
    }
}

// pair_geo plus the derivative weights of the lane's four corners:
//   dw[d][q] = scale_d * phi'_d * (+1 if corner q is the right neighbour along d else -1) * prod_{d' != d} w_{d'}(q)
// so that  dy/dx_d = sum_q dw[d][q] * value(q)  (summed over both lanes of the pair; reference linear_interpolate.cuh:122-150) and the
// second-order scatter weight of corner q is  sum_d dL_ddLdx[d] * dw[d][q]  (reference lotd_hash_only.h:472-695 walks the faces instead).
__device__ __forceinline__ void pair_geo_d(const LevelDesc& L, const FastLevel& X, uint32_t gfo, uint32_t pbase, bool smooth, float x, float y, float z,
                                           uint32_t side, Geo2& g, float (&dw)[3][4]) {
    const uint32_t Ry = L.res[1], Rz = L.res[2];
    float p[3], sd[3];   // sd = scale * phi'
    const float xv[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float sc = X.scale[d];
        float v = __fmaf_rn(xv[d], sc, 0.5f);
        const float fl = floorf(v);
        g.c[d] = (uint32_t)fl;
        v -= fl;
        p[d] = smooth ? v * v * (3.0f - 2.0f * v) : v;
        sd[d] = smooth ? sc * (6.0f * v * (1.0f - v)) : sc;
    }
    g.key = g.c[0] | (g.c[1] << 10) | (g.c[2] << 20);
    const float w3[3][2] = {{1.0f - p[0], p[0]}, {1.0f - p[1], p[1]}, {1.0f - p[2], p[2]}};
    const uint32_t nf = L.n_feat;
    const uint32_t base = L.offset + gfo + pbase;
    const bool dense = L.type == NR3D_LOD_DENSE;
    const uint32_t size = L.size;
    const uint32_t y0 = g.c[1] * 2654435761u, y1 = y0 + 2654435761u, z0 = g.c[2] * 805459861u, z1 = z0 + 805459861u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        // bits of this corner along (x, y, z): the lane's side bit sits on z for Dense levels and on x for Hash levels
        const uint32_t b[3] = {dense ? (uint32_t)(q & 1) : side, dense ? (uint32_t)(q >> 1) : (uint32_t)(q & 1), dense ? side : (uint32_t)(q >> 1)};
        float wsel[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) wsel[d] = b[d] ? w3[d][1] : w3[d][0];
        if (dense) {
            const uint32_t cell = ((g.c[0] + b[0]) * Ry + (g.c[1] + b[1])) * Rz + g.c[2] + b[2];
            g.e[q] = base + cell * nf;
        } else {
            const uint32_t hh = (g.c[0] + side) ^ ((q & 1) ? y1 : y0) ^ ((q >> 1) ? z1 : z0);
            g.e[q] = base + (X.hmask ? (hh & X.hmask) : (hh % size)) * nf;
        }
        g.w[q] = (wsel[0] * wsel[1]) * wsel[2];
        dw[0][q] = (b[0] ? sd[0] : -sd[0]) * (wsel[1] * wsel[2]);
        dw[1][q] = (b[1] ? sd[1] : -sd[1]) * (wsel[0] * wsel[2]);
        dw[2][q] = (b[2] ? sd[2] : -sd[2]) * (wsel[0] * wsel[1]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// F features of one table entry: fp32 tables accumulate in fp32; fp16 tables accumulate every term in half like the reference
// (linear_interpolate.cuh:118) and scatter with packed-half reductions.  The access width is free on B200 (one LSU slot per distinct
// 128-byte line, one L2 reduction packet per distinct 16-byte chunk), so wider pseudo levels cost the same number of memory operations.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ __half2 as_half2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2_bits(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }

template <typename PT, int F>
__device__ __forceinline__ void load_feats(const PT* p, float (&v)[F]) {
    if constexpr (std::is_same<PT, float>::value) {
        if constexpr (F == 2) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(p));
            v[0] = t.x; v[1] = t.y;
        } else {
#pragma unroll
            for (int j = 0; j < F / 4; ++j) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
                v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            }
        }
    } else {
        uint32_t u[F / 2];
        if constexpr (F == 2) {
            u[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
        } else if constexpr (F == 4) {
            const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
            u[0] = t.x; u[1] = t.y;
        } else {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
            u[0] = t.x; u[1] = t.y; u[2] = t.z; u[3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < F / 2; ++j) {
            const float2 t = __half22float2(as_half2(u[j]));
            v[2 * j] = t.x; v[2 * j + 1] = t.y;
        }
    }
}

// one reduction instruction per 16 bytes (fp32: v2 / v4; fp16: f16x2 / v2.f16x2 / v4.f16x2)
template <typename PT, int F>
__device__ __forceinline__ void red_feats(PT* p, const float (&v)[F]) {
    if constexpr (std::is_same<PT, float>::value) {
        if constexpr (F == 2) {
            red_add_v2_f32(p, v[0], v[1]);
        } else {
#pragma unroll
            for (int j = 0; j < F / 4; ++j) red_add_v4_f32(p + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    } else {
        if constexpr (F == 2) {
            red_add_h2(p, __floats2half2_rn(v[0], v[1]));
        } else if constexpr (F == 4) {
            asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(p), "r"(h2_bits(v[0], v[1])), "r"(h2_bits(v[2], v[3])) : "memory");
        } else {
            asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(h2_bits(v[0], v[1])), "r"(h2_bits(v[2], v[3])),
                         "r"(h2_bits(v[4], v[5])), "r"(h2_bits(v[6], v[7])) : "memory");
        }
    }
}

template <typename PT> __device__ __forceinline__ float ldcs_f(const PT* p);
template <> __device__ __forceinline__ float ldcs_f<float>(const float* p) { return __ldcs(p); }
template <> __device__ __forceinline__ float ldcs_f<__half>(const __half* p) { return __half2float(__ldcs(p)); }

// predicated streaming load (0 when off): one instruction instead of a branch with reconvergence bookkeeping around it
template <typename PT> __device__ __forceinline__ float ldcs_f_if(bool on, const PT* p);
template <> __device__ __forceinline__ float ldcs_f_if<float>(bool on, const float* p) {
    float v = 0.f;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.cs.f32 %0, [%1];\n\t}" : "+f"(v) : "l"(p), "r"((int)on) : "memory");
    return v;
}
template <> __device__ __forceinline__ float ldcs_f_if<__half>(bool on, const __half* p) {
    unsigned short u = 0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.cs.b16 %0, [%1];\n\t}" : "+h"(u) : "l"(p), "r"((int)on) : "memory");
    return __half2float(*reinterpret_cast<__half*>(&u));
}

}  // namespace nr3d
