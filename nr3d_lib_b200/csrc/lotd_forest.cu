// lotd_forest.cu -- LoTD over a forest of blocks: every block owns one copy of the level tables, a point
// carries the index of the block it lies in and block-local coordinates in [0,1]^3, and lattice corners on a block face are
// fetched from / scattered into the NEIGHBOUR block so that the field is continuous across blocks.
//
// Behavioural contract: csrc/lotd/include/lotd/lotd_forest.h
//   forward (+dy/dx)        kernel_lod_forest                              :161-338, forest_fwd_n_linear :31-139
//   dL/dparam               kernel_lod_forest_backward_grid                :417-560, forest_bwd_n_linear :340-407
//   d(dL/dx)/dparam         kernel_lod_forest_backward_input_backward_grid :639-800
//   d(dL/dx)/dx             kernel_lod_forest_backward_input_backward_input:933-1066 (Dense / VM / Hash only)
//   block lookup            csrc/forest/forest.h:25-57 (kaolin-style octree walk), :88-95
// Differences from the single-block encoder: scale = res (not res - 2), so cells run 0..res and corner coordinate 0 means "left
// neighbour's res-1", res+1 "right neighbour's 0", 1..res the block's own 0..res-1; a corner whose block is not in the forest
// (or any remapped corner when continuity is disabled) contributes nothing.  Level types: Dense, VM, NPlaneMul, CP, Hash; D = 3.
// Outputs are row-major ([N, n_enc], [N, n_enc, 3]) like the reference's forest kernels.
// One thread per (point, pseudo level); the corner maths is shared with the generic kernels (lotd_device.cuh).
#include "lotd_kernels.cuh"

namespace nr3d {

struct ForestRef {
    const uint8_t* octree;
    const int32_t* exsum;
    const int16_t* block_ks;  // [n_trees, 3]
    uint32_t level, level_poffset, n_trees, continuity;
};

// == identify (forest.h:25-57): index of block k at `level` in the SPC point hierarchy, -1 if absent
__device__ __forceinline__ int32_t forest_identify(const ForestRef& fr, int kx, int ky, int kz) {
    const int maxval = (1 << fr.level) - 1;
    if (kx < 0 || ky < 0 || kz < 0 || kx > maxval || ky > maxval || kz > maxval) return -1;
    int ord = 0;
    for (uint32_t l = 0; l < fr.level; ++l) {
        const uint32_t depth = fr.level - l - 1;
        const uint32_t mask = 1u << depth;
        const uint32_t child = (((mask & (uint32_t)kx) << 2) | ((mask & (uint32_t)ky) << 1) | (mask & (uint32_t)kz)) >> depth;
        const uint32_t bits = fr.octree[ord];
        if (!(bits & (1u << child))) return -1;
        ord = fr.exsum[ord] + __popc(bits & ((2u << child) - 1u));
    }
    return ord - (int32_t)fr.level_poffset;
}

struct ForestCtx {
    Ctx<3> c;
    int bk[3];            // integer coordinates of the point's block
    uint64_t block_off;   // element offset of the point's block inside `params`
    uint32_t level_off;   // element offset of the level inside a block
};

template <int F>
__device__ __forceinline__ bool forest_setup(const LotdTable& tab, const LotdIn& in, const ForestRef& fr, uint64_t i, uint32_t pl, ForestCtx& fc) {
    Ctx<3>& c = fc.c;
    const uint32_t level = tab.map_level[pl];
    if ((int32_t)level > in.max_level) return false;
    uint32_t block_ind = 0;
    if (in.batch_inds) {
        const int64_t b = in.batch_inds[i];
        if (b < 0) return false;
        block_ind = (uint32_t)b;
    } else if (in.batch_data_size) {
        block_ind = (uint32_t)(i / in.batch_data_size);
    }
    fc.block_off = in.batch_offsets ? (uint64_t)in.batch_offsets[block_ind] : (uint64_t)block_ind * tab.n_params;
    const int16_t* k = fr.block_ks + (uint64_t)block_ind * 3;
    fc.bk[0] = k[0]; fc.bk[1] = k[1]; fc.bk[2] = k[2];
    const LevelDesc& L = tab.lv[level];
    fc.level_off = L.offset;
    c.base = fc.block_off + L.offset;
    c.type = L.type;
    c.n_feat = L.n_feat;
    c.size = L.size;
    c.gfo = (uint32_t)tab.map_cnt[pl] * F;
    const float* xp = in.x + i * 3;
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        c.res[d] = L.res[d];
        c.scale[d] = (float)c.res[d];  // NOTE: for forest (lotd_forest.h:239)
        float v = xp[d] * c.scale[d] + 0.5f;
        const float fl = floorf(v);
        c.cell[d] = (uint32_t)fl;
        v -= (float)c.cell[d];
        if (smooth) {
            c.p[d] = v * v * (3.0f - 2.0f * v);
            c.dp[d] = 6.0f * v * (1.0f - v);
            c.d2p[d] = 6.0f - 12.0f * v;
        } else {
            c.p[d] = v;
            c.dp[d] = 1.0f;
            c.d2p[d] = 0.0f;
        }
    }
    return true;
}

// Continuity rule (lotd_forest.h:53-88): corner coordinate -> (owning block, block-local coordinate).  Returns false when the
// corner has no owner; *base receives the element offset of the owner's level table.
__device__ __forceinline__ bool forest_remap(const LotdTable& tab, const LotdIn& in, const ForestRef& fr, const ForestCtx& fc,
                                             const uint32_t* pos, uint32_t* lp, uint64_t* base) {
    int k[3];
    bool changed = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const uint32_t p = pos[d], R = fc.c.res[d];
        if (p == 0) { k[d] = fc.bk[d] - 1; lp[d] = R - 1; changed = true; }
        else if (p == R + 1) { k[d] = fc.bk[d] + 1; lp[d] = 0; changed = true; }
        else { k[d] = fc.bk[d]; lp[d] = p - 1; }
    }
    uint64_t off = fc.block_off;
    if (changed) {
        if (!fr.continuity) return false;
        const int32_t nb = forest_identify(fr, k[0], k[1], k[2]);
        if (nb < 0) return false;
        off = in.batch_offsets ? (uint64_t)in.batch_offsets[nb] : (uint64_t)nb * tab.n_params;
    }
    *base = off + fc.level_off;
    return true;
}

__device__ __forceinline__ bool forest_type_ok(uint32_t t) {
    return t == NR3D_LOD_DENSE || t == NR3D_LOD_VM || t == NR3D_LOD_NPLANEMUL || t == NR3D_LOD_CP || t == NR3D_LOD_HASH;
}

template <int F, typename PT, bool DYDX>
__global__ void __launch_bounds__(kLotdThreads)
forest_fwd_kernel(const __grid_constant__ LotdTable tab, const LotdIn in, const ForestRef fr, PT* __restrict__ y, float* __restrict__ dydx) {
    using C = Cvt<PT>;
    constexpr int D = 3;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.N) return;
    const uint32_t pl = blockIdx.y;
    PT r[F];
    float gr[F][D];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        r[f] = C::zero();
#pragma unroll
        for (int d = 0; d < D; ++d) gr[f][d] = 0.f;
    }
    ForestCtx fc;
    if (forest_setup<F>(tab, in, fr, i, pl, fc) && forest_type_ok(fc.c.type)) {
        const Ctx<D>& c = fc.c;
        const PT* params = reinterpret_cast<const PT*>(in.params);
        const bool vec_ok = in.vec_ok;
        PT v[1 << D][F];
#pragma unroll
        for (int idx = 0; idx < (1 << D); ++idx) {
            uint32_t pos[D], lp[D];
#pragma unroll
            for (int d = 0; d < D; ++d) pos[d] = c.cell[d] + ((idx >> d) & 1);
            uint64_t base;
            if (forest_remap(tab, in, fr, fc, pos, lp, &base)) corner_val<D, F, PT>(c, params + base, lp, v[idx], vec_ok);
            else {
#pragma unroll
                for (int f = 0; f < F; ++f) v[idx][f] = C::zero();
            }
        }
#pragma unroll
        for (int idx = 0; idx < (1 << D); ++idx) {
            uint32_t pos[D];
            const float w = corner_weight<D>(c, idx, pos);
#pragma unroll
            for (int f = 0; f < F; ++f) r[f] = C::add(r[f], C::from_f(w * C::to_f(v[idx][f])));
        }
        if (DYDX) {
#pragma unroll
            for (int gd = 0; gd < D; ++gd) {
#pragma unroll
                for (int idx = 0; idx < (1 << (D - 1)); ++idx) {
                    uint32_t pos[D];
                    int li;
                    const float w = face_weight<D>(c, gd, idx, c.scale[gd] * c.dp[gd], pos, &li);
                    const int ri = li + (1 << gd);
#pragma unroll
                    for (int f = 0; f < F; ++f) gr[f][gd] += w * (C::to_f(v[ri][f]) - C::to_f(v[li][f]));
                }
            }
        }
    }
    PT* yo = y + i * tab.n_enc + pl * F;
#pragma unroll
    for (int f = 0; f < F; ++f) st_cs(yo + f, r[f]);
    if (DYDX) {
        float* go = dydx + (i * tab.n_enc + pl * F) * D;
#pragma unroll
        for (int f = 0; f < F; ++f)
#pragma unroll
            for (int d = 0; d < D; ++d) st_cs(go + f * D + d, gr[f][d]);
    }
}

template <int F, typename PT, bool SECOND>
__global__ void __launch_bounds__(kLotdThreads)
forest_bwd_param_kernel(const __grid_constant__ LotdTable tab, const LotdIn in, const ForestRef fr, const PT* __restrict__ dLdy, int64_t gs_n,
                        int64_t gs_f, const float* __restrict__ ddx, PT* __restrict__ grad_params) {
    using C = Cvt<PT>;
    constexpr int D = 3;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.N) return;
    const uint32_t pl = blockIdx.y;
    ForestCtx fc;
    if (!forest_setup<F>(tab, in, fr, i, pl, fc) || !forest_type_ok(fc.c.type)) return;
    const Ctx<D>& c = fc.c;
    const PT* params = reinterpret_cast<const PT*>(in.params);
    const bool vec_ok = in.vec_ok;
    float grad[F];
    {
        const PT* gp = dLdy + (int64_t)i * gs_n + (int64_t)(pl * F) * gs_f;
#pragma unroll
        for (int f = 0; f < F; ++f) grad[f] = C::to_f(gp[(int64_t)f * gs_f]);
    }
    if (!SECOND) {
#pragma unroll 1
        for (int idx = 0; idx < (1 << D); ++idx) {
            uint32_t pos[D], lp[D];
            const float w = corner_weight<D>(c, idx, pos);
            uint64_t base;
            if (forest_remap(tab, in, fr, fc, pos, lp, &base)) corner_add_grad<D, F, PT>(c, params + base, grad_params + base, lp, grad, w, vec_ok);
        }
    } else {
        float gin[D];
#pragma unroll
        for (int d = 0; d < D; ++d) gin[d] = c.scale[d] * ddx[i * D + d] * c.dp[d];
#pragma unroll 1
        for (int gd = 0; gd < D; ++gd) {
#pragma unroll 1
            for (int idx = 0; idx < (1 << (D - 1)); ++idx) {
                uint32_t pos[D], lp[D];
                int li;
                const float w = face_weight<D>(c, gd, idx, gin[gd], pos, &li);
                uint64_t base;
                pos[gd] = c.cell[gd];
                if (forest_remap(tab, in, fr, fc, pos, lp, &base)) corner_add_grad<D, F, PT>(c, params + base, grad_params + base, lp, grad, -w, vec_ok);
                pos[gd] = c.cell[gd] + 1;
                if (forest_remap(tab, in, fr, fc, pos, lp, &base)) corner_add_grad<D, F, PT>(c, params + base, grad_params + base, lp, grad, w, vec_ok);
            }
        }
    }
}

template <int F, typename PT>
__global__ void __launch_bounds__(kLotdThreads)
forest_bwdbwd_input_kernel(const __grid_constant__ LotdTable tab, const LotdIn in, const ForestRef fr, const PT* __restrict__ dLdy, int64_t gs_n,
                           int64_t gs_f, const float* __restrict__ ddx, float* __restrict__ dLdx) {
    using C = Cvt<PT>;
    constexpr int D = 3;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.N) return;
    const uint32_t pl = blockIdx.y;
    ForestCtx fc;
    if (!forest_setup<F>(tab, in, fr, i, pl, fc)) return;
    const Ctx<D>& c = fc.c;
    if (!(c.type == NR3D_LOD_DENSE || c.type == NR3D_LOD_HASH || c.type == NR3D_LOD_VM)) return;   // lotd_forest.h:1022-1050
    const PT* params = reinterpret_cast<const PT*>(in.params);
    const bool vec_ok = in.vec_ok;
    float grad[F];
    {
        const PT* gp = dLdy + (int64_t)i * gs_n + (int64_t)(pl * F) * gs_f;
#pragma unroll
        for (int f = 0; f < F; ++f) grad[f] = C::to_f(gp[(int64_t)f * gs_f]);
    }
    float S[1 << D];
#pragma unroll
    for (int idx = 0; idx < (1 << D); ++idx) {
        uint32_t pos[D], lp[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pos[d] = c.cell[d] + ((idx >> d) & 1);
        uint64_t base;
        S[idx] = forest_remap(tab, in, fr, fc, pos, lp, &base) ? corner_dot<D, F, PT>(c, params + base, lp, grad, vec_ok) : 0.f;
    }
    float gin_other[D], gin_diag[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float gi = ddx[i * D + d];
        gin_other[d] = c.scale[d] * gi * c.dp[d];
        gin_diag[d] = (c.scale[d] * gi) * (c.scale[d] * c.d2p[d]);
    }
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
#pragma unroll
    for (int gd = 0; gd < D; ++gd) {
        float out = 0.f;
#pragma unroll
        for (int idx = 0; idx < (1 << (D - 1)); ++idx) {
            if (smooth) {
                uint32_t pos[D];
                int li;
                const float w = face_weight<D>(c, gd, idx, gin_diag[gd], pos, &li);
                out += S[li] * (-w);
                out += S[li + (1 << gd)] * w;
            }
#pragma unroll
            for (int og = 0; og < D - 1; ++og) {
                const int o = og >= gd ? og + 1 : og;
                float w = gin_other[o] * (c.dp[gd] * c.scale[gd]);
                int li = 0;
#pragma unroll
                for (int ng = 0; ng < D - 1; ++ng) {
                    const int dim = ng >= o ? ng + 1 : ng;
                    if ((idx & (1 << ng)) == 0) {
                        if (dim != gd) w *= 1.0f - c.p[dim];
                        else w *= -1.0f;
                    } else {
                        if (dim != gd) w *= c.p[dim];
                        li += 1 << dim;
                    }
                }
                out += S[li] * (-w);
                out += S[li + (1 << o)] * w;
            }
        }
        atomicAdd(dLdx + i * D + gd, out);
    }
}

static int forest_prepare(const nr3d_lotd_meta* meta, const nr3d_forest_meta* forest, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                          const void* x, const void* params, const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size,
                          int32_t max_level, void* stream, LotdLaunch& L, ForestRef& fr) {
    NR3D_CHECK(forest != nullptr, "LoTDEncoding: null forest meta");
    NR3D_CHECK(input_dtype == NR3D_F32, "LoTDEncoding: lotd-forest needs fp32 points (<input,param> -> (float, half), (float, float))");
    NR3D_CHECK(meta != nullptr && meta->n_dims_to_encode == 3, "LoTDEncoding::fwd: lotd-forest only supports `n_dims_to_encode`==3");
    NR3D_CHECK(forest->octree && forest->exsum && forest->block_ks, "LoTDEncoding: forest.octree / forest.exsum / forest.block_ks must be given");
    NR3D_CHECK(forest->level <= 15, "LoTDEncoding: forest level must be <= 15 (block coordinates are int16)");
    for (uint32_t l = 0; l < meta->n_levels; ++l) {
        const uint32_t t = meta->level_types[l];
        NR3D_CHECK(t == NR3D_LOD_DENSE || t == NR3D_LOD_VM || t == NR3D_LOD_NPLANEMUL || t == NR3D_LOD_CP || t == NR3D_LOD_HASH,
                   "LoTDEncoding: lotd-forest supports Dense / VM / NPlaneMul / CP / Hash levels only");
    }
    if (int rc = build_launch(meta, input_dtype, param_dtype, N, x, params, batch_inds, batch_offsets, batch_data_size, max_level, stream, L)) return rc;
    fr.octree = forest->octree; fr.exsum = forest->exsum; fr.block_ks = forest->block_ks;
    fr.level = forest->level; fr.level_poffset = forest->level_poffset; fr.n_trees = forest->n_trees;
    fr.continuity = forest->continuity_enabled ? 1u : 0u;
    return 0;
}

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_lotd_forest_fwd(const nr3d_lotd_meta* meta, const nr3d_forest_meta* forest, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                         const void* x, const void* params, const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size,
                         int32_t max_level, void* y, void* dy_dx, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    ForestRef fr;
    if (int rc = forest_prepare(meta, forest, input_dtype, param_dtype, N, x, params, batch_inds, batch_offsets, batch_data_size, max_level, stream, L, fr)) return rc;
    NR3D_CHECK(x && params && y, "LoTDEncoding::fwd: null argument");
    const dim3 grid((unsigned)div_up<uint64_t>(N, kLotdThreads), L.tab.n_pseudo, 1);
    NR3D_LOTD_DISPATCH_F_PT(
        if (dy_dx) forest_fwd_kernel<F, PT, true><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, fr, (PT*)y, (float*)dy_dx);
        else forest_fwd_kernel<F, PT, false><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, fr, (PT*)y, nullptr))
    NR3D_LAUNCH_CHECK("lotd_forest_fwd");
    return 0;
}

int nr3d_lotd_forest_bwd_param(const nr3d_lotd_meta* meta, const nr3d_forest_meta* forest, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                               const void* dL_dy, int64_t s_n, int64_t s_f, const void* dL_ddLdx, const void* x, const void* params,
                               const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size, int32_t max_level,
                               void* dL_dparam, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    ForestRef fr;
    if (int rc = forest_prepare(meta, forest, input_dtype, param_dtype, N, x, params, batch_inds, batch_offsets, batch_data_size, max_level, stream, L, fr)) return rc;
    NR3D_CHECK(x && params && dL_dy && dL_dparam, "LoTDEncoding::bwd: null argument");
    L.in.vec_ok = L.in.vec_ok && ((reinterpret_cast<uintptr_t>(dL_dparam) & 7u) == 0);
    const dim3 grid((unsigned)div_up<uint64_t>(N, kLotdThreads), L.tab.n_pseudo, 1);
    const float* ddx = (const float*)dL_ddLdx;
    NR3D_LOTD_DISPATCH_F_PT(
        if (ddx) forest_bwd_param_kernel<F, PT, true><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, fr, (const PT*)dL_dy, s_n, s_f, ddx, (PT*)dL_dparam);
        else forest_bwd_param_kernel<F, PT, false><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, fr, (const PT*)dL_dy, s_n, s_f, nullptr, (PT*)dL_dparam))
    NR3D_LAUNCH_CHECK("lotd_forest_bwd_param");
    return 0;
}

int nr3d_lotd_forest_bwd_bwd_dx(const nr3d_lotd_meta* meta, const nr3d_forest_meta* forest, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                                const void* dL_ddLdx, const void* dL_dy, int64_t s_n, int64_t s_f, const void* x, const void* params,
                                const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size, int32_t max_level,
                                void* dL_dx, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    ForestRef fr;
    if (int rc = forest_prepare(meta, forest, input_dtype, param_dtype, N, x, params, batch_inds, batch_offsets, batch_data_size, max_level, stream, L, fr)) return rc;
    NR3D_CHECK(x && params && dL_dy && dL_ddLdx && dL_dx, "LoTDEncoding::bwd_bwd_input: null argument");
    const dim3 grid((unsigned)div_up<uint64_t>(N, kLotdThreads), L.tab.n_pseudo, 1);
    NR3D_LOTD_DISPATCH_F_PT((forest_bwdbwd_input_kernel<F, PT><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, fr, (const PT*)dL_dy, s_n, s_f, (const float*)dL_ddLdx, (float*)dL_dx)))
    NR3D_LAUNCH_CHECK("lotd_forest_bwdbwd_input");
    return 0;
}

}  // extern "C"
