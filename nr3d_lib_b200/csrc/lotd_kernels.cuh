// lotd_kernels.cuh -- LoTD kernels (generic over level type), one thread per (point, pseudo level).
// Instantiated per input dimension in lotd_d{2,3,4}.cu.  The Dense/Hash tuned kernels live in lotd_fast.cu.
#pragma once
#include "lotd_device.cuh"
#include <string.h>

namespace nr3d {

constexpr int kLotdThreads = 256;

// =================================================================================================
// forward (+ optional dy/dx).  Reference: kernel_lod (lotd_encoding.h:113-428) and
// kernel_lod_hash_only[_with_dydx] (lotd_hash_only.h:15-378).  Every output element is written.
// =================================================================================================
template <int D, int F, typename PT, bool DYDX>
__global__ void __launch_bounds__(kLotdThreads)
lotd_fwd_kernel(const __grid_constant__ LotdTable tab, const LotdIn in, PT* __restrict__ y, int64_t ys_n, int64_t ys_f,
                float* __restrict__ dydx, int64_t ds_n, int64_t ds_f) {
    using C = Cvt<PT>;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.N) return;
    const uint32_t pl = blockIdx.y;
    const uint32_t ofo = pl * F;
    PT r[F];
    float gr[F][D];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        r[f] = C::zero();
#pragma unroll
        for (int d = 0; d < D; ++d) gr[f][d] = 0.f;
    }
    Ctx<D> c;
    if (lotd_setup<D, F>(tab, in, i, pl, c)) {
        const PT* g = reinterpret_cast<const PT*>(in.params) + c.base;
        const bool vec_ok = in.vec_ok;
        if (is_nlinear(c.type)) {
            PT v[1 << D][F];
            all_corner_vals<D, F, PT>(c, g, v, vec_ok);
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
        } else if (c.type == NR3D_LOD_NPLANESUM) {
            if constexpr (D > 2) {  // lotd_encoding.h:268-351
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    PT v[1 << (D - 1)][F];
                    float wv[1 << (D - 1)];
#pragma unroll
                    for (int idx = 0; idx < (1 << (D - 1)); ++idx) {
                        uint32_t pp[D - 1];
                        float w = 1.0f;
#pragma unroll
                        for (int d2 = 0; d2 < D - 1; ++d2) {
                            const int d3 = d2 >= j ? d2 + 1 : d2;
                            if ((idx & (1 << d2)) == 0) { w *= 1.0f - c.p[d3]; pp[d2] = c.cell[d3]; }
                            else { w *= c.p[d3]; pp[d2] = c.cell[d3] + 1; }
                        }
                        wv[idx] = w;
                        load_feats<F>(g + (uint64_t)idx_nplane_sub<D>(c.res, pp, j) * c.n_feat + c.gfo, v[idx], vec_ok);
                    }
#pragma unroll
                    for (int idx = 0; idx < (1 << (D - 1)); ++idx)
#pragma unroll
                        for (int f = 0; f < F; ++f) r[f] = C::add(r[f], C::from_f(wv[idx] * C::to_f(v[idx][f])));
                    if (DYDX) {
#pragma unroll
                        for (int g2 = 0; g2 < D - 1; ++g2) {
                            const int g3 = g2 >= j ? g2 + 1 : g2;
#pragma unroll
                            for (int idx = 0; idx < (1 << (D - 2)); ++idx) {
                                float w = c.scale[g3] * c.dp[g3];
                                int li = 0;
#pragma unroll
                                for (int ng = 0; ng < D - 2; ++ng) {
                                    const int d2 = ng >= g2 ? ng + 1 : ng;
                                    const int d3 = d2 >= j ? d2 + 1 : d2;
                                    if ((idx & (1 << ng)) == 0) w *= 1.0f - c.p[d3];
                                    else { w *= c.p[d3]; li += 1 << d2; }
                                }
                                const int ri = li + (1 << g2);
#pragma unroll
                                for (int f = 0; f < F; ++f) gr[f][g3] += w * (C::to_f(v[ri][f]) - C::to_f(v[li][f]));
                            }
                        }
                    }
                }
            }
        } else {  // CPfast, lotd_encoding.h:353-410
            PT Lv[D][F], Rv[D][F];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                load_feats<F>(g + (uint64_t)idx_cp_line<D>(c.res, c.cell[k], k) * c.n_feat + c.gfo, Lv[k], vec_ok);
                load_feats<F>(g + (uint64_t)idx_cp_line<D>(c.res, c.cell[k] + 1, k) * c.n_feat + c.gfo, Rv[k], vec_ok);
            }
            float acc[F];
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f] = 1.0f;
#pragma unroll
            for (int k = 0; k < D; ++k)
#pragma unroll
                for (int f = 0; f < F; ++f) acc[f] *= (1.0f - c.p[k]) * C::to_f(Lv[k][f]) + c.p[k] * C::to_f(Rv[k][f]);
#pragma unroll
            for (int f = 0; f < F; ++f) r[f] = C::from_f(acc[f]);
            if (DYDX) {
#pragma unroll
                for (int gd = 0; gd < D; ++gd) {
                    const float w = c.scale[gd] * c.dp[gd];
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        float t = w * (C::to_f(Rv[gd][f]) - C::to_f(Lv[gd][f]));
#pragma unroll
                        for (int k = 0; k < D; ++k)
                            if (k != gd) t *= (1.0f - c.p[k]) * C::to_f(Lv[k][f]) + c.p[k] * C::to_f(Rv[k][f]);
                        gr[f][gd] = t;
                    }
                }
            }
        }
    }
    PT* yo = y + (int64_t)i * ys_n + (int64_t)ofo * ys_f;
#pragma unroll
    for (int f = 0; f < F; ++f) st_cs(yo + (int64_t)f * ys_f, r[f]);
    if (DYDX) {
        float* go = dydx + (int64_t)i * ds_n + (int64_t)ofo * ds_f;
        if (ds_f == D && ((F * D) % 2 == 0) && ((reinterpret_cast<uintptr_t>(go) & 7u) == 0)) {
            // the thread's F * D derivatives are contiguous (row-major dy_dx): 8-byte stores
            const float* flat = &gr[0][0];
#pragma unroll
            for (int e = 0; e < F * D; e += 2) __stcs(reinterpret_cast<float2*>(go + e), make_float2(flat[e], flat[e + 1]));
        } else {
#pragma unroll
            for (int f = 0; f < F; ++f)
#pragma unroll
                for (int d = 0; d < D; ++d) st_cs(go + (int64_t)f * ds_f + d, gr[f][d]);
        }
    }
}

// =================================================================================================
// dL/dparam.  Reference: kernel_lod_backward_grid (lotd_encoding.h:467-711),
// kernel_lod_hashonly_backward_grid (lotd_hash_only.h:380-470).
// SECOND = true adds the second-order term d(dL/dx)/dparam . dL_ddLdx instead
// (kernel_lod_backward_input_backward_grid, lotd_encoding.h:764-1041).
// =================================================================================================
// Pseudo levels whose small sub-tables are accumulated in shared memory by lotd_bwd_param_priv_kernel (the plain kernel skips them).
constexpr int kMaxPriv = 64;
struct PrivPlan {
    uint32_t n;                 // number of privatised pseudo levels
    uint32_t n_batches;         // scenes with a shared-memory copy each
    uint32_t skip[NR3D_MAX_PSEUDO_LEVELS / 32];   // bit pl set: pseudo level pl belongs to the privatised kernel
    uint8_t pl[kMaxPriv];
};

// weight of every lattice corner: the n-linear weight (first order), or -- second order -- the sum over the derivative dimensions gd of
// gin[gd] * (+1 right / -1 left of gd) * prod_{d != gd} w_d  (the reference walks the 2^(D-1) face corners of every gd and scatters
// left and right separately, lotd_encoding.h:764-1041: 3x the reductions)
template <int D, bool SECOND>
__device__ __forceinline__ void corner_weights(const Ctx<D>& c, const float* gin, float* cw) {
#pragma unroll
    for (int idx = 0; idx < (1 << D); ++idx) {
        if (!SECOND) {
            float w = 1.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) w *= ((idx >> d) & 1) ? c.p[d] : 1.0f - c.p[d];
            cw[idx] = w;
        } else {
            float acc = 0.f;
#pragma unroll
            for (int gd = 0; gd < D; ++gd) {
                float w = gin[gd];
#pragma unroll
                for (int d = 0; d < D; ++d)
                    if (d != gd) w *= ((idx >> d) & 1) ? c.p[d] : 1.0f - c.p[d];
                acc += ((idx >> gd) & 1) ? w : -w;
            }
            cw[idx] = acc;
        }
    }
}

template <int D, int F, typename PT, bool SECOND>
__global__ void __launch_bounds__(kLotdThreads)
lotd_bwd_param_kernel(const __grid_constant__ LotdTable tab, const LotdIn in, const __grid_constant__ PrivPlan plan, const PT* __restrict__ dLdy,
                      int64_t gs_n, int64_t gs_f, const float* __restrict__ ddx, PT* __restrict__ grad_params) {
    using C = Cvt<PT>;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.N) return;
    const uint32_t pl = blockIdx.y;
    if (plan.skip[pl >> 5] & (1u << (pl & 31))) return;   // handled by lotd_bwd_param_priv_kernel
    Ctx<D> c;
    if (!lotd_setup<D, F>(tab, in, i, pl, c)) return;
    const PT* g = reinterpret_cast<const PT*>(in.params) + c.base;
    PT* gg = grad_params + c.base;
    const bool vec_ok = in.vec_ok;
    float grad[F];
    {
        const PT* gp = dLdy + (int64_t)i * gs_n + (int64_t)(pl * F) * gs_f;
#pragma unroll
        for (int f = 0; f < F; ++f) grad[f] = C::to_f(gp[(int64_t)f * gs_f]);
    }
    float gin[D];
    if (SECOND) {
#pragma unroll
        for (int d = 0; d < D; ++d) gin[d] = c.scale[d] * ddx[i * D + d] * c.dp[d];
    }
    if (is_nlinear(c.type)) {
        float cw[1 << D];
        corner_weights<D, SECOND>(c, gin, cw);
        nlinear_scatter<D, F, PT>(c, g, gg, GlobalSink<F, PT>{gg, c.n_feat, c.gfo, vec_ok}, cw, grad, vec_ok);
    } else if (c.type == NR3D_LOD_NPLANESUM) {
        if constexpr (D > 2) {
#pragma unroll 1
            for (int j = 0; j < D; ++j) {
                if (!SECOND) {  // lotd_encoding.h:597-651
#pragma unroll
                    for (int idx = 0; idx < (1 << (D - 1)); ++idx) {
                        uint32_t pp[D - 1];
                        float w = 1.0f;
#pragma unroll
                        for (int d2 = 0; d2 < D - 1; ++d2) {
                            const int d3 = d2 >= j ? d2 + 1 : d2;
                            if ((idx & (1 << d2)) == 0) { w *= 1.0f - c.p[d3]; pp[d2] = c.cell[d3]; }
                            else { w *= c.p[d3]; pp[d2] = c.cell[d3] + 1; }
                        }
                        float wg[F];
#pragma unroll
                        for (int f = 0; f < F; ++f) wg[f] = grad[f] * w;
                        scatter_add<F>(gg + (uint64_t)idx_nplane_sub<D>(c.res, pp, j) * c.n_feat + c.gfo, wg, vec_ok);
                    }
                } else {  // lotd_encoding.h:903-967
#pragma unroll
                    for (int g2 = 0; g2 < D - 1; ++g2) {
                        const int g3 = g2 >= j ? g2 + 1 : g2;
#pragma unroll
                        for (int idx = 0; idx < (1 << (D - 2)); ++idx) {
                            float w = gin[g3];
                            uint32_t pp[D - 1];
#pragma unroll
                            for (int ng = 0; ng < D - 2; ++ng) {
                                const int d2 = ng >= g2 ? ng + 1 : ng;
                                const int d3 = d2 >= j ? d2 + 1 : d2;
                                if ((idx & (1 << ng)) == 0) { w *= 1.0f - c.p[d3]; pp[d2] = c.cell[d3]; }
                                else { w *= c.p[d3]; pp[d2] = c.cell[d3] + 1; }
                            }
                            float wg[F];
                            pp[g2] = c.cell[g3];
#pragma unroll
                            for (int f = 0; f < F; ++f) wg[f] = grad[f] * (-w);
                            scatter_add<F>(gg + (uint64_t)idx_nplane_sub<D>(c.res, pp, j) * c.n_feat + c.gfo, wg, vec_ok);
                            pp[g2] = c.cell[g3] + 1;
#pragma unroll
                            for (int f = 0; f < F; ++f) wg[f] = grad[f] * w;
                            scatter_add<F>(gg + (uint64_t)idx_nplane_sub<D>(c.res, pp, j) * c.n_feat + c.gfo, wg, vec_ok);
                        }
                    }
                }
            }
        }
    } else {  // CPfast
        PT Lv[D][F], Rv[D][F];
        uint64_t il[D], ir[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            il[k] = (uint64_t)idx_cp_line<D>(c.res, c.cell[k], k) * c.n_feat + c.gfo;
            ir[k] = (uint64_t)idx_cp_line<D>(c.res, c.cell[k] + 1, k) * c.n_feat + c.gfo;
            load_feats<F>(g + il[k], Lv[k], vec_ok);
            load_feats<F>(g + ir[k], Rv[k], vec_ok);
        }
        if (!SECOND) {  // lotd_encoding.h:653-705
#pragma unroll
            for (int gd = 0; gd < D; ++gd) {
                float gl[F], a[F], b[F];
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    float t = grad[f];
#pragma unroll
                    for (int k = 0; k < D; ++k)
                        if (k != gd) t *= (1.0f - c.p[k]) * C::to_f(Lv[k][f]) + c.p[k] * C::to_f(Rv[k][f]);
                    gl[f] = t;
                    a[f] = gl[f] * (1.0f - c.p[gd]);
                    b[f] = gl[f] * c.p[gd];
                }
                scatter_add<F>(gg + il[gd], a, vec_ok);
                scatter_add<F>(gg + ir[gd], b, vec_ok);
            }
        } else {  // lotd_encoding.h:970-1038 (intended maths; the reference indexes a 2-wide array with F_pl)
#pragma unroll
            for (int ld = 0; ld < D; ++ld) {
#pragma unroll
                for (int gd = 0; gd < D; ++gd) {
                    const float wl = (ld != gd) ? 1.0f - c.p[ld] : -1.0f;
                    const float wr = (ld != gd) ? c.p[ld] : 1.0f;
                    float a[F], b[F];
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        float t = grad[f] * c.scale[gd] * ddx[i * D + gd] * c.dp[gd];
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            if (k == ld) continue;
                            const float nl = (k != gd) ? 1.0f - c.p[k] : -1.0f;
                            const float nr = (k != gd) ? c.p[k] : 1.0f;
                            t *= nl * C::to_f(Lv[k][f]) + nr * C::to_f(Rv[k][f]);
                        }
                        a[f] = t * wl;
                        b[f] = t * wr;
                    }
                    scatter_add<F>(gg + il[ld], a, vec_ok);
                    scatter_add<F>(gg + ir[ld], b, vec_ok);
                }
            }
        }
    }
}

// =================================================================================================
// dL/dparam for pseudo levels with tiny, heavily shared sub-tables (small Dense, CP, the lines of VM): a persistent grid of CTAs,
// each with a shared-memory copy of the sub-table for every scene, walks the points; the copies are added to the gradient table
// once per CTA at the end.  2 Mi points onto a 96-entry line table are ~90 000 reductions per entry when sent to L2 directly --
// the L2 slice applies same-address reductions one after the other -- and a few hundred after privatisation.
// grid = (persistent CTAs, privatised pseudo levels); dynamic shared memory = n_batches * n_small * F floats.
// =================================================================================================
template <int D, int F, typename PT, bool SECOND>
__global__ void __launch_bounds__(kLotdThreads)
lotd_bwd_param_priv_kernel(const __grid_constant__ LotdTable tab, const LotdIn in, const __grid_constant__ PrivPlan plan, const PT* __restrict__ dLdy,
                           int64_t gs_n, int64_t gs_f, const float* __restrict__ ddx, PT* __restrict__ grad_params) {
    using C = Cvt<PT>;
    extern __shared__ float sm[];
    const uint32_t pl = plan.pl[blockIdx.y];
    const LevelDesc& L = tab.lv[tab.map_level[pl]];
    const uint32_t n_small = small_entries(L.type, L.res, D, L.size);
    const uint32_t n_slots = plan.n_batches * n_small * F;
    for (uint32_t e = threadIdx.x; e < n_slots; e += blockDim.x) sm[e] = 0.f;
    __syncthreads();
    const bool vec_ok = in.vec_ok;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < in.N; i += (uint64_t)gridDim.x * blockDim.x) {
        Ctx<D> c;
        if (!lotd_setup<D, F>(tab, in, i, pl, c)) continue;
        const PT* g = reinterpret_cast<const PT*>(in.params) + c.base;
        PT* gg = grad_params + c.base;
        float grad[F];
        {
            const PT* gp = dLdy + (int64_t)i * gs_n + (int64_t)(pl * F) * gs_f;
#pragma unroll
            for (int f = 0; f < F; ++f) grad[f] = C::to_f(gp[(int64_t)f * gs_f]);
        }
        float gin[D];
        if (SECOND) {
#pragma unroll
            for (int d = 0; d < D; ++d) gin[d] = c.scale[d] * ddx[i * D + d] * c.dp[d];
        }
        float cw[1 << D];
        corner_weights<D, SECOND>(c, gin, cw);
        const bool mine = c.batch < plan.n_batches;     // scenes beyond the planned count go straight to the table
        const PrivSink<F, PT> sink{GlobalSink<F, PT>{gg, c.n_feat, c.gfo, vec_ok}, sm + (size_t)(mine ? c.batch : 0) * n_small * F, mine ? n_small : 0u};
        nlinear_scatter<D, F, PT>(c, g, gg, sink, cw, grad, vec_ok);
    }
    __syncthreads();
    const uint32_t gfo = (uint32_t)tab.map_cnt[pl] * F;
    for (uint32_t e = threadIdx.x; e < plan.n_batches * n_small; e += blockDim.x) {
        float v[F];
        bool any = false;
#pragma unroll
        for (int f = 0; f < F; ++f) { v[f] = sm[e * F + f]; any |= (v[f] != 0.f); }
        if (!any) continue;
        const uint32_t b = e / n_small, ent = e - b * n_small;
        const uint64_t boff = in.batch_offsets ? (uint64_t)in.batch_offsets[b] : (uint64_t)b * tab.n_params;
        scatter_add<F>(grad_params + boff + L.offset + (uint64_t)ent * L.n_feat + gfo, v, vec_ok);
    }
}

// =================================================================================================
// second order: d(dL/dx)/dx . dL_ddLdx  -> atomicAdd into dL_dx [N, D].
// Reference: kernel_lod_backward_input_backward_input (lotd_encoding.h:1157-1298) with
// bwd_input_bwd_input_n_linear (lotd_encoding.h:1043-1155). Only Dense / Hash / VM / VecZMatXoY contribute.
// =================================================================================================
template <int D, int F, typename PT>
__global__ void __launch_bounds__(kLotdThreads)
lotd_bwdbwd_input_kernel(const __grid_constant__ LotdTable tab, const LotdIn in, const PT* __restrict__ dLdy, int64_t gs_n,
                         int64_t gs_f, const float* __restrict__ ddx, float* __restrict__ dLdx) {
    using C = Cvt<PT>;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.N) return;
    const uint32_t pl = blockIdx.y;
    Ctx<D> c;
    if (!lotd_setup<D, F>(tab, in, i, pl, c)) return;
    if (!(c.type == NR3D_LOD_DENSE || c.type == NR3D_LOD_HASH || c.type == NR3D_LOD_VM || c.type == NR3D_LOD_VECZMATXOY)) return;
    const PT* g = reinterpret_cast<const PT*>(in.params) + c.base;
    const bool vec_ok = in.vec_ok;
    float grad[F];
    {
        const PT* gp = dLdy + (int64_t)i * gs_n + (int64_t)(pl * F) * gs_f;
#pragma unroll
        for (int f = 0; f < F; ++f) grad[f] = C::to_f(gp[(int64_t)f * gs_f]);
    }
    // S[idx] = sum_f corner_value[idx][f] * dL_dy[f]
    float S[1 << D];
#pragma unroll
    for (int idx = 0; idx < (1 << D); ++idx) {
        uint32_t pos[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pos[d] = c.cell[d] + ((idx >> d) & 1);
        const float s = corner_dot<D, F, PT>(c, g, pos, grad, vec_ok);
        S[idx] = s;
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
            if (smooth) {  // diagonal of the Hessian
                uint32_t pos[D];
                int li;
                const float w = face_weight<D>(c, gd, idx, gin_diag[gd], pos, &li);
                out += S[li] * (-w);
                out += S[li + (1 << gd)] * w;
            }
            if constexpr (D > 1) {
#pragma unroll
                for (int og = 0; og < D - 1; ++og) {
                    const int o = og >= gd ? og + 1 : og;  // the other derivative dim
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
        }
        atomicAdd(dLdx + i * D + gd, out);
    }
}

// =================================================================================================
// dL_dx[n,d] = sum_j dL_dy[n,j] * dy_dx[n,j,d]   (reference: at::mul + at::sum_out, lotd_hash_only.h:852-856)
// dL_ddLdy[n,j] = sum_d dL_ddLdx[n,d] * dy_dx[n,j,d]   (lotd_hash_only.h:995-999)
// =================================================================================================
template <int D, typename PT>
__global__ void __launch_bounds__(kLotdThreads)
lotd_bwd_input_kernel(uint64_t N, uint32_t n_enc, const PT* __restrict__ dLdy, int64_t gs_n, int64_t gs_f,
                      const float* __restrict__ dydx, int64_t ds_n, int64_t ds_f, float* __restrict__ dLdx) {
    using C = Cvt<PT>;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    const PT* gp = dLdy + (int64_t)i * gs_n;
    const float* dp = dydx + (int64_t)i * ds_n;
#pragma unroll 4
    for (uint32_t j = 0; j < n_enc; ++j) {
        const float gy = C::to_f(gp[(int64_t)j * gs_f]);
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] += gy * dp[(int64_t)j * ds_f + d];
    }
#pragma unroll
    for (int d = 0; d < D; ++d) dLdx[i * D + d] = acc[d];
}

// Same contraction for ROW-MAJOR dy_dx ([N, n_enc * D] contiguous: generic metas, forest, the sorted fast path): a thread per point
// would read 32 different lines per load there, so a warp takes one point per pass and streams its n_enc * D derivatives with coalesced
// 128-byte loads; the D partial sums meet through shuffles.  Bound: HBM stream of dy_dx (12 B per feature and point).
template <int D, typename PT>
__global__ void __launch_bounds__(kLotdThreads)
lotd_bwd_input_rows_kernel(uint64_t N, uint32_t n_enc, const PT* __restrict__ dLdy, int64_t gs_n, const float* __restrict__ dydx,
                           float* __restrict__ dLdx) {
    using C = Cvt<PT>;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t width = n_enc * D;
    for (uint64_t i = warp; i < N; i += n_warps) {
        const float* dp = dydx + i * width;
        const PT* gp = dLdy + (int64_t)i * gs_n;
        float acc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] = 0.f;
        for (uint32_t e = lane; e < width; e += 32) {
            const float v = __ldcs(dp + e) * C::to_f(gp[e / D]);
            const uint32_t dd = e % D;
#pragma unroll
            for (int d = 0; d < D; ++d) acc[d] += (dd == (uint32_t)d) ? v : 0.f;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], m);
        }
        if (lane < (uint32_t)D) {
            float out = acc[0];
#pragma unroll
            for (int d = 1; d < D; ++d) out = (lane == (uint32_t)d) ? acc[d] : out;
            dLdx[i * D + lane] = out;
        }
    }
}

template <int D, typename PT>
__global__ void __launch_bounds__(kLotdThreads)
lotd_ddLdy_kernel(uint64_t N, uint32_t n_enc, const float* __restrict__ ddx, const float* __restrict__ dydx, int64_t ds_n,
                  int64_t ds_f, PT* __restrict__ out) {
    using C = Cvt<PT>;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * n_enc) return;
    const uint64_t i = t / n_enc;
    const uint32_t j = (uint32_t)(t - i * n_enc);
    const float* dp = dydx + (int64_t)i * ds_n + (int64_t)j * ds_f;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) acc += ddx[i * D + d] * dp[d];
    out[t] = C::from_f(acc);
}

// =================================================================================================
// lod_get_grid_index (lotd_encoding.h:1300-1433): int64 [N, n_enc, 2^D], Dense / Hash only
// =================================================================================================
template <int D, int F>
__global__ void __launch_bounds__(kLotdThreads)
lotd_grid_index_kernel(const __grid_constant__ LotdTable tab, const LotdIn in, int64_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.N) return;
    const uint32_t pl = blockIdx.y;
    Ctx<D> c;
    if (!lotd_setup<D, F>(tab, in, i, pl, c)) return;
    int64_t* o = out + ((int64_t)i * tab.n_enc + (int64_t)pl * F) * (1 << D);
#pragma unroll
    for (int idx = 0; idx < (1 << D); ++idx) {
        uint32_t pos[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pos[d] = c.cell[d] + ((idx >> d) & 1);
        uint32_t cellidx;
        if (c.type == NR3D_LOD_DENSE) cellidx = idx_dense<D>(c.res, pos);
        else if (c.type == NR3D_LOD_HASH) cellidx = idx_hash<D>(pos, c.size);
        else continue;
        // the reference does this arithmetic in uint32 (batch_offset + level_offset + index + f)
        const uint32_t ind = (uint32_t)c.base + cellidx * c.n_feat + c.gfo;
#pragma unroll
        for (int f = 0; f < F; ++f) o[idx + f * (1 << D)] = (int64_t)(uint32_t)(ind + f);
    }
}

// ------------------------------------------------------------------------------------------------
// per-dimension launchers (defined in lotd_d{2,3,4}.cu through NR3D_LOTD_DEFINE_DIM)
// ------------------------------------------------------------------------------------------------
struct LotdLaunch {
    LotdTable tab;
    LotdIn in;
    int fpl;        // 2, 4, 8
    int half;       // param dtype: 0 float, 1 half
    cudaStream_t stream;
    uint32_t n_batches = 0;   // scenes behind `params` if the caller knows (0: unknown -> no shared-memory privatisation)
};

// which pseudo levels go to the privatised dL/dparam kernel (host side)
template <int D>
inline PrivPlan make_priv_plan(const LotdLaunch& L, int F, size_t* smem_bytes) {
    PrivPlan plan;
    memset(&plan, 0, sizeof(plan));
    *smem_bytes = 0;
    uint32_t B = L.n_batches;
    if (B == 0 && !L.in.batch_inds) B = L.in.batch_data_size ? (uint32_t)(L.in.N / L.in.batch_data_size) : 1u;
    if (B == 0) return plan;
    plan.n_batches = B;
    for (uint32_t pl = 0; pl < L.tab.n_pseudo && plan.n < (uint32_t)kMaxPriv; ++pl) {
        const LevelDesc& lv = L.tab.lv[L.tab.map_level[pl]];
        const uint32_t n_small = small_entries(lv.type, lv.res, D, lv.size);
        if (n_small == 0 || (int32_t)L.tab.map_level[pl] > L.in.max_level) continue;
        const size_t bytes = (size_t)B * n_small * F * sizeof(float);
        // worth it when the points outnumber the privatised entries and every scene's copy fits the default shared-memory window
        if (bytes > 48 * 1024 || L.in.N < 2ull * B * n_small) continue;
        plan.pl[plan.n++] = (uint8_t)pl;
        plan.skip[pl >> 5] |= 1u << (pl & 31);
        if (bytes > *smem_bytes) *smem_bytes = bytes;
    }
    return plan;
}

// fills L from the C-ABI arguments (defined in lotd_api.cu; shared with lotd_forest.cu)
int build_launch(const nr3d_lotd_meta* m, int32_t input_dtype, int32_t param_dtype, uint64_t N, const void* x, const void* params,
                 const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size, int32_t max_level, void* stream,
                 LotdLaunch& L);

template <int D> int lotd_launch_fwd(const LotdLaunch& L, void* y, int64_t ys_n, int64_t ys_f, float* dydx, int64_t ds_n, int64_t ds_f);
template <int D> int lotd_launch_bwd_param(const LotdLaunch& L, const void* dLdy, int64_t gs_n, int64_t gs_f, const float* ddx, void* grad_params);
template <int D> int lotd_launch_bwdbwd_input(const LotdLaunch& L, const void* dLdy, int64_t gs_n, int64_t gs_f, const float* ddx, float* dLdx);
template <int D> int lotd_launch_bwd_input(const LotdLaunch& L, const void* dLdy, int64_t gs_n, int64_t gs_f, const float* dydx, int64_t ds_n, int64_t ds_f, float* dLdx);
template <int D> int lotd_launch_ddLdy(const LotdLaunch& L, const float* ddx, const float* dydx, int64_t ds_n, int64_t ds_f, void* out);
template <int D> int lotd_launch_grid_index(const LotdLaunch& L, int64_t* out);

#define NR3D_LOTD_DISPATCH_F_PT(...)                                           \
    switch (L.fpl * 2 + L.half) {                                              \
    case 4:  { constexpr int F = 2; using PT = float;  __VA_ARGS__; } break;         \
    case 5:  { constexpr int F = 2; using PT = __half; __VA_ARGS__; } break;         \
    case 8:  { constexpr int F = 4; using PT = float;  __VA_ARGS__; } break;         \
    case 9:  { constexpr int F = 4; using PT = __half; __VA_ARGS__; } break;         \
    case 16: { constexpr int F = 8; using PT = float;  __VA_ARGS__; } break;         \
    case 17: { constexpr int F = 8; using PT = __half; __VA_ARGS__; } break;         \
    default: return fail("LoTDEncoding: `n_feat_per_pseudo_lvl` must be one of [2,4,8]"); }

#define NR3D_LOTD_DEFINE_DIM(D)                                                                                          \
    template <> int lotd_launch_fwd<D>(const LotdLaunch& L, void* y, int64_t ys_n, int64_t ys_f, float* dydx,            \
                                       int64_t ds_n, int64_t ds_f) {                                                      \
        if (L.in.N == 0) return 0;                                                                                        \
        const dim3 grid((unsigned)div_up<uint64_t>(L.in.N, kLotdThreads), L.tab.n_pseudo, 1);                             \
        NR3D_LOTD_DISPATCH_F_PT(                                                                                          \
            if (dydx) lotd_fwd_kernel<D, F, PT, true><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, (PT*)y, ys_n, ys_f, dydx, ds_n, ds_f); \
            else lotd_fwd_kernel<D, F, PT, false><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, (PT*)y, ys_n, ys_f, nullptr, 0, 0))        \
        NR3D_LAUNCH_CHECK("lotd_fwd");                                                                                    \
        return 0;                                                                                                         \
    }                                                                                                                     \
    template <> int lotd_launch_bwd_param<D>(const LotdLaunch& L, const void* dLdy, int64_t gs_n, int64_t gs_f,          \
                                             const float* ddx, void* gp) {                                                \
        if (L.in.N == 0) return 0;                                                                                        \
        const dim3 grid((unsigned)div_up<uint64_t>(L.in.N, kLotdThreads), L.tab.n_pseudo, 1);                             \
        size_t priv_smem = 0;                                                                                             \
        const PrivPlan plan = make_priv_plan<D>(L, L.fpl, &priv_smem);                                                    \
        if (plan.n < L.tab.n_pseudo) {                                                                                    \
            NR3D_LOTD_DISPATCH_F_PT(                                                                                      \
                if (ddx) lotd_bwd_param_kernel<D, F, PT, true><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, plan, (const PT*)dLdy, gs_n, gs_f, ddx, (PT*)gp); \
                else lotd_bwd_param_kernel<D, F, PT, false><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, plan, (const PT*)dLdy, gs_n, gs_f, nullptr, (PT*)gp)) \
            NR3D_LAUNCH_CHECK("lotd_bwd_param");                                                                          \
        }                                                                                                                 \
        if (plan.n) {                                                                                                     \
            const uint64_t want = div_up<uint64_t>(L.in.N, kLotdThreads);                                                 \
            const dim3 pgrid((unsigned)(want < (uint64_t)kSMs * 3 ? want : (uint64_t)kSMs * 3), plan.n, 1);               \
            NR3D_LOTD_DISPATCH_F_PT(                                                                                      \
                if (ddx) lotd_bwd_param_priv_kernel<D, F, PT, true><<<pgrid, kLotdThreads, priv_smem, L.stream>>>(L.tab, L.in, plan, (const PT*)dLdy, gs_n, gs_f, ddx, (PT*)gp); \
                else lotd_bwd_param_priv_kernel<D, F, PT, false><<<pgrid, kLotdThreads, priv_smem, L.stream>>>(L.tab, L.in, plan, (const PT*)dLdy, gs_n, gs_f, nullptr, (PT*)gp)) \
            NR3D_LAUNCH_CHECK("lotd_bwd_param_priv");                                                                     \
        }                                                                                                                 \
        return 0;                                                                                                         \
    }                                                                                                                     \
    template <> int lotd_launch_bwdbwd_input<D>(const LotdLaunch& L, const void* dLdy, int64_t gs_n, int64_t gs_f,       \
                                                const float* ddx, float* dLdx) {                                          \
        if (L.in.N == 0) return 0;                                                                                        \
        const dim3 grid((unsigned)div_up<uint64_t>(L.in.N, kLotdThreads), L.tab.n_pseudo, 1);                             \
        NR3D_LOTD_DISPATCH_F_PT((lotd_bwdbwd_input_kernel<D, F, PT><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, (const PT*)dLdy, gs_n, gs_f, ddx, dLdx))) \
        NR3D_LAUNCH_CHECK("lotd_bwdbwd_input");                                                                           \
        return 0;                                                                                                         \
    }                                                                                                                     \
    template <> int lotd_launch_bwd_input<D>(const LotdLaunch& L, const void* dLdy, int64_t gs_n, int64_t gs_f,          \
                                             const float* dydx, int64_t ds_n, int64_t ds_f, float* dLdx) {                \
        if (L.in.N == 0) return 0;                                                                                        \
        const unsigned grid = (unsigned)div_up<uint64_t>(L.in.N, kLotdThreads);                                           \
        if (gs_f == 1 && ds_f == D && ds_n == (int64_t)L.tab.n_enc * D && L.tab.n_enc * D >= 96) { /* wide row-major dy_dx: warp per point (measured: 2x at 96 floats per row, 0.5x at 60) */ \
            const unsigned rgrid = (unsigned)(div_up<uint64_t>(L.in.N, kLotdThreads / 32) < (uint64_t)kSMs * 16 ? div_up<uint64_t>(L.in.N, kLotdThreads / 32) : (uint64_t)kSMs * 16); \
            if (L.half) lotd_bwd_input_rows_kernel<D, __half><<<rgrid, kLotdThreads, 0, L.stream>>>(L.in.N, L.tab.n_enc, (const __half*)dLdy, gs_n, dydx, dLdx); \
            else lotd_bwd_input_rows_kernel<D, float><<<rgrid, kLotdThreads, 0, L.stream>>>(L.in.N, L.tab.n_enc, (const float*)dLdy, gs_n, dydx, dLdx); \
        } else                                                                                                            \
        if (L.half) lotd_bwd_input_kernel<D, __half><<<grid, kLotdThreads, 0, L.stream>>>(L.in.N, L.tab.n_enc, (const __half*)dLdy, gs_n, gs_f, dydx, ds_n, ds_f, dLdx); \
        else lotd_bwd_input_kernel<D, float><<<grid, kLotdThreads, 0, L.stream>>>(L.in.N, L.tab.n_enc, (const float*)dLdy, gs_n, gs_f, dydx, ds_n, ds_f, dLdx); \
        NR3D_LAUNCH_CHECK("lotd_bwd_input");                                                                              \
        return 0;                                                                                                         \
    }                                                                                                                     \
    template <> int lotd_launch_ddLdy<D>(const LotdLaunch& L, const float* ddx, const float* dydx, int64_t ds_n,         \
                                         int64_t ds_f, void* out) {                                                       \
        if (L.in.N == 0) return 0;                                                                                        \
        const unsigned grid = (unsigned)div_up<uint64_t>(L.in.N * L.tab.n_enc, kLotdThreads);                             \
        if (L.half) lotd_ddLdy_kernel<D, __half><<<grid, kLotdThreads, 0, L.stream>>>(L.in.N, L.tab.n_enc, ddx, dydx, ds_n, ds_f, (__half*)out); \
        else lotd_ddLdy_kernel<D, float><<<grid, kLotdThreads, 0, L.stream>>>(L.in.N, L.tab.n_enc, ddx, dydx, ds_n, ds_f, (float*)out); \
        NR3D_LAUNCH_CHECK("lotd_ddLdy");                                                                                  \
        return 0;                                                                                                         \
    }                                                                                                                     \
    template <> int lotd_launch_grid_index<D>(const LotdLaunch& L, int64_t* out) {                                       \
        if (L.in.N == 0) return 0;                                                                                        \
        const dim3 grid((unsigned)div_up<uint64_t>(L.in.N, kLotdThreads), L.tab.n_pseudo, 1);                             \
        switch (L.fpl) {                                                                                                  \
        case 2: lotd_grid_index_kernel<D, 2><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, out); break;               \
        case 4: lotd_grid_index_kernel<D, 4><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, out); break;               \
        case 8: lotd_grid_index_kernel<D, 8><<<grid, kLotdThreads, 0, L.stream>>>(L.tab, L.in, out); break;               \
        default: return fail("LoTDEncoding: `n_feat_per_pseudo_lvl` must be one of [2,4,8]"); }                           \
        NR3D_LAUNCH_CHECK("lotd_grid_index");                                                                             \
        return 0;                                                                                                         \
    }

}  // namespace nr3d
