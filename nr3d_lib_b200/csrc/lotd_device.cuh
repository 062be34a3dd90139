// lotd_device.cuh -- device-side maths of the LoTD encoder (all level types), written for sm_100a.
//
// Parity notes (what each block reproduces in the reference):
//   pos_fract            csrc/lotd/include/lotd/lotd_cuda.h:959-1077   (scale = res-2, +0.5, floor, smoothstep)
//   index functions      csrc/lotd/include/lotd/lotd_cuda.h:92-296     (uint32 arithmetic, bit-exact)
//   corner values        csrc/lotd/include/lotd/lotd_cuda.h:298-492
//   n-linear weights     csrc/lotd/include/lotd/linear_interpolate.cuh:92-150
//   gradient scatter     csrc/lotd/include/lotd/lotd_cuda.h:494-829
#pragma once
#include "common.cuh"

namespace nr3d {

struct LevelDesc {
    uint32_t res[4];
    uint32_t type, n_feat, size, offset;
};
// Kernel-parameter copy of the meta: 1.6 KB in the constant bank (the reference passes a 2.5 KB LoDMetaRef by value).
struct LotdTable {
    LevelDesc lv[NR3D_MAX_LEVELS];
    uint8_t map_level[NR3D_MAX_PSEUDO_LEVELS];
    uint8_t map_cnt[NR3D_MAX_PSEUDO_LEVELS];
    uint32_t n_levels, n_pseudo, n_enc, n_params, interp, fpl, pad0, pad1;
};

struct LotdIn {
    uint64_t N;
    const float* x;
    const void* params;
    const int64_t* batch_inds;
    const int64_t* batch_offsets;
    uint32_t batch_data_size;
    int32_t max_level;
    uint32_t vec_ok;  // 1: level base pointers are 8-byte aligned (no user-supplied batch_offsets)
    uint32_t x_half;  // 1: `x` points at __half coordinates (the reference's <half, half, half> combination, lotd_hash_only.h:776)
};

template <int D>
struct Ctx {
    uint32_t res[D];
    uint32_t cell[D];
    float scale[D], p[D], dp[D], d2p[D];
    uint32_t type, n_feat, size, gfo;
    uint32_t batch;  // scene (batch) index of the point
    uint64_t base;   // element offset of this level's table inside `params`
};

// returns false when the (point, pseudo level) pair is skipped (level > max_level or batch index < 0)
template <int D, int F>
__device__ __forceinline__ bool lotd_setup(const LotdTable& tab, const LotdIn& in, uint64_t i, uint32_t pl, Ctx<D>& c) {
    const uint32_t level = tab.map_level[pl];
    if ((int32_t)level > in.max_level) return false;
    uint32_t batch_ind = 0;
    if (in.batch_inds) {
        const int64_t b = in.batch_inds[i];
        if (b < 0) return false;
        batch_ind = (uint32_t)b;
    } else if (in.batch_data_size) {
        batch_ind = (uint32_t)(i / in.batch_data_size);
    }
    const uint64_t batch_offset = in.batch_offsets ? (uint64_t)in.batch_offsets[batch_ind] : (uint64_t)batch_ind * tab.n_params;
    const LevelDesc& L = tab.lv[level];
    c.base = batch_offset + L.offset;
    c.batch = batch_ind;
    c.type = L.type;
    c.n_feat = L.n_feat;
    c.size = L.size;
    c.gfo = (uint32_t)tab.map_cnt[pl] * F;
    const float* xp = in.x + i * D;
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        c.res[d] = L.res[d];
        c.scale[d] = (float)(c.res[d] - 2u);
        float v;
        if (in.x_half) {
            // half points: the reference computes position, cell and fraction in HALF precision (COMPUTE_T = __half, lotd_cuda.h:959-977:
            // val = x * scale + 0.5 with two roundings, cell = floor(val), val -= cell) -- reproduced so that the cells are the reference's;
            // the interpolation weights are then formed in fp32 from that half fraction (the reference multiplies them out in half)
            const __half xh = reinterpret_cast<const __half*>(in.x)[i * D + d];
            const __half sc = __uint2half_rn(c.res[d] - 2u);
            const __half val = __hadd(__hmul(xh, sc), __float2half_rn(0.5f));
            const float fl = floorf(__half2float(val));
            c.cell[d] = (uint32_t)fl;
            c.scale[d] = __half2float(sc);
            v = __half2float(__hsub(val, __uint2half_rn(c.cell[d])));
        } else {
            v = xp[d] * c.scale[d] + 0.5f;
            const float fl = floorf(v);
            c.cell[d] = (uint32_t)fl;
            v -= (float)c.cell[d];
        }
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

// ------------------------------------------------------------------------------------------------
// index functions (all uint32, must be bit-exact)
// ------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ uint32_t idx_dense(const uint32_t* res, const uint32_t* pos) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {  // last dim contiguous
        index += pos[d] * stride;
        stride *= res[d];
    }
    return index;
}
template <int D>
__device__ __forceinline__ uint32_t idx_hash(const uint32_t* pos, uint32_t size) {
    constexpr uint32_t primes[4] = {1u, 2654435761u, 805459861u, 3674653429u};
    uint32_t h = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) h ^= pos[d] * primes[d];
    return h % size;
}
// "N-plane" index of the plane that drops dimension `jump` (no per-plane offset: reference quirk Q3)
template <int D>
__device__ __forceinline__ uint32_t idx_nplane(const uint32_t* res, const uint32_t* pos, int jump) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (int d2 = 0; d2 < D - 1; ++d2) {
        const int d3 = d2 >= jump ? d2 + 1 : d2;
        index += pos[D - 1 - d3] * stride;
        stride *= res[D - 1 - d3];
    }
    return index;
}
// NPlaneSum variant: plane-local coordinates + offset jump*stride (self-consistent for cubic res only)
template <int D>
__device__ __forceinline__ uint32_t idx_nplane_sub(const uint32_t* res, const uint32_t* pos_plane, int jump) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (int d2 = 0; d2 < D - 1; ++d2) {
        const int d3 = d2 >= jump ? d2 + 1 : d2;
        index += pos_plane[D - 2 - d2] * stride;
        stride *= res[D - 1 - d3];
    }
    return (uint32_t)jump * stride + index;
}
template <int D>
__device__ __forceinline__ uint32_t idx_cp_line(const uint32_t* res, uint32_t pos_line, int line_dim) {
    uint32_t acc = 0;
#pragma unroll
    for (int d = 0; d < D; ++d)
        if (d < line_dim) acc += res[d];
    return acc + pos_line;
}
// VM: lines [sum(res)] then planes; plane k drops dimension k
template <int D>
__device__ __forceinline__ void idx_vm(const uint32_t* res, const uint32_t* pos, uint32_t* plane, uint32_t* line) {
    uint32_t acc_line = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        line[k] = acc_line + pos[k];
        acc_line += res[k];
    }
    uint32_t acc_plane = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const int rev_jump = D - 1 - k;
        uint32_t stride = 1, index = 0;
#pragma unroll
        for (int d2 = 0; d2 < D - 1; ++d2) {
            const int d3 = d2 >= rev_jump ? d2 + 1 : d2;
            index += pos[D - 1 - d3] * stride;
            stride *= res[D - 1 - d3];
        }
        plane[k] = acc_line + acc_plane + index;
        acc_plane += stride;
    }
}

// ------------------------------------------------------------------------------------------------
// parameter loads
// ------------------------------------------------------------------------------------------------
template <int F>
__device__ __forceinline__ void load_feats(const float* g, float* v, bool vec_ok) {
    if (vec_ok) {
#pragma unroll
        for (int f = 0; f < F; f += 2) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(g + f));
            v[f] = t.x;
            v[f + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int f = 0; f < F; ++f) v[f] = __ldg(g + f);
    }
}
template <int F>
__device__ __forceinline__ void load_feats(const __half* g, __half* v, bool vec_ok) {
    if (vec_ok) {
#pragma unroll
        for (int f = 0; f < F; f += 2) {  // level / scene offsets are even -> 4-byte aligned
            const __half2 t = __ldg(reinterpret_cast<const __half2*>(g + f));
            v[f] = __low2half(t);
            v[f + 1] = __high2half(t);
        }
    } else {  // user-supplied batch_offsets may be odd: scalar accesses
#pragma unroll
        for (int f = 0; f < F; ++f) v[f] = __ldg(g + f);
    }
}

// value of one lattice corner for the "n-linear" level types
template <int D, int F, typename PT>
__device__ __forceinline__ void corner_val(const Ctx<D>& c, const PT* __restrict__ g, const uint32_t* pos, PT* v, bool vec_ok) {
    using C = Cvt<PT>;
    switch (c.type) {
    case NR3D_LOD_DENSE:
        load_feats<F>(g + (uint64_t)idx_dense<D>(c.res, pos) * c.n_feat + c.gfo, v, vec_ok);
        break;
    case NR3D_LOD_HASH:
        load_feats<F>(g + (uint64_t)idx_hash<D>(pos, c.size) * c.n_feat + c.gfo, v, vec_ok);
        break;
    case NR3D_LOD_VM:
        if constexpr (D == 3) {
            uint32_t pl[D], ln[D];
            idx_vm<D>(c.res, pos, pl, ln);
#pragma unroll
            for (int f = 0; f < F; ++f) v[f] = C::zero();
#pragma unroll
            for (int k = 0; k < D; ++k) {
                PT a[F], b[F];
                load_feats<F>(g + (uint64_t)pl[k] * c.n_feat + c.gfo, a, vec_ok);
                load_feats<F>(g + (uint64_t)ln[k] * c.n_feat + c.gfo, b, vec_ok);
#pragma unroll
                for (int f = 0; f < F; ++f) v[f] = C::add(v[f], C::from_f(C::to_f(a[f]) * C::to_f(b[f])));
            }
        }
        break;
    case NR3D_LOD_VECZMATXOY:
        if constexpr (D == 3) {
            const uint32_t ln = pos[2];
            const uint32_t pl = c.res[2] + pos[1] + pos[0] * c.res[0];
            PT a[F], b[F];
            load_feats<F>(g + (uint64_t)pl * c.n_feat + c.gfo, a, vec_ok);
            load_feats<F>(g + (uint64_t)ln * c.n_feat + c.gfo, b, vec_ok);
#pragma unroll
            for (int f = 0; f < F; ++f) v[f] = C::from_f(C::to_f(a[f]) * C::to_f(b[f]));
        }
        break;
    case NR3D_LOD_NPLANEMUL: {
        float r[F];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            PT a[F];
            load_feats<F>(g + (uint64_t)idx_nplane<D>(c.res, pos, j) * c.n_feat + c.gfo, a, vec_ok);
#pragma unroll
            for (int f = 0; f < F; ++f) r[f] = (j == 0) ? C::to_f(a[f]) : r[f] * C::to_f(a[f]);
        }
#pragma unroll
        for (int f = 0; f < F; ++f) v[f] = C::from_f(r[f]);
    } break;
    case NR3D_LOD_CP: {
        float r[F];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            PT a[F];
            load_feats<F>(g + (uint64_t)idx_cp_line<D>(c.res, pos[k], k) * c.n_feat + c.gfo, a, vec_ok);
#pragma unroll
            for (int f = 0; f < F; ++f) r[f] = (k == 0) ? C::to_f(a[f]) : r[f] * C::to_f(a[f]);
        }
#pragma unroll
        for (int f = 0; f < F; ++f) v[f] = C::from_f(r[f]);
    } break;
    default:
#pragma unroll
        for (int f = 0; f < F; ++f) v[f] = C::zero();
        break;
    }
}

// values of all 2^D lattice corners of one (point, pseudo level).  Same arithmetic (and rounding points) as corner_val per corner,
// but every distinct table entry is loaded once: a VM corner is sum_k plane_k * line_k with only 4 distinct plane entries and 2 distinct
// line entries per k (18 loads instead of 48), a CP corner is the product of D line entries out of 2 D (6 loads instead of 24).
template <int D, int F, typename PT>
__device__ __forceinline__ void all_corner_vals(const Ctx<D>& c, const PT* __restrict__ g, PT (&v)[1 << D][F], bool vec_ok) {
    using C = Cvt<PT>;
    if (c.type == NR3D_LOD_VM) {
        if constexpr (D == 3) {
#pragma unroll
            for (int idx = 0; idx < (1 << D); ++idx)
#pragma unroll
                for (int f = 0; f < F; ++f) v[idx][f] = C::zero();
#pragma unroll
            for (int k = 0; k < D; ++k) {      // k ascending: the same summation order as corner_val
                PT Lv[2][F], Pv[4][F];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint32_t pos[D], pl[D], ln[D];
                    int bb = 0;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        if (d == k) pos[d] = c.cell[d];
                        else { pos[d] = c.cell[d] + ((b >> bb) & 1); ++bb; }
                    }
                    idx_vm<D>(c.res, pos, pl, ln);
                    load_feats<F>(g + (uint64_t)pl[k] * c.n_feat + c.gfo, Pv[b], vec_ok);
                    if (b == 0) {
                        load_feats<F>(g + (uint64_t)ln[k] * c.n_feat + c.gfo, Lv[0], vec_ok);
                        load_feats<F>(g + (uint64_t)(ln[k] + 1u) * c.n_feat + c.gfo, Lv[1], vec_ok);
                    }
                }
#pragma unroll
                for (int idx = 0; idx < (1 << D); ++idx) {
                    const int a = (idx >> k) & 1;
                    int b = 0, bb = 0;
#pragma unroll
                    for (int d = 0; d < D; ++d)
                        if (d != k) { b |= ((idx >> d) & 1) << bb; ++bb; }
#pragma unroll
                    for (int f = 0; f < F; ++f) v[idx][f] = C::add(v[idx][f], C::from_f(C::to_f(Pv[b][f]) * C::to_f(Lv[a][f])));
                }
            }
            return;
        }
    }
    if (c.type == NR3D_LOD_CP) {
        PT Lv[D][2][F];
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
            for (int a = 0; a < 2; ++a) load_feats<F>(g + (uint64_t)idx_cp_line<D>(c.res, c.cell[k] + a, k) * c.n_feat + c.gfo, Lv[k][a], vec_ok);
#pragma unroll
        for (int idx = 0; idx < (1 << D); ++idx) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                float r = C::to_f(Lv[0][idx & 1][f]);
#pragma unroll
                for (int k = 1; k < D; ++k) r *= C::to_f(Lv[k][(idx >> k) & 1][f]);
                v[idx][f] = C::from_f(r);
            }
        }
        return;
    }
#pragma unroll
    for (int idx = 0; idx < (1 << D); ++idx) {
        uint32_t pos[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pos[d] = c.cell[d] + ((idx >> d) & 1);
        corner_val<D, F, PT>(c, g, pos, v[idx], vec_ok);
    }
}

// sum_f corner_value[f] * grad[f] for the level types with a second-order dL/dx (Dense / Hash / VM / VecZMatXoY).  VM keeps the
// un-rounded products of the reference (calc_dLdx_dim_vm_impl, lotd_cuda.h:887-918).
template <int D, int F, typename PT>
__device__ __forceinline__ float corner_dot(const Ctx<D>& c, const PT* __restrict__ g, const uint32_t* pos, const float* grad, bool vec_ok) {
    using C = Cvt<PT>;
    float s = 0.f;
    if (c.type == NR3D_LOD_VM) {
        if constexpr (D == 3) {
            uint32_t plx[D], lnx[D];
            idx_vm<D>(c.res, pos, plx, lnx);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                PT a[F], b[F];
                load_feats<F>(g + (uint64_t)plx[k] * c.n_feat + c.gfo, a, vec_ok);
                load_feats<F>(g + (uint64_t)lnx[k] * c.n_feat + c.gfo, b, vec_ok);
#pragma unroll
                for (int f = 0; f < F; ++f) s += C::to_f(a[f]) * C::to_f(b[f]) * grad[f];
            }
        }
    } else if (c.type == NR3D_LOD_VECZMATXOY) {
        if constexpr (D == 3) {
            PT a[F], b[F];
            load_feats<F>(g + (uint64_t)(c.res[2] + pos[1] + pos[0] * c.res[0]) * c.n_feat + c.gfo, a, vec_ok);
            load_feats<F>(g + (uint64_t)pos[2] * c.n_feat + c.gfo, b, vec_ok);
#pragma unroll
            for (int f = 0; f < F; ++f) s += C::to_f(a[f]) * C::to_f(b[f]) * grad[f];
        }
    } else {
        PT v[F];
        corner_val<D, F, PT>(c, g, pos, v, vec_ok);
#pragma unroll
        for (int f = 0; f < F; ++f) s += C::to_f(v[f]) * grad[f];
    }
    return s;
}

// ------------------------------------------------------------------------------------------------
// gradient scatter
// ------------------------------------------------------------------------------------------------
template <int F>
__device__ __forceinline__ void scatter_add(float* addr, const float* v, bool vec_ok) {
    if (vec_ok) {
#pragma unroll
        for (int f = 0; f < F; f += 2) red_add_v2_f32(addr + f, v[f], v[f + 1]);
    } else {
#pragma unroll
        for (int f = 0; f < F; ++f) red_add_f32(addr + f, v[f]);
    }
}
template <int F>
__device__ __forceinline__ void scatter_add(__half* addr, const float* v, bool vec_ok) {
    if (vec_ok) {
#pragma unroll
        for (int f = 0; f < F; f += 2) red_add_h2(addr + f, __halves2half2(__float2half_rn(v[f]), __float2half_rn(v[f + 1])));
    } else {  // possibly odd element offset: one half atomic per feature (same rounding per element as the packed reduction)
#pragma unroll
        for (int f = 0; f < F; ++f) atomicAdd(addr + f, __float2half_rn(v[f]));
    }
}

// d(corner value)/d(params) * (grad * weight) for the n-linear types (reference: add_grid_gridient_*_impl)
template <int D, int F, typename PT>
__device__ __forceinline__ void corner_add_grad(const Ctx<D>& c, const PT* __restrict__ g, PT* __restrict__ gg,
                                                const uint32_t* pos, const float* grad, float w, bool vec_ok) {
    using C = Cvt<PT>;
    float wg[F];
#pragma unroll
    for (int f = 0; f < F; ++f) wg[f] = grad[f] * w;
    switch (c.type) {
    case NR3D_LOD_DENSE:
        scatter_add<F>(gg + (uint64_t)idx_dense<D>(c.res, pos) * c.n_feat + c.gfo, wg, vec_ok);
        break;
    case NR3D_LOD_HASH:
        scatter_add<F>(gg + (uint64_t)idx_hash<D>(pos, c.size) * c.n_feat + c.gfo, wg, vec_ok);
        break;
    case NR3D_LOD_VM:
        if constexpr (D == 3) {
            uint32_t pl[D], ln[D];
            idx_vm<D>(c.res, pos, pl, ln);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const uint64_t ip = (uint64_t)pl[k] * c.n_feat + c.gfo, il = (uint64_t)ln[k] * c.n_feat + c.gfo;
                PT a[F], b[F];
                load_feats<F>(g + ip, a, vec_ok);
                load_feats<F>(g + il, b, vec_ok);
                float ga[F], gb[F];
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    ga[f] = wg[f] * C::to_f(b[f]);
                    gb[f] = wg[f] * C::to_f(a[f]);
                }
                scatter_add<F>(gg + ip, ga, vec_ok);
                scatter_add<F>(gg + il, gb, vec_ok);
            }
        }
        break;
    case NR3D_LOD_VECZMATXOY:
        if constexpr (D == 3) {
            const uint64_t il = (uint64_t)pos[2] * c.n_feat + c.gfo;
            const uint64_t ip = (uint64_t)(c.res[2] + pos[1] + pos[0] * c.res[0]) * c.n_feat + c.gfo;
            PT a[F], b[F];
            load_feats<F>(g + ip, a, vec_ok);
            load_feats<F>(g + il, b, vec_ok);
            float ga[F], gb[F];
#pragma unroll
            for (int f = 0; f < F; ++f) {
                ga[f] = wg[f] * C::to_f(b[f]);
                gb[f] = wg[f] * C::to_f(a[f]);
            }
            scatter_add<F>(gg + ip, ga, vec_ok);
            scatter_add<F>(gg + il, gb, vec_ok);
        }
        break;
    case NR3D_LOD_NPLANEMUL:
    case NR3D_LOD_CP: {
        uint64_t id[D];
        PT a[D][F];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const uint32_t ii = (c.type == NR3D_LOD_CP) ? idx_cp_line<D>(c.res, pos[k], k) : idx_nplane<D>(c.res, pos, k);
            id[k] = (uint64_t)ii * c.n_feat + c.gfo;
            load_feats<F>(g + id[k], a[k], vec_ok);
        }
#pragma unroll
        for (int gd = 0; gd < D; ++gd) {
            float gp[F];
#pragma unroll
            for (int f = 0; f < F; ++f) {
                float cur = wg[f];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k != gd) cur *= C::to_f(a[k][f]);
                gp[f] = cur;
            }
            scatter_add<F>(gg + id[gd], gp, vec_ok);
        }
    } break;
    default:
        break;
    }
}

// ------------------------------------------------------------------------------------------------
// Aggregated gradient scatter of one (point, pseudo level) for the n-linear level types, given the weight cw[c] of every lattice
// corner c (bit d of c set -> cell[d] + 1).  Same sums as corner_add_grad over all corners (reference add_grid_gridient_*_impl,
// lotd_cuda.h:494-829, called once per corner -- and once per corner per derivative dimension in the second-order pass,
// lotd_encoding.h:764-1041), but contributions that land on the SAME table entry are summed in registers first:
//   VM      a line entry is shared by the 4 corners with the same bit k, a plane entry by 2: 18 reductions / 18 loads instead of 48 / 96
//   CP      a line entry is shared by 2^(D-1) corners:                                      2 D reductions instead of D 2^D
//   NPlane  a plane entry is shared by 2 corners:                                           D 2^(D-1) instead of D 2^D
// which matters most exactly where the reference is slowest: the tiny line / plane tables are same-address reduction hot spots.
// ------------------------------------------------------------------------------------------------
// Where the summed contributions go.  GlobalSink: L2 reductions into the gradient table.  PrivSink: entries below `n_small` (the
// whole table of a small Dense / CP level, the line part of a VM level) are accumulated in the CTA's shared-memory copy of the
// table first -- these tiny sub-tables receive thousands of reductions per entry and serialise in the L2 slices otherwise.
template <int F, typename PT>
struct GlobalSink {
    PT* gg;            // level base inside the gradient array
    uint32_t n_feat, gfo;
    bool vec_ok;
    __device__ __forceinline__ void add(uint32_t entry, const float* v) const { scatter_add<F>(gg + (uint64_t)entry * n_feat + gfo, v, vec_ok); }
};
template <int F, typename PT>
struct PrivSink {
    GlobalSink<F, PT> g;
    float* sm;         // this scene's slice of the shared-memory table: [n_small][F]
    uint32_t n_small;
    __device__ __forceinline__ void add(uint32_t entry, const float* v) const {
        if (entry < n_small) {
#pragma unroll
            for (int f = 0; f < F; ++f) atomicAdd(sm + entry * F + f, v[f]);
        } else {
            g.add(entry, v);
        }
    }
};

template <int D, int F, typename PT, typename Sink>
__device__ __forceinline__ void nlinear_scatter(const Ctx<D>& c, const PT* __restrict__ g, PT* __restrict__ gg, const Sink& sink, const float* cw,
                                                const float* grad, bool vec_ok) {
    using C = Cvt<PT>;
    if (c.type == NR3D_LOD_DENSE || c.type == NR3D_LOD_HASH) {
#pragma unroll
        for (int idx = 0; idx < (1 << D); ++idx) {
            uint32_t pos[D];
#pragma unroll
            for (int d = 0; d < D; ++d) pos[d] = c.cell[d] + ((idx >> d) & 1);
            float wg[F];
#pragma unroll
            for (int f = 0; f < F; ++f) wg[f] = grad[f] * cw[idx];
            sink.add(c.type == NR3D_LOD_DENSE ? idx_dense<D>(c.res, pos) : idx_hash<D>(pos, c.size), wg);
        }
        return;
    }
    if (c.type == NR3D_LOD_VM) {
        if constexpr (D == 3) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                // entries of line k (2) and plane k (4; plane k ignores dimension k)
                uint32_t el[2], ep[4];
                PT Lv[2][F], Pv[4][F];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    // corner with bit k = 0 and the two other bits taken from b (in increasing dimension order)
                    uint32_t pos[D], pl[D], ln[D];
                    int bb = 0;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        if (d == k) pos[d] = c.cell[d];
                        else { pos[d] = c.cell[d] + ((b >> bb) & 1); ++bb; }
                    }
                    idx_vm<D>(c.res, pos, pl, ln);
                    ep[b] = pl[k];
                    load_feats<F>(g + (uint64_t)ep[b] * c.n_feat + c.gfo, Pv[b], vec_ok);
                    if (b == 0) { el[0] = ln[k]; el[1] = ln[k] + 1u; }
                }
                load_feats<F>(g + (uint64_t)el[0] * c.n_feat + c.gfo, Lv[0], vec_ok);
                load_feats<F>(g + (uint64_t)el[1] * c.n_feat + c.gfo, Lv[1], vec_ok);
                float gl[2][F], gp[4][F];
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    gl[0][f] = gl[1][f] = 0.f;
#pragma unroll
                    for (int b = 0; b < 4; ++b) gp[b][f] = 0.f;
                }
#pragma unroll
                for (int idx = 0; idx < (1 << D); ++idx) {
                    const int a = (idx >> k) & 1;
                    int b = 0, bb = 0;
#pragma unroll
                    for (int d = 0; d < D; ++d)
                        if (d != k) { b |= ((idx >> d) & 1) << bb; ++bb; }
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        const float t = cw[idx] * grad[f];
                        gl[a][f] += t * C::to_f(Pv[b][f]);
                        gp[b][f] += t * C::to_f(Lv[a][f]);
                    }
                }
                sink.add(el[0], gl[0]);
                sink.add(el[1], gl[1]);
#pragma unroll
                for (int b = 0; b < 4; ++b) sink.add(ep[b], gp[b]);
            }
        }
        return;
    }
    if (c.type == NR3D_LOD_CP) {
        uint32_t el[D][2];
        PT Lv[D][2][F];
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                el[k][a] = idx_cp_line<D>(c.res, c.cell[k] + a, k);
                load_feats<F>(g + (uint64_t)el[k][a] * c.n_feat + c.gfo, Lv[k][a], vec_ok);
            }
        float acc[D][2][F];
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
            for (int f = 0; f < F; ++f) acc[k][0][f] = acc[k][1][f] = 0.f;
#pragma unroll
        for (int idx = 0; idx < (1 << D); ++idx) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    float t = cw[idx] * grad[f];
#pragma unroll
                    for (int k2 = 0; k2 < D; ++k2)
                        if (k2 != k) t *= C::to_f(Lv[k2][(idx >> k2) & 1][f]);
                    acc[k][(idx >> k) & 1][f] += t;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            sink.add(el[k][0], acc[k][0]);
            sink.add(el[k][1], acc[k][1]);
        }
        return;
    }
    // NPlaneMul, VecZMatXoY (rare) and anything else: per corner, straight to the table
#pragma unroll 1
    for (int idx = 0; idx < (1 << D); ++idx) {
        uint32_t pos[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pos[d] = c.cell[d] + ((idx >> d) & 1);
        corner_add_grad<D, F, PT>(c, g, gg, pos, grad, cw[idx], vec_ok);
    }
}

// number of leading table entries of a level that are worth privatising in shared memory (0: none)
__host__ __device__ inline uint32_t small_entries(uint32_t type, const uint32_t* res, int D, uint32_t size) {
    uint32_t lines = 0;
    for (int d = 0; d < D; ++d) lines += res[d];
    if (type == NR3D_LOD_CP) return lines;          // the whole table
    if (type == NR3D_LOD_VM) return lines;          // the lines; the planes follow
    if (type == NR3D_LOD_DENSE) return size;        // the whole table (the launcher checks that it fits)
    return 0;
}

// corner position for corner id `idx` (bit d set -> cell+1) and its n-linear weight
template <int D>
__device__ __forceinline__ float corner_weight(const Ctx<D>& c, int idx, uint32_t* pos) {
    float w = 1.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        if ((idx & (1 << d)) == 0) {
            w *= 1.0f - c.p[d];
            pos[d] = c.cell[d];
        } else {
            w *= c.p[d];
            pos[d] = c.cell[d] + 1;
        }
    }
    return w;
}
// weight for the (D-1)-face corner `idx` (bits over the dims != g), starting from w0; sets pos for dims != g
template <int D>
__device__ __forceinline__ float face_weight(const Ctx<D>& c, int g, int idx, float w0, uint32_t* pos, int* left_idx) {
    float w = w0;
    int li = 0;
#pragma unroll
    for (int ng = 0; ng < D - 1; ++ng) {
        const int dim = ng >= g ? ng + 1 : ng;
        if ((idx & (1 << ng)) == 0) {
            w *= 1.0f - c.p[dim];
            pos[dim] = c.cell[dim];
        } else {
            w *= c.p[dim];
            pos[dim] = c.cell[dim] + 1;
            li += 1 << dim;
        }
    }
    *left_idx = li;
    return w;
}

__device__ __forceinline__ bool is_nlinear(uint32_t type) { return type != NR3D_LOD_NPLANESUM && type != NR3D_LOD_CPFAST; }

}  // namespace nr3d
