// pipeline_ops.cu -- the streaming glue between the marcher, the encoder and the compositor.
//
// The reference strings these together from torch element-wise / index / reduce kernels, one full pass over the samples
// each (graphics/raymarch/occgrid_raymarch.py:92-110: index_select x2, addcmul, sub; graphics/nerf/nerf_ray_query.py:182:
// alpha = 1 - exp(-sigma * deltas)).  On B200 those passes cost as much as the encoder itself once it runs at its own
// roofline (profiles/r1_s2_m2_launches.txt: 25 % of the M2 step), so here each group is one kernel:
//   * march_samples:       (ridx, t_starts, t_ends, rays) -> sample positions + interval lengths, one pass;
//   * density_alpha_fwd:   features [S, C] -> sigma = softplus(gain * sum_c h) -> alpha = 1 - exp(-sigma * delta)
//                          (the density head that stands in for the decoder in the M2 workload);
//   * density_alpha_bwd:   dL/dalpha -> dL/dh [S, C] (one coalesced row store per sample).
// Bound: HBM streaming (fwd 4C + 12 B / sample, bwd 4C + 16 B / sample).
#include "common.cuh"
#include <algorithm>

namespace nr3d {

__global__ void __launch_bounds__(256)
march_samples_kernel(uint64_t S, const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ t_starts,
                     const float* __restrict__ t_ends, const int32_t* __restrict__ ridx32, const int64_t* __restrict__ ridx64,
                     float* __restrict__ samples, float* __restrict__ deltas) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const int64_t r = ridx32 ? (int64_t)ridx32[i] : ridx64[i];
    const float t0 = __ldcs(t_starts + i);
    if (deltas) __stcs(deltas + i, __ldcs(t_ends + i) - t0);
    // == torch.addcmul(rays_o[ridx], rays_d[ridx], t_starts): with value = 1 ATen's CUDA functor is a + b * c, which nvcc contracts
    // into one FMA -- reproduced so that the sample positions are bit-identical to the reference's (checked on B200)
    const float ox = __ldg(rays_o + r * 3), oy = __ldg(rays_o + r * 3 + 1), oz = __ldg(rays_o + r * 3 + 2);
    const float dx = __ldg(rays_d + r * 3), dy = __ldg(rays_d + r * 3 + 1), dz = __ldg(rays_d + r * 3 + 2);
    __stcs(samples + i * 3 + 0, fmaf(dx, t0, ox));
    __stcs(samples + i * 3 + 1, fmaf(dy, t0, oy));
    __stcs(samples + i * 3 + 2, fmaf(dz, t0, oz));
}

__device__ __forceinline__ float softplus1(float v) {  // == F.softplus(v) (beta 1, threshold 20)
    return v > 20.0f ? v : log1pf(expf(v));
}

// G lanes per row (G = 8 for C = 32: each lane one float4); a warp pass covers kU * 32 / G rows with kU independent loads per lane in
// flight (the kernel is a pure HBM stream: memory-level parallelism is what it needs).
constexpr int kU = 4;

template <int G, bool VEC4>
__global__ void __launch_bounds__(256)
density_alpha_fwd_kernel(uint64_t S, uint32_t C, const float* __restrict__ h, int64_t h_stride, const float* __restrict__ deltas, float gain,
                         float* __restrict__ sigma, float* __restrict__ alpha) {
    const uint32_t lane = threadIdx.x & 31, sub = lane % G, rl = lane / G;
    constexpr uint32_t kRows = 32 / G;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t row0 = warp * (kRows * kU); row0 < S; row0 += n_warps * (kRows * kU)) {
        float s[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const uint64_t row = row0 + (uint64_t)u * kRows + rl;
            s[u] = 0.f;
            if (row < S) {
                const float* hr = h + (int64_t)row * h_stride;
                if (VEC4) {
                    for (uint32_t c = sub * 4; c < C; c += G * 4) {
                        const float4 v = __ldcs(reinterpret_cast<const float4*>(hr + c));
                        s[u] += (v.x + v.y) + (v.z + v.w);
                    }
                } else {
                    for (uint32_t c = sub; c < C; c += G) s[u] += __ldcs(hr + c);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
#pragma unroll
            for (int m = G / 2; m > 0; m >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], m);
        }
        // lane `sub` < kU of every group finishes row u = sub of its group (spreads the transcendental work over the lanes)
        float mine = s[0];
#pragma unroll
        for (int u = 1; u < kU; ++u) mine = (sub == (uint32_t)u) ? s[u] : mine;
        const uint64_t row = row0 + (uint64_t)sub * kRows + rl;
        if (sub < (uint32_t)kU && row < S) {
            const float sg = softplus1(mine * gain);
            __stcs(sigma + row, sg);
            __stcs(alpha + row, 1.0f - expf(-sg * __ldcs(deltas + row)));
        }
    }
}

template <int G, bool VEC4>
__global__ void __launch_bounds__(256)
density_alpha_bwd_kernel(uint64_t S, uint32_t C, const float* __restrict__ d_alpha, const float* __restrict__ d_sigma_extra,
                         const float* __restrict__ sigma, const float* __restrict__ alpha, const float* __restrict__ deltas, float gain,
                         float* __restrict__ d_h) {
    const uint32_t lane = threadIdx.x & 31, sub = lane % G, rl = lane / G;
    constexpr uint32_t kRows = 32 / G;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t row0 = warp * (kRows * kU); row0 < S; row0 += n_warps * (kRows * kU)) {
        // lane `sub` < kU of every group computes the coefficient of row u = sub of its group, then the group shares them
        float mine = 0.f;
        {
            const uint64_t row = row0 + (uint64_t)sub * kRows + rl;
            if (sub < (uint32_t)kU && row < S) {
                // alpha = 1 - exp(-sigma * delta):  d alpha / d sigma = delta * (1 - alpha)
                // sigma = softplus(gain * s):       d sigma / d s     = gain * sigmoid(gain * s) = gain * (1 - exp(-sigma))
                const float sg = __ldcs(sigma + row);
                float g_sigma = __ldcs(d_alpha + row) * __ldcs(deltas + row) * (1.0f - __ldcs(alpha + row));
                if (d_sigma_extra) g_sigma += __ldcs(d_sigma_extra + row);
                const float sig = sg > 20.0f ? 1.0f : (1.0f - expf(-sg));
                mine = g_sigma * sig * gain;
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const float c = __shfl_sync(0xffffffffu, mine, (int)(rl * G + u));
            const uint64_t row = row0 + (uint64_t)u * kRows + rl;
            if (row >= S) continue;
            float* dr = d_h + row * (uint64_t)C;
            if (VEC4) {
                for (uint32_t k = sub * 4; k < C; k += G * 4) __stcs(reinterpret_cast<float4*>(dr + k), make_float4(c, c, c, c));
            } else {
                for (uint32_t k = sub; k < C; k += G) __stcs(dr + k, c);
            }
        }
    }
}

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_march_samples(uint64_t S, const float* rays_o, const float* rays_d, const float* t_starts, const float* t_ends,
                       const void* ridx, int32_t ridx_dtype, float* samples, float* deltas, void* stream) {
    if (S == 0) return 0;
    NR3D_CHECK(rays_o && rays_d && t_starts && ridx && samples, "march_samples: null argument");
    NR3D_CHECK(deltas == nullptr || t_ends != nullptr, "march_samples: deltas requested without t_ends");
    NR3D_CHECK(ridx_dtype == NR3D_I32 || ridx_dtype == NR3D_I64, "march_samples: ridx must be int32 or int64");
    march_samples_kernel<<<(unsigned)div_up<uint64_t>(S, 256), 256, 0, (cudaStream_t)stream>>>(
        S, rays_o, rays_d, t_starts, t_ends, ridx_dtype == NR3D_I32 ? (const int32_t*)ridx : nullptr,
        ridx_dtype == NR3D_I64 ? (const int64_t*)ridx : nullptr, samples, deltas);
    NR3D_LAUNCH_CHECK("march_samples");
    return 0;
}

int nr3d_density_alpha_fwd(uint64_t S, uint32_t C, const float* h, int64_t h_stride, const float* deltas, float gain, float* sigma,
                           float* alpha, void* stream) {
    if (S == 0) return 0;
    NR3D_CHECK(h && deltas && sigma && alpha && C > 0, "density_alpha_fwd: null argument");
    const bool vec4 = (C % 4 == 0) && (h_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(h) & 15u) == 0);
    const unsigned grid = (unsigned)std::min<uint64_t>(div_up<uint64_t>(S, 4 * kU * 8), (uint64_t)kSMs * 8);   // 8 warps x 16 rows per CTA pass
    if (vec4) density_alpha_fwd_kernel<8, true><<<grid, 256, 0, (cudaStream_t)stream>>>(S, C, h, h_stride, deltas, gain, sigma, alpha);
    else density_alpha_fwd_kernel<8, false><<<grid, 256, 0, (cudaStream_t)stream>>>(S, C, h, h_stride, deltas, gain, sigma, alpha);
    NR3D_LAUNCH_CHECK("density_alpha_fwd");
    return 0;
}

int nr3d_density_alpha_bwd(uint64_t S, uint32_t C, const float* d_alpha, const float* d_sigma_extra, const float* sigma, const float* alpha,
                           const float* deltas, float gain, float* d_h, void* stream) {
    if (S == 0) return 0;
    NR3D_CHECK(d_alpha && sigma && alpha && deltas && d_h && C > 0, "density_alpha_bwd: null argument");
    const bool vec4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_h) & 15u) == 0);
    const unsigned grid = (unsigned)std::min<uint64_t>(div_up<uint64_t>(S, 4 * kU * 8), (uint64_t)kSMs * 8);
    if (vec4) density_alpha_bwd_kernel<8, true><<<grid, 256, 0, (cudaStream_t)stream>>>(S, C, d_alpha, d_sigma_extra, sigma, alpha, deltas, gain, d_h);
    else density_alpha_bwd_kernel<8, false><<<grid, 256, 0, (cudaStream_t)stream>>>(S, C, d_alpha, d_sigma_extra, sigma, alpha, deltas, gain, d_h);
    NR3D_LAUNCH_CHECK("density_alpha_bwd");
    return 0;
}

}  // extern "C"
