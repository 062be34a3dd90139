// lotd_fused_bwd.cu -- backward of the fused LoTD encoder + density decoder (forward: lotd_fused.cu).
//
// Reference behaviour: autograd through LoTDNeRF.forward_density (nr3d_lib/models/fields/nerf/lotd_nerf.py:136-178):
//     h = encoding(x);  h1 = relu(h W1^T + b1);  out = h1 W2^T + b2;  sigma = activation(out[..., 0])
// i.e. per training step the [N, 32] features, the [N, 64] hidden activations and both their gradients cross HBM (at least
// 2 x (128 + 256) B per sample) around four GEMM launches and the LoTD scatter.  Here one persistent kernel does, per 128-point tile:
//   1. encode     re-gather the features (the forward keeps nothing but sigma) -> F  [128 x 32] bf16 (+ a column of ones), shared memory
//                 load the upstream gradient rows                               -> G  [128 x 16] bf16 (dL/dout, activation folded in)
//   2. MMA-A      D1 = F W1^T (TMEM)               epilogue A: h1 = relu(D1 + b1) -> H [128 x 64] bf16, ReLU mask kept in a register
//   3. MMA-B      dH = G W2 (TMEM)                 and  dW2^T [64 x 16] += H^T G          (contraction over the tile's points)
//                 epilogue B: dh = mask ? dH : 0 -> DH [128 x 64] bf16 (overwrites H)
//   4. MMA-C      dF = DH W1 (TMEM)                and  [dW1 | db1] [64 x 40] += DH^T [F | 1]
//                 epilogue C: dF -> shared memory (fp32)
//   5. scatter    dL/dparam += corner weights x dF with the run-merged two-lanes-per-point scatter of lotd_fast.cu (red.global.add.v2.f32)
// All MMAs are tcgen05.mma.kind::f16 (bf16 operands, fp32 accumulators in TMEM) issued by one thread; the point contractions read the SAME
// shared-memory tiles as MN-major operands (lotd_umma.cuh), so no transposed copy is made.  The weight-gradient accumulators live in TMEM
// for the whole kernel (M = 64 tiles: rows 16q + r sit in TMEM lane 32q + r) and reach HBM once per CTA.
// Precision: operands are rounded to bf16 (like the forward and like tcnn's fp16 MLPs); gradients match an fp32 autograd reference
// to ~1e-2 relative to the tensor's largest entry (tests/test_fused_gpu.py).  dL/dx is not produced (positions are not trained in the
// reference's density path; NeuS-style callers use the unfused operators).
#include "lotd_umma.cuh"

namespace nr3d {

constexpr int kBwThreads = 256;            // 128 points per tile, two lanes per point
constexpr uint32_t kBwTmemCols = 128;      // [0, 64) D1 -> dH -> dF;  [64, 104) dW1 | db1;  [104, 120) dW2^T
constexpr uint32_t kColW1 = 64, kColW2 = 104;
constexpr int kBwCtasPerSm = 4;            // 4 x 128 TMEM columns = the SM's 512
constexpr int kDfStride = 33;              // floats per staged dF row (conflict-free for row-per-lane writes and pair reads)
// shared-memory map (bytes); tiles are [k-block of 8 channels][row-block of 8][8 rows][16 B]
constexpr uint32_t kBoF = 0;                         // F tile: 5 channel blocks x 2048 (features 0..31, block 4 = ones column + zeros)
constexpr uint32_t kBoG = kBoF + 5 * 2048;           // G tile: 2 channel blocks x 2048
constexpr uint32_t kBoR0End = 128 * kDfStride * 4;   // ... the region [0, 16896) is reused for the fp32 dF rows
constexpr uint32_t kBoH = 16896;                     // H / DH tile: 8 x 2048
constexpr uint32_t kBoW1 = kBoH + 16384;             // W1  [64 n][32 k]  K-major  (layer 1, B operand)           4096
constexpr uint32_t kBoW2t = kBoW1 + 4096;            // W2^T [64 n][16 k] K-major  (dH = G W2, B operand)         2048
constexpr uint32_t kBoW1t = kBoW2t + 2048;           // W1^T [32 n][64 k] K-major  (dF = DH W1, B operand)        4096
constexpr uint32_t kBoB1 = kBoW1t + 4096;            // 64 f32
constexpr uint32_t kBoIdx = kBoB1 + 256;             // 128 u32 original index per row
constexpr uint32_t kBoBar = kBoIdx + 512;            // 1 mbarrier
constexpr uint32_t kBoTmem = kBoBar + 16;
constexpr uint32_t kBwSmem = kBoTmem + 16;
static_assert(kBoG + 2 * 2048 <= kBoR0End && kBoR0End <= kBoH, "shared-memory map");

struct FusedBwd {
    const uint4* w1c;    // W1   [64, 32] bf16, K-major core-matrix order
    const uint4* w2tc;   // W2^T [64, 16] bf16
    const uint4* w1tc;   // W1^T [32, 64] bf16
    const float* b1;     // [64] or null
    const float* sigma;    // [N] forward output (activation derivative), may be null when d_sigma is null
    const float* d_sigma;  // [N] dL/dsigma or null
    const float* d_out16;  // [N, 16] dL/d(decoder output) or null (added to the d_sigma term in column 0)
    int32_t activation;    // 0 identity, 1 exp, 2 softplus, 3 relu
    float* dparams;      // [n_params] f32, accumulated
    float* dw1;          // [64, 32] f32, accumulated
    float* db1;          // [64]
    float* dw2;          // [16, 64]
    float* db2;          // [16]
};

// run-merged scatter of one pseudo level (F = 2, fp32 tables): the backward loop body of lotd_pair_bwd_kernel (lotd_fast.cu)
__device__ __forceinline__ void scatter_level(const LevelDesc& L, const FastLevel& X, uint32_t gfo, bool smooth, float x, float y, float z, uint32_t side, int k,
                                              bool live, float g0, float g1, float* __restrict__ grad) {
    Geo2 g;
    pair_geo(L, X, gfo, 0u, smooth, x, y, z, side, g);
    float cx[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) { cx[q][0] = g.w[q] * g0; cx[q][1] = g.w[q] * g1; }
    bool issue = live;
    if (L.res[0] <= 1024u && L.res[1] <= 1024u && L.res[2] <= 1024u) {
        const uint32_t key = live ? g.key : (0xffffffffu - (uint32_t)k);
        const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 2);
        const uint32_t hmask = __ballot_sync(0xffffffffu, k == 0 || key != prev) & 0x55555555u;
        if (hmask != 0x55555555u) {
            const uint32_t le = hmask & (0xffffffffu >> (31 - 2 * k));
            const int s0 = (31 - __clz(le)) >> 1;
            const uint32_t above = hmask & (0xffffffffu << (2 * k + 1));
            const int e0 = above ? ((__ffs(above) - 1) >> 1) : 16;
            const int r = e0 - s0, j = k - s0;
            const int rmax = __reduce_max_sync(0xffffffffu, r);
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
                if (d >= rmax) break;
                const bool take = j + d < r;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        const float u = __shfl_down_sync(0xffffffffu, cx[q][f], 2 * d);
                        if (take) cx[q][f] += u;
                    }
            }
            issue = live && j == 0;
        }
    }
    if (issue) {
#pragma unroll
        for (int q = 0; q < 4; ++q) red_add_v2_f32(grad + g.e[q], cx[q][0], cx[q][1]);
    }
}

// OUT16: an upstream gradient for all 16 decoder outputs is given (a.d_out16); otherwise only dL/dsigma feeds column 0 and the bias
// gradient db2 needs one register instead of eight (the kernel is compiled for 64 registers per thread: 4 CTAs / SM).
template <bool OUT16>
__global__ void __launch_bounds__(kBwThreads, kBwCtasPerSm)
lotd_fused_density_bwd_kernel(const __grid_constant__ LotdTable tab, const FastIn in, const FusedBwd a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t side = lane & 1;
    const int k = lane >> 1;
    const int m = tid >> 1;   // row of this thread pair's point inside the tile
    uint32_t* idx_s = reinterpret_cast<uint32_t*>(smem + kBoIdx);
    float* df_s = reinterpret_cast<float*>(smem);
    const uint32_t bar = smem_u32(smem + kBoBar);

    // ---- one-time setup ----
    for (int i = tid; i < (4096 + 2048 + 4096) / 16; i += kBwThreads) {
        uint4 v;
        if (i < 256) v = __ldg(a.w1c + i);
        else if (i < 384) v = __ldg(a.w2tc + (i - 256));
        else v = __ldg(a.w1tc + (i - 384));
        reinterpret_cast<uint4*>(smem + kBoW1)[i] = v;
    }
    if (tid < 64) reinterpret_cast<float*>(smem + kBoB1)[tid] = a.b1 ? __ldg(a.b1 + tid) : 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kBoTmem);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kBwTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint64_t n_tiles = (in.N + 127) / 128;
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const float* params = reinterpret_cast<const float*>(in.params);
    uint32_t phase = 0;        // parity of the next barrier completion
    bool first_tile = true;
    constexpr int NB2 = OUT16 ? 8 : 1;
    float db2_acc[NB2];        // this lane's dL/dout columns (side * 8 + j), summed over the CTA's tiles
#pragma unroll
    for (int j = 0; j < NB2; ++j) db2_acc[j] = 0.f;
    const uint32_t quad = warp & 3, half = warp >> 2;     // TMEM lanes 32 quad .. +31, column half
    const uint32_t erow = quad * 32 + lane;               // tile row this thread owns in the epilogues

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t p = tile * 128 + m;
        const bool active = p < in.N;
        // ---- 1. encode + upstream gradient rows ----
        float4 rec = make_float4(0.5f, 0.5f, 0.5f, 0.f);
        if (active) rec = __ldcs(in.xs + p);
        const uint32_t i = __float_as_uint(rec.w);
        if (side == 0) idx_s[m] = active ? i : 0xffffffffu;
        uint8_t* f_row = smem + kBoF + (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll 2
        for (uint32_t pl = 0; pl < 16; ++pl) {
            const uint32_t level = tab.map_level[pl];
            float r0 = 0.f, r1 = 0.f;
            if (active && (int32_t)level <= in.max_level) {
                Geo2 g;
                pair_geo(tab.lv[level], in.fl[level], (uint32_t)tab.map_cnt[pl] * 2u, 0u, smooth, rec.x, rec.y, rec.z, side, g);
                float2 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float2*>(params + g.e[q]));
#pragma unroll
                for (int q = 0; q < 4; ++q) { r0 += g.w[q] * v[q].x; r1 += g.w[q] * v[q].y; }
            }
            r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
            r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
            if (side == 0) *reinterpret_cast<uint32_t*>(f_row + (pl >> 2) * 2048 + (pl & 3) * 4) = pack_bf16(r0, r1);
        }
        {
            // channel block 4 of the F tile: (1, 0, 0, 0, 0, 0, 0, 0) per row -- the ones column that turns dW1's MMA into [dW1 | db1]
            if (side == 1) *reinterpret_cast<uint4*>(f_row + 4 * 2048) = make_uint4(active ? 0x00003f80u : 0u, 0u, 0u, 0u);
            // G row: lane `side` fills columns side * 8 .. + 7
            float gv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) gv[j] = 0.f;
            if (active) {
                if (OUT16) {
                    const float4* src = reinterpret_cast<const float4*>(a.d_out16 + (uint64_t)i * 16 + side * 8);
                    const float4 u0 = __ldcs(src), u1 = __ldcs(src + 1);
                    gv[0] = u0.x; gv[1] = u0.y; gv[2] = u0.z; gv[3] = u0.w; gv[4] = u1.x; gv[5] = u1.y; gv[6] = u1.z; gv[7] = u1.w;
                }
                if (side == 0 && a.d_sigma) {
                    const float ds = __ldcs(a.d_sigma + i);
                    float dact = 1.f;   // d activation / d out0 expressed through the forward's output sigma
                    if (a.activation != 0) {
                        const float sg = __ldcs(a.sigma + i);
                        dact = a.activation == 1 ? sg : (a.activation == 2 ? (sg > 20.f ? 1.f : 1.f - __expf(-sg)) : (sg > 0.f ? 1.f : 0.f));
                    }
                    gv[0] += ds * dact;
                }
            }
#pragma unroll
            for (int j = 0; j < NB2; ++j) db2_acc[j] += gv[j];
            *reinterpret_cast<uint4*>(smem + kBoG + side * 2048 + (m >> 3) * 128 + (m & 7) * 16) =
                make_uint4(pack_bf16(gv[0], gv[1]), pack_bf16(gv[2], gv[3]), pack_bf16(gv[4], gv[5]), pack_bf16(gv[6], gv[7]));
        }

        // ---- 2. MMA-A: D1 = F W1^T ----
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint64_t da = umma_desc(smem_u32(smem + kBoF), 2048, 128), db = umma_desc(smem_u32(smem + kBoW1), 1024, 128);
#pragma unroll
            for (uint32_t s = 0; s < 2; ++s)
                umma_f16(tmem, da + (uint64_t)((s * 2 * 2048) >> 4), db + (uint64_t)((s * 2 * 1024) >> 4), umma_idesc(128, 64), s);
            umma_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();
        // epilogue A: hidden activations (bf16) + ReLU mask of this thread's (row, 32 columns)
        uint32_t relu_mask = 0;
        {
            const float* b1 = reinterpret_cast<const float*>(smem + kBoB1) + half * 32;
            uint8_t* h_row = smem + kBoH + (erow >> 3) * 128 + (erow & 7) * 16 + (half * 4) * 2048;
#pragma unroll
            for (int c16 = 0; c16 < 2; ++c16) {
                uint32_t v[16];
                NR3D_TMEM_LD16(tmem + ((quad * 32) << 16) + half * 32 + c16 * 16, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float u0 = __uint_as_float(v[2 * j]) + b1[c16 * 16 + 2 * j], u1 = __uint_as_float(v[2 * j + 1]) + b1[c16 * 16 + 2 * j + 1];
                    relu_mask |= (u0 > 0.f ? 1u : 0u) << (c16 * 16 + 2 * j);
                    relu_mask |= (u1 > 0.f ? 1u : 0u) << (c16 * 16 + 2 * j + 1);
                    h[j] = pack_bf16(fmaxf(u0, 0.f), fmaxf(u1, 0.f));
                }
                *reinterpret_cast<uint4*>(h_row + (c16 * 2) * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(h_row + (c16 * 2 + 1) * 2048) = make_uint4(h[4], h[5], h[6], h[7]);
            }
        }

        // ---- 3. MMA-B: dH = G W2;  dW2^T += H^T G ----
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint32_t sG = smem_u32(smem + kBoG), sH = smem_u32(smem + kBoH);
            umma_f16(tmem, umma_desc(sG, 2048, 128), umma_desc(smem_u32(smem + kBoW2t), 1024, 128), umma_idesc(128, 64), 0);
            // contraction over the 128 points: eight K = 16 steps, both operands MN-major (LBO = 128: next 8 points, SBO = 2048: next 8 channels)
            const uint64_t da = umma_desc(sH, 128, 2048), db = umma_desc(sG, 128, 2048);
#pragma unroll
            for (uint32_t s = 0; s < 8; ++s)
                umma_f16(tmem + kColW2, da + (uint64_t)((s * 256) >> 4), db + (uint64_t)((s * 256) >> 4), umma_idesc(64, 16, 1, 1), (first_tile && s == 0) ? 0u : 1u);
            umma_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();
        // epilogue B: masked hidden gradient (bf16) over the hidden tile
        {
            uint8_t* h_row = smem + kBoH + (erow >> 3) * 128 + (erow & 7) * 16 + (half * 4) * 2048;
#pragma unroll
            for (int c16 = 0; c16 < 2; ++c16) {
                uint32_t v[16];
                NR3D_TMEM_LD16(tmem + ((quad * 32) << 16) + half * 32 + c16 * 16, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float u0 = ((relu_mask >> (c16 * 16 + 2 * j)) & 1u) ? __uint_as_float(v[2 * j]) : 0.f;
                    const float u1 = ((relu_mask >> (c16 * 16 + 2 * j + 1)) & 1u) ? __uint_as_float(v[2 * j + 1]) : 0.f;
                    h[j] = pack_bf16(u0, u1);
                }
                *reinterpret_cast<uint4*>(h_row + (c16 * 2) * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(h_row + (c16 * 2 + 1) * 2048) = make_uint4(h[4], h[5], h[6], h[7]);
            }
        }

        // ---- 4. MMA-C: dF = DH W1;  [dW1 | db1] += DH^T [F | 1] ----
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint32_t sH = smem_u32(smem + kBoH), sF = smem_u32(smem + kBoF);
            const uint64_t da = umma_desc(sH, 2048, 128), db = umma_desc(smem_u32(smem + kBoW1t), 512, 128);
#pragma unroll
            for (uint32_t s = 0; s < 4; ++s)   // K = 64 hidden channels
                umma_f16(tmem, da + (uint64_t)((s * 2 * 2048) >> 4), db + (uint64_t)((s * 2 * 512) >> 4), umma_idesc(128, 32), s);
            const uint64_t ta = umma_desc(sH, 128, 2048), tb = umma_desc(sF, 128, 2048);
#pragma unroll
            for (uint32_t s = 0; s < 8; ++s)   // K = 128 points
                umma_f16(tmem + kColW1, ta + (uint64_t)((s * 256) >> 4), tb + (uint64_t)((s * 256) >> 4), umma_idesc(64, 40, 1, 1), (first_tile && s == 0) ? 0u : 1u);
            umma_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1u;
        tc_fence_after();
        // epilogue C: dF rows (fp32) into shared memory; thread (quad, half) holds row erow, columns half * 16 .. + 15
        {
            uint32_t v[16];
            NR3D_TMEM_LD16(tmem + ((quad * 32) << 16) + half * 16, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) df_s[erow * kDfStride + half * 16 + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        // ---- 5. scatter ----
        {
            const bool live = active;
            const float x = live ? rec.x : 0.5f, y = live ? rec.y : 0.5f, z = live ? rec.z : 0.5f;
            for (uint32_t pl = 0; pl < 16; ++pl) {
                const uint32_t level = tab.map_level[pl];
                if ((int32_t)level > in.max_level) continue;   // uniform
                const float g0 = live ? df_s[m * kDfStride + 2 * pl] : 0.f, g1 = live ? df_s[m * kDfStride + 2 * pl + 1] : 0.f;
                scatter_level(tab.lv[level], in.fl[level], (uint32_t)tab.map_cnt[pl] * 2u, smooth, x, y, z, side, k, live, g0, g1, a.dparams);
            }
        }
        __syncthreads();   // the dF rows / idx are free for the next tile's F and G
        first_tile = false;
    }

    // ---- weight gradients: TMEM accumulators -> HBM, once per CTA ----
    if (!first_tile) {
        tc_fence_after();
        if (warp < 4) {
            // M = 64 accumulators: row 16 q + r lives in TMEM lane 32 q + r (r < 16)
            const uint32_t row = quad * 16 + lane;
            uint32_t v[16];
            // [dW1 | db1]: 40 columns
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                NR3D_TMEM_LD16(tmem + ((quad * 32) << 16) + kColW1 + c * 16, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (lane < 16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(a.dw1 + row * 32 + c * 16 + j, __uint_as_float(v[j]));
                }
            }
            {
                uint32_t u[8];
                NR3D_TMEM_LD8(tmem + ((quad * 32) << 16) + kColW1 + 32, u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (lane < 16) atomicAdd(a.db1 + row, __uint_as_float(u[0]));
            }
            NR3D_TMEM_LD16(tmem + ((quad * 32) << 16) + kColW2, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(a.dw2 + j * 64 + row, __uint_as_float(v[j]));
            }
        }
        // db2: sum this lane's eight columns over the lanes of the same side, then over the CTA through HBM atomics
#pragma unroll
        for (int j = 0; j < NB2; ++j) {
            float s = db2_acc[j];
#pragma unroll
            for (int mm = 2; mm < 32; mm <<= 1) s += __shfl_xor_sync(0xffffffffu, s, mm);
            if (lane < 2 && (OUT16 || side == 0)) atomicAdd(a.db2 + side * 8 + j, s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kBwTmemCols) : "memory");
}

int make_table_public(const nr3d_lotd_meta* m, LotdTable& tab);  // lotd_fast.cu

}  // namespace nr3d

using namespace nr3d;

extern "C" int nr3d_lotd_fused_density_bwd(const nr3d_lotd_meta* meta, uint64_t N, const void* xs, const void* params, int32_t max_level,
                                           const void* w1_packed, const void* w2t_packed, const void* w1t_packed, const float* b1, int32_t activation,
                                           const float* sigma, const float* d_sigma, const float* d_out16, float* dL_dparam, float* dW1, float* db1,
                                           float* dW2, float* db2, void* stream) {
    NR3D_CHECK(meta != nullptr, "fused_density_bwd: null meta");
    NR3D_CHECK(meta->hash_only && meta->n_dims_to_encode == 3 && meta->n_feat_per_pseudo_lvl == 2 && meta->n_pseudo_levels == 16,
               "fused_density_bwd: needs a Dense/Hash-only meta with D=3, F=2 and 16 pseudo levels (32 features)");
    NR3D_CHECK(N < (1ull << 32) - 1, "fused_density_bwd: N must be < 2^32 - 1");
    NR3D_CHECK(activation >= 0 && activation <= 3, "fused_density_bwd: activation code %d not in [0, 3]", (int)activation);
    if (N == 0) return 0;
    NR3D_CHECK(xs && params && w1_packed && w2t_packed && w1t_packed && dL_dparam && dW1 && db1 && dW2 && db2, "fused_density_bwd: null argument");
    NR3D_CHECK(d_sigma || d_out16, "fused_density_bwd: neither d_sigma nor d_out16 given");
    NR3D_CHECK(!d_sigma || activation == 0 || sigma, "fused_density_bwd: d_sigma needs the forward's sigma for the activation derivative");
    NR3D_CHECK(((uintptr_t)w1_packed & 15) == 0 && ((uintptr_t)w2t_packed & 15) == 0 && ((uintptr_t)w1t_packed & 15) == 0 &&
               (!d_out16 || ((uintptr_t)d_out16 & 15) == 0) && ((uintptr_t)dL_dparam & 7) == 0,
               "fused_density_bwd: packed weights / d_out16 must be 16-byte aligned, dL_dparam 8-byte aligned");
    LotdTable tab;
    make_table_public(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), nullptr, params, max_level, meta->n_params, 0u, meta->n_pseudo_levels};
    fast_levels(meta, in);
    FusedBwd a{reinterpret_cast<const uint4*>(w1_packed), reinterpret_cast<const uint4*>(w2t_packed), reinterpret_cast<const uint4*>(w1t_packed), b1,
               sigma, d_sigma, d_out16, activation, dL_dparam, dW1, db1, dW2, db2};
    static bool configured = false;
    if (!configured) {
        NR3D_CHECK(cudaFuncSetAttribute(lotd_fused_density_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwSmem) == cudaSuccess &&
                   cudaFuncSetAttribute(lotd_fused_density_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwSmem) == cudaSuccess,
                   "fused_density_bwd: cannot reserve %d bytes of shared memory", (int)kBwSmem);
        configured = true;
    }
    const uint64_t n_tiles = div_up<uint64_t>(N, 128);
    const unsigned grid = (unsigned)(n_tiles < (uint64_t)kSMs * kBwCtasPerSm ? n_tiles : (uint64_t)kSMs * kBwCtasPerSm);
    if (d_out16) lotd_fused_density_bwd_kernel<true><<<grid, kBwThreads, kBwSmem, (cudaStream_t)stream>>>(tab, in, a);
    else lotd_fused_density_bwd_kernel<false><<<grid, kBwThreads, kBwSmem, (cudaStream_t)stream>>>(tab, in, a);
    NR3D_LAUNCH_CHECK("lotd_fused_density_bwd");
    return 0;
}
