// lotd_fused.cu -- LoTD encoder + density decoder fused in one kernel, forward only.
//
// Reference behaviour: LoTDNeRF.query_density / forward_density (nr3d_lib/models/fields/nerf/lotd_nerf.py:136-178):
//     h = encoding(x)  ->  density_decoder(h) = Linear(32, 64) -> ReLU -> Linear(64, 1 + n_extra)  ->  activation(out[..., 0])
// (models/blocks/mlp.py, D = 1 hidden layer, W = 64: the decoder of the NGP-style configs).  The reference materialises the
// [N, 32] feature tensor in HBM (128 B/sample written and read back) and runs the MLP as separate GEMM launches.
// Here the features of a CTA's 128 points never leave the SM:
//   1. encode: the "two lanes per point" gather loop of lotd_fast.cu writes bf16 features straight into the K-major
//      core-matrix layout that tcgen05.mma reads from shared memory (A1: 128 x 32);
//   2. layer 1: D1[128 x 64] (TMEM, fp32) = A1 . W1^T   -- two tcgen05.mma.kind::f16 (K = 16 each), one elected thread;
//   3. epilogue 1: all 8 warps pull their TMEM quadrant with tcgen05.ld, add bias, ReLU, round to bf16 and store the hidden
//      tile as the next A operand (A2: 128 x 64);
//   4. layer 2: D2[128 x 16] = A2 . W2^T (four MMAs), epilogue 2 applies the density activation and writes sigma (and,
//      optionally, all 16 decoder outputs) at the points' ORIGINAL indices.
// The kernel is persistent (6 CTAs per SM, each looping over 128-point tiles): tcgen05.alloc / dealloc serialise per SM, so
// paying them per tile would cost more than the whole encode (measured: 2.06 ms vs 0.65 ms for 4 Mi points).
// Operands are bf16 with fp32 accumulation: results match an fp32 MLP fed with bf16-rounded operands to ~1e-3 (test tolerance
// 2e-2 against the plain fp32 reference).  Completion is tracked with mbarriers (tcgen05.commit); every wait is bounded and
// traps instead of hanging.
#include "lotd_umma.cuh"

namespace nr3d {

constexpr int kFusedThreads = 256;  // 128 points per CTA
constexpr uint32_t kTmemCols = 64;  // D1: columns [0, 64); D2 reuses columns [0, 16) once epilogue 1 has drained D1
constexpr int kFusedCtasPerSm = 6;   // persistent CTAs: TMEM is allocated once per CTA, not once per 128 points
// shared-memory map (bytes); core matrix = 8 rows x 16 bytes, [k-block][row-block][8][16 B]
constexpr uint32_t kOffA1 = 0;                   // 4 k-blocks x 16 row-blocks x 128 B =  8192
constexpr uint32_t kOffA2 = kOffA1 + 8192;       // 8 x 16 x 128                      = 16384
constexpr uint32_t kOffW1 = kOffA2 + 16384;      // 4 k-blocks x 8 col-blocks x 128   =  4096
constexpr uint32_t kOffW2 = kOffW1 + 4096;       // 8 x 2 x 128                       =  2048
constexpr uint32_t kOffB1 = kOffW2 + 2048;       // 64 f32
constexpr uint32_t kOffB2 = kOffB1 + 256;        // 16 f32
constexpr uint32_t kOffIdx = kOffB2 + 64;        // 128 u32: original index per row (0xffffffff = padding row)
constexpr uint32_t kOffBar = kOffIdx + 512;      // 2 mbarriers
constexpr uint32_t kOffTmem = kOffBar + 16;      // TMEM base address
constexpr uint32_t kFusedSmem = kOffTmem + 16;

struct FusedDec {
    const uint4* w1c;   // W1 [64, 32] bf16 in core-matrix order (4096 B)
    const uint4* w2c;   // W2 [16, 64] bf16 in core-matrix order (2048 B)
    const float* b1;    // [64] or null
    const float* b2;    // [16] or null
    int32_t activation; // 0 identity, 1 exp, 2 softplus, 3 relu
};

__global__ void __launch_bounds__(kFusedThreads)
lotd_fused_density_kernel(const __grid_constant__ LotdTable tab, const FastIn in, const FusedDec dec, float* __restrict__ sigma,
                          float* __restrict__ out16) {
    __shared__ __align__(128) uint8_t smem[kFusedSmem];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t side = lane & 1;
    const int m = tid >> 1;  // row of this point inside the CTA's tile
    uint32_t* idx_s = reinterpret_cast<uint32_t*>(smem + kOffIdx);
    const uint32_t bar1 = smem_u32(smem + kOffBar), bar2 = bar1 + 8;

    // ---- one-time setup: weights, biases, barriers ----
    for (int i = tid; i < (4096 + 2048) / 16; i += kFusedThreads)
        reinterpret_cast<uint4*>(smem + kOffW1)[i] = i < 256 ? __ldg(dec.w1c + i) : __ldg(dec.w2c + (i - 256));
    if (tid < 64) reinterpret_cast<float*>(smem + kOffB1)[tid] = dec.b1 ? __ldg(dec.b1 + tid) : 0.f;
    if (tid < 16) reinterpret_cast<float*>(smem + kOffB2)[tid] = dec.b2 ? __ldg(dec.b2 + tid) : 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar2) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint64_t n_tiles = (in.N + 127) / 128;
    uint32_t parity = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, parity ^= 1u) {
    const uint64_t p = tile * 128 + m;
    const bool active = p < in.N;
    // ---- 1. encode: bf16 features into the A1 operand tile ----
    float4 rec = make_float4(0.5f, 0.5f, 0.5f, 0.f);
    if (active) rec = __ldcs(in.xs + p);
    if (side == 0) idx_s[m] = active ? __float_as_uint(rec.w) : 0xffffffffu;
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const float* params = reinterpret_cast<const float*>(in.params);
    uint8_t* a1_row = smem + kOffA1 + (m >> 3) * 128 + (m & 7) * 16;
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
        // feature columns (2 pl, 2 pl + 1): k-block pl / 4, byte (2 pl % 8) * 2 inside the row's 16 bytes
        if (side == 0) *reinterpret_cast<uint32_t*>(a1_row + (pl >> 2) * 2048 + (pl & 3) * 4) = pack_bf16(r0, r1);
    }

    // ---- 2. layer 1 on the tensor core ----
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint64_t da = umma_desc(smem_u32(smem + kOffA1), 2048, 128), db = umma_desc(smem_u32(smem + kOffW1), 1024, 128);
#pragma unroll
        for (uint32_t s = 0; s < 2; ++s)   // K = 32 = 2 x 16: each step advances two k-blocks
            umma_f16(tmem, da + (uint64_t)((s * 2 * 2048) >> 4), db + (uint64_t)((s * 2 * 1024) >> 4), umma_idesc(128, 64), s);
        umma_commit(bar1);
    }
    mbar_wait(bar1, parity);
    tc_fence_after();

    // ---- 3. epilogue 1: bias + ReLU -> bf16 hidden tile (A2) ----
    {
        const uint32_t quad = warp & 3, half = warp >> 2;           // TMEM lanes 32 quad .. +31, columns 32 half .. +31
        const uint32_t row = quad * 32 + lane;
        const float* b1 = reinterpret_cast<const float*>(smem + kOffB1) + half * 32;
        uint8_t* a2_row = smem + kOffA2 + (row >> 3) * 128 + (row & 7) * 16 + (half * 4) * 2048;
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
            uint32_t v[16];
            NR3D_TMEM_LD16(tmem + ((quad * 32) << 16) + half * 32 + c16 * 16, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                h[j] = pack_bf16(fmaxf(__uint_as_float(v[2 * j]) + b1[c16 * 16 + 2 * j], 0.f), fmaxf(__uint_as_float(v[2 * j + 1]) + b1[c16 * 16 + 2 * j + 1], 0.f));
            *reinterpret_cast<uint4*>(a2_row + (c16 * 2) * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(a2_row + (c16 * 2 + 1) * 2048) = make_uint4(h[4], h[5], h[6], h[7]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- 4. layer 2 + epilogue 2 ----
    if (tid == 0) {
        const uint64_t da = umma_desc(smem_u32(smem + kOffA2), 2048, 128), db = umma_desc(smem_u32(smem + kOffW2), 256, 128);
#pragma unroll
        for (uint32_t s = 0; s < 4; ++s)   // K = 64 = 4 x 16
            umma_f16(tmem, da + (uint64_t)((s * 2 * 2048) >> 4), db + (uint64_t)((s * 2 * 256) >> 4), umma_idesc(128, 16), s);
        umma_commit(bar2);
    }
    if (warp < 4) {
        mbar_wait(bar2, parity);
        tc_fence_after();
        uint32_t v[16];
        NR3D_TMEM_LD16(tmem + ((warp * 32) << 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t row = warp * 32 + lane;
        const uint32_t i = idx_s[row];
        if (i != 0xffffffffu) {
            const float* b2 = reinterpret_cast<const float*>(smem + kOffB2);
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) + b2[j];
            float s = o[0];
            if (dec.activation == 1) s = __expf(s);
            else if (dec.activation == 2) s = s > 20.f ? s : log1pf(__expf(s));
            else if (dec.activation == 3) s = fmaxf(s, 0.f);
            sigma[i] = s;
            if (out16) {
                float4* dst = reinterpret_cast<float4*>(out16 + (uint64_t)i * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();   // idx_s / A1 / TMEM are free for the next tile
    tc_fence_after();
    }  // tile loop
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
}

int make_table_public(const nr3d_lotd_meta* m, LotdTable& tab);  // lotd_fast.cu

}  // namespace nr3d

using namespace nr3d;

extern "C" int nr3d_lotd_fused_density_fwd(const nr3d_lotd_meta* meta, uint64_t N, const void* xs, const void* params, int32_t max_level,
                                           const void* w1_packed, const float* b1, const void* w2_packed, const float* b2, int32_t activation,
                                           float* sigma, float* out16, void* stream) {
    NR3D_CHECK(meta != nullptr, "fused_density: null meta");
    NR3D_CHECK(meta->hash_only && meta->n_dims_to_encode == 3 && meta->n_feat_per_pseudo_lvl == 2 && meta->n_pseudo_levels == 16,
               "fused_density: needs a Dense/Hash-only meta with D=3, F=2 and 16 pseudo levels (32 features)");
    NR3D_CHECK(N < (1ull << 32) - 1, "fused_density: N must be < 2^32 - 1");
    NR3D_CHECK(activation >= 0 && activation <= 3, "fused_density: activation code %d not in [0, 3]", (int)activation);
    if (N == 0) return 0;
    NR3D_CHECK(xs && params && w1_packed && w2_packed && sigma, "fused_density: null argument");
    NR3D_CHECK(((uintptr_t)w1_packed & 15) == 0 && ((uintptr_t)w2_packed & 15) == 0 && (!out16 || ((uintptr_t)out16 & 15) == 0),
               "fused_density: packed weights / out16 must be 16-byte aligned");
    LotdTable tab;
    make_table_public(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), nullptr, params, max_level, meta->n_params, 0u, meta->n_pseudo_levels};
    fast_levels(meta, in);
    FusedDec dec{reinterpret_cast<const uint4*>(w1_packed), reinterpret_cast<const uint4*>(w2_packed), b1, b2, activation};
    const uint64_t n_tiles = div_up<uint64_t>(N, 128);
    const unsigned grid = (unsigned)(n_tiles < (uint64_t)kSMs * kFusedCtasPerSm ? n_tiles : (uint64_t)kSMs * kFusedCtasPerSm);
    lotd_fused_density_kernel<<<grid, kFusedThreads, 0, (cudaStream_t)stream>>>(tab, in, dec, sigma, out16);
    NR3D_LAUNCH_CHECK("lotd_fused_density");
    return 0;
}
