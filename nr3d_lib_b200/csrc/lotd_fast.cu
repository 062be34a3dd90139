// lotd_fast.cu -- B200 fast path of the LoTD encoder for the Dense/Hash ("hash-only") configuration, D = 3, F_pl = 2,
// fp32 or fp16 parameters, single scene: the workload of BASELINE.json's headline metric.
//
// What bounds this workload on B200 (measured, scripts/ubench_mem.cu, scripts/ubench_pair.cu -> profiles/r1_ubench_*.txt):
//   * the 48.5 MB parameter table is L2 resident, so DRAM only sees x, y, dL_dy (about 1.3 GB per fwd+bwd step);
//   * a warp-wide gather costs one LSU slot per DISTINCT 128-BYTE LINE it touches (287 G lanes/s when every lane has its
//     own line = 148 SMs x 1.965 GHz; 572 G lanes/s when lane pairs share a line), independent of the access width;
//   * a warp-wide red.global.add costs one L2 slot per distinct 32-byte sector / 16-byte chunk (231 G lanes/s random,
//     460 G lanes/s when lane pairs share a chunk), f32 / v2.f32 / v4.f32 cost the same, and same-address reductions
//     from different SMs serialise in the L2 slice (coarse levels: up to 4.5x slower).
// So the levers are (1) lanes of one instruction sharing lines / sectors, (2) merging same-address contributions before they
// reach L2, (3) fewer instructions (both kernels end up ~85 % issue-bound).  This file implements them:
//   1. points are binned once per step by a 128^3 cell key (x fastest) with a counting sort; forward and backward walk
//      the points in that order, so coarse and middle levels hit few lines per warp;
//   2. TWO ADJACENT LANES share one point (see "pair layout" below): the x-neighbour corners of a Hash level (z-neighbours
//      of a Dense level) are fetched / scattered by the two lanes of a pair in the same instruction and coalesce in hardware;
//   3. in the backward pass, runs of points in the same cell sum their corner contributions through a shared-memory tile
//      and issue one reduction per corner per run;
//   4. y and dL_dy are accessed as [N, n_enc] rows staged through shared memory (one coalesced 128-byte access per point),
//      so the sort permutation costs no partial-sector traffic.
// (The thread-per-point kernels of the first iteration lost the A/B by 20 % -- profiles/r1_ab_pair_layout.txt -- and are gone.)
// Results are identical to the generic kernels up to fp32 summation order (same index functions, same weights).
#include "lotd_pair.cuh"
#include <string.h>

namespace nr3d {

// tunables (overridable with -D for A/B runs, see scripts/build_variants.py)
#ifndef NR3D_BIN_RES        // 0: chosen per call from the number of points (about two points per bin), else fixed (A/B runs)
#define NR3D_BIN_RES 0
#endif
#ifndef NR3D_BIN_ORDER      // 0: x fastest, 1: z fastest
#define NR3D_BIN_ORDER 0
#endif
#ifndef NR3D_FWD_UNROLL
#define NR3D_FWD_UNROLL 1
#endif
#ifndef NR3D_BWD_SHORTRUN     // 1: runs of 2-3 points in one cell are summed through two shuffle steps, the head of the run issues the reductions
#define NR3D_BWD_SHORTRUN 1     // A/B on B200: 2083 -> 2203 Msamples/s (profiles/r1_ab_tunables.txt)
#endif
#ifndef NR3D_FWD_THREADS
#define NR3D_FWD_THREADS 256
#endif
#ifndef NR3D_BWD_THREADS
#define NR3D_BWD_THREADS 128
#endif
// bins per axis of the point sort.  The kernels like about two points per bin (A/B on B200: 4 Mi uniform points 128^3 > 64^3, 256^3;
// 30 Mi ray samples 256^3 > 192^3 > 128^3, profiles/r1_ab_tunables.txt), so the resolution follows the point count.
static inline uint32_t bin_res_for(uint64_t N) {
    if (NR3D_BIN_RES) return NR3D_BIN_RES;
    return N >= (12ull << 20) ? 256u : 128u;
}
constexpr int kFastThreads = NR3D_FWD_THREADS;
constexpr int kBwdThreads = NR3D_BWD_THREADS;
constexpr int kScanBlockF = 1024;

__device__ __forceinline__ uint32_t bin_key(float x, float y, float z, uint32_t res) {
    const uint32_t bx = min(res - 1, (uint32_t)fmaxf(x * (float)res, 0.f));
    const uint32_t by = min(res - 1, (uint32_t)fmaxf(y * (float)res, 0.f));
    const uint32_t bz = min(res - 1, (uint32_t)fmaxf(z * (float)res, 0.f));
    return NR3D_BIN_ORDER == 0 ? (bz * res + by) * res + bx : (bx * res + by) * res + bz;
}

// pass 1: bin key and rank of the point inside its bin (the rank makes the scatter pass atomic-free)
__global__ void __launch_bounds__(256) sort_hist_kernel(uint64_t N, uint32_t res, const float* __restrict__ x, uint32_t* __restrict__ hist, uint2* __restrict__ keyrank) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t k = bin_key(x[i * 3], x[i * 3 + 1], x[i * 3 + 2], res);
    const uint32_t r = atomicAdd(hist + k, 1u);
    keyrank[i] = make_uint2(k, r);
}

// exclusive scan of `hist` in place (three launches, like pack_ops.cu's scan but for uint32)
__global__ void __launch_bounds__(kScanBlockF) scanu_block_sums(uint32_t n, const uint32_t* __restrict__ v, uint32_t* __restrict__ bs) {
    __shared__ uint32_t ws[32];
    const uint32_t i = blockIdx.x * kScanBlockF + threadIdx.x;
    uint32_t s = i < n ? v[i] : 0;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t t = ws[threadIdx.x];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
        if (threadIdx.x == 0) bs[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(kScanBlockF) scanu_of_sums(uint32_t nb, uint32_t* __restrict__ bs) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += kScanBlockF) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t x = i < nb ? bs[i] : 0;
        uint32_t s = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, s, d); if ((threadIdx.x & 31) >= d) s += t; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = ws[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, d); if (threadIdx.x >= d) w += t; }
            ws[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t off = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
        const uint32_t carry = carry_s;
        if (i < nb) bs[i] = carry + off + s - x;
        __syncthreads();
        if (threadIdx.x == kScanBlockF - 1) carry_s = carry + off + s;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kScanBlockF) scanu_apply(uint32_t n, uint32_t* __restrict__ v, const uint32_t* __restrict__ bs) {
    __shared__ uint32_t ws[32];
    const uint32_t i = blockIdx.x * kScanBlockF + threadIdx.x;
    const uint32_t x = i < n ? v[i] : 0;
    uint32_t s = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, s, d); if ((threadIdx.x & 31) >= d) s += t; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = ws[threadIdx.x];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, d); if (threadIdx.x >= d) w += t; }
        ws[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t off = bs[blockIdx.x] + ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0);
    if (i < n) v[i] = off + s - x;
}

// pass 2: sorted record = (x, y, z, original index) written with ONE 16-byte store per point
__global__ void __launch_bounds__(256) sort_scatter_kernel(uint64_t N, const float* __restrict__ x, const uint2* __restrict__ keyrank,
                                                           const uint32_t* __restrict__ offsets, float4* __restrict__ xs) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint2 kr = keyrank[i];
    const uint32_t pos = __ldg(offsets + kr.x) + kr.y;
    xs[pos] = make_float4(x[i * 3 + 0], x[i * 3 + 1], x[i * 3 + 2], __uint_as_float((uint32_t)i));
}

// ------------------------------------------------------------------------------------------------------------
// per-(point, level) geometry shared by forward and backward
// ------------------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------------------
// "pair" layout (default).  Measured on B200 (scripts/ubench_pair.cu -> profiles/r1_ubench_pair.txt): a warp-wide gather
// costs one LSU slot per DISTINCT 128-BYTE LINE (two lanes in one line, even in different sectors: 572 G lanes/s vs 287),
// a warp-wide red.global costs one L2 slot per DISTINCT 32-BYTE SECTOR (two lanes in one sector: 460 G lanes/s vs 231).
// The two corners of a Hash level that differ in x sit at (x ^ h) and ((x+1) ^ h): the same 128-byte line 15/16 of the time
// and the same sector 3/4 of the time, for odd x as well -- which the one-thread-per-point layout above can only exploit
// for even x (16-byte accesses).  So here TWO ADJACENT LANES share one point: lane side s handles the four corners with
// x + s (Hash) or z + s (Dense, z fastest), the pair's partial sums meet with one shuffle, and the hardware coalesces the
// pair's accesses.  Hash-level cost drops from 6 to ~4.25 lines (gather) and from 6 to 5 sectors (reduction) per point.
// ------------------------------------------------------------------------------------------------------------
constexpr int kPairRowStride = 34;   // floats per staged row: row k starts at bank 2k -> conflict-free pair writes and row reads
constexpr int kPairTileStride = 12;  // floats per lane in the run-merge tile (8 used): conflict-free 16-byte stores


// parameter-type specifics: fp32 tables accumulate in fp32; fp16 tables accumulate every term in half like the reference
// (linear_interpolate.cuh:118) and scatter with packed-half reductions.
template <typename PT> struct PairIO;
template <> struct PairIO<float> {
    static __device__ __forceinline__ float2 load2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
    static __device__ __forceinline__ float ldcs(const float* p) { return __ldcs(p); }
    static __device__ __forceinline__ void red2(float* p, float a, float b) { red_add_v2_f32(p, a, b); }
};
template <> struct PairIO<__half> {
    static __device__ __forceinline__ float2 load2(const __half* p) { return __half22float2(__ldg(reinterpret_cast<const __half2*>(p))); }
    static __device__ __forceinline__ float ldcs(const __half* p) { return __half2float(__ldcs(p)); }
    static __device__ __forceinline__ void red2(__half* p, float a, float b) { red_add_h2(p, __floats2half2_rn(a, b)); }
};

template <typename PT>
__global__ void __launch_bounds__(kFastThreads)
lotd_pair_fwd_kernel(const __grid_constant__ LotdTable tab, const FastIn in, PT* __restrict__ y, int64_t ys_n, int64_t ys_f) {
    using C = Cvt<PT>;
    const PT* params = reinterpret_cast<const PT*>(in.params);
    __shared__ float rows[kFastThreads / 32][16 * kPairRowStride];
    const uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool active = p < in.N;
    const int lane = threadIdx.x & 31;
    const uint32_t side = lane & 1;
    const int k = lane >> 1;  // point slot inside the warp
    float* myrows = rows[threadIdx.x >> 5];
    float4 rec = make_float4(0.5f, 0.5f, 0.5f, 0.f);
    if (active) rec = __ldcs(in.xs + p);
    const float x = rec.x, yv = rec.y, z = rec.z;
    const uint64_t i = __float_as_uint(rec.w);
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const bool staged = (ys_f == 1);
    uint32_t chunk_base = 0;
    constexpr int kFwdUnroll = NR3D_FWD_UNROLL;
#pragma unroll kFwdUnroll
    for (uint32_t pl = 0; pl < tab.n_pseudo; ++pl) {
        const uint32_t level = tab.map_level[pl];
        float r0 = 0.f, r1 = 0.f;  // (for fp16 tables these always hold half-representable values)
        if ((int32_t)level <= in.max_level) {
            Geo2 g;
            pair_geo(tab.lv[level], (uint32_t)tab.map_cnt[pl] * 2u, smooth, x, yv, z, side, g);
            float2 v[4];  // all four loads are issued before the first use
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = PairIO<PT>::load2(params + g.e[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                r0 = C::to_f(C::add(C::from_f(r0), C::from_f(g.w[q] * v[q].x)));
                r1 = C::to_f(C::add(C::from_f(r1), C::from_f(g.w[q] * v[q].y)));
            }
            r0 = C::to_f(C::add(C::from_f(r0), C::from_f(__shfl_xor_sync(0xffffffffu, r0, 1))));
            r1 = C::to_f(C::add(C::from_f(r1), C::from_f(__shfl_xor_sync(0xffffffffu, r1, 1))));
        }
        const float mine = side ? r1 : r0;  // lane `side` owns feature 2 * pl + side of its point
        if (staged) {
            const uint32_t c = pl * 2u - chunk_base;
            myrows[k * kPairRowStride + c + side] = mine;
            const bool last = (pl + 1 == tab.n_pseudo);
            if (c + 2 == 32 || last) {  // flush the chunk: one coalesced row store per point of the warp
                const uint32_t width = c + 2;
                __syncwarp();
#pragma unroll 4
                for (int r = 0; r < 16; ++r) {
                    const uint64_t ir = __shfl_sync(0xffffffffu, (uint32_t)i, 2 * r);
                    const bool ok = __shfl_sync(0xffffffffu, (int)active, 2 * r);
                    if (ok && (uint32_t)lane < width) st_cs(y + (int64_t)ir * ys_n + chunk_base + lane, C::from_f(myrows[r * kPairRowStride + lane]));
                }
                __syncwarp();
                chunk_base += 32;
            }
        } else if (active) {
            st_cs(y + (int64_t)i * ys_n + (int64_t)(pl * 2 + side) * ys_f, C::from_f(mine));
        }
    }
}

// SECOND = false: dL/dparam.  SECOND = true: d(dL/dx)/dparam . dL_ddLdx (second-order backward of NeuS-style eikonal terms,
// reference kernel_lod_hashonly_backward_input_backward_grid, lotd_hash_only.h:472-695): the same scatter with the corner weights
// replaced by sum_d ddx[d] * dw[d][corner] (pair_geo_d), ddx = dL_ddLdx [N,3] read at the point's original index.
template <typename PT, bool SECOND>
__global__ void __launch_bounds__(kBwdThreads)
lotd_pair_bwd_kernel(const __grid_constant__ LotdTable tab, const FastIn in, const PT* __restrict__ dLdy, int64_t gs_n, int64_t gs_f,
                     const float* __restrict__ ddx, PT* __restrict__ grad) {
    using C = Cvt<PT>;
    __shared__ __align__(16) float tile[kBwdThreads / 32][32 * kPairTileStride];
    __shared__ float rows[kBwdThreads / 32][16 * kPairRowStride];
    const uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool active = p < in.N;
    const int lane = threadIdx.x & 31;
    const uint32_t side = lane & 1;
    const int k = lane >> 1;
    float* mytile = tile[threadIdx.x >> 5];
    float* myrows = rows[threadIdx.x >> 5];
    float4 rec = make_float4(0.5f, 0.5f, 0.5f, 0.f);
    if (active) rec = __ldcs(in.xs + p);
    const float x = rec.x, yv = rec.y, z = rec.z;
    const uint64_t i = __float_as_uint(rec.w);
    const PT* grow = dLdy + (int64_t)i * gs_n;
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const bool staged = (gs_f == 1);
    float gx[3] = {0.f, 0.f, 0.f};
    if (SECOND && active) { gx[0] = __ldg(ddx + i * 3); gx[1] = __ldg(ddx + i * 3 + 1); gx[2] = __ldg(ddx + i * 3 + 2); }
    uint32_t chunk_base = 0;
    for (uint32_t pl = 0; pl < tab.n_pseudo; ++pl) {
        const uint32_t level = tab.map_level[pl];
        float g0 = 0.f, g1 = 0.f;
        if (staged) {
            if (pl * 2u == chunk_base + 32u) chunk_base += 32u;
            if (pl * 2u == chunk_base) {  // stage the next (up to) 32 features of the warp's 16 rows: one coalesced row read per point
                const uint32_t width = min(32u, tab.n_enc - chunk_base);
                __syncwarp();
#pragma unroll 4
                for (int r = 0; r < 16; ++r) {
                    const uint64_t ir = __shfl_sync(0xffffffffu, (uint32_t)i, 2 * r);
                    const bool ok = __shfl_sync(0xffffffffu, (int)active, 2 * r);
                    if (ok && (uint32_t)lane < width) myrows[r * kPairRowStride + lane] = PairIO<PT>::ldcs(dLdy + (int64_t)ir * gs_n + chunk_base + lane);
                }
                __syncwarp();
            }
            g0 = myrows[k * kPairRowStride + pl * 2u - chunk_base];
            g1 = myrows[k * kPairRowStride + pl * 2u - chunk_base + 1];
        } else if (active) {
            g0 = C::to_f(grow[(int64_t)(pl * 2) * gs_f]);
            g1 = C::to_f(grow[(int64_t)(pl * 2 + 1) * gs_f]);
        }
        if ((int32_t)level > in.max_level) continue;  // uniform
        const LevelDesc& L = tab.lv[level];
        Geo2 g;
        if (SECOND) {
            float dw[3][4];
            pair_geo_d(L, (uint32_t)tab.map_cnt[pl] * 2u, smooth, x, yv, z, side, g, dw);
#pragma unroll
            for (int q = 0; q < 4; ++q) g.w[q] = gx[0] * dw[0][q] + gx[1] * dw[1][q] + gx[2] * dw[2][q];
        } else {
            pair_geo(L, (uint32_t)tab.map_cnt[pl] * 2u, smooth, x, yv, z, side, g);
        }
        // points of one run (consecutive points in the same cell) merge their contributions before touching L2
        const bool can_key = L.res[0] <= 1024u && L.res[1] <= 1024u && L.res[2] <= 1024u;
        uint32_t hmask = 0x55555555u;  // bit 2k set: point k starts a run
        if (can_key) {
            const uint32_t key = active ? g.key : (0xffffffffu - (uint32_t)k);
            const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 2);
            hmask = __ballot_sync(0xffffffffu, k == 0 || key != prev) & 0x55555555u;
        }
        // run of my point: points [s0, e0), length r, my position j
        const uint32_t le = hmask & (0xffffffffu >> (31 - 2 * k));
        const int s0 = (31 - __clz(le)) >> 1;
        const uint32_t above = hmask & (0xffffffffu << (2 * k + 1));
        const int e0 = above ? ((__ffs(above) - 1) >> 1) : 16;
        const int r = e0 - s0, j = k - s0;
        const bool longrun = active && r >= 4;  // shorter runs are not worth the detour through shared memory
        // every lane ends up with at most four contributions (entry g.e[q], value pair); cv[q]: this lane issues corner q
        float cx[4], cy[4];
        bool cv[4];
        if (!__any_sync(0xffffffffu, longrun)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { cx[q] = g.w[q] * g0; cy[q] = g.w[q] * g1; cv[q] = active; }
        } else {
            *reinterpret_cast<float4*>(mytile + lane * kPairTileStride) = make_float4(g.w[0] * g0, g.w[0] * g1, g.w[1] * g0, g.w[1] * g1);
            *reinterpret_cast<float4*>(mytile + lane * kPairTileStride + 4) = make_float4(g.w[2] * g0, g.w[2] * g1, g.w[3] * g0, g.w[3] * g1);
            __syncwarp();
            // A long run is reduced by its first 4 * nch positions: position j sums corner q = j & 3 over the points
            // s0 + c, s0 + c + nch, ... (c = j >> 2), the nch partial sums of a corner then meet through two shuffles and
            // positions 0..3 issue ONE reduction per corner for the whole run.
            const int nch = r >> 2;  // 1..4 groups of four positions
            const int qj = j & 3, c = j >> 2;
            float2 acc = make_float2(0.f, 0.f);
            if (longrun && c < nch) {
                for (int m = s0 + c; m < e0; m += nch) {
                    const float2 t = *reinterpret_cast<const float2*>(mytile + (2 * m + side) * kPairTileStride + qj * 2);
                    acc.x += t.x; acc.y += t.y;
                }
            }
            const int lim = nch << 2;
            float tx = __shfl_down_sync(0xffffffffu, acc.x, 16), ty = __shfl_down_sync(0xffffffffu, acc.y, 16);
            if (longrun && j + 8 < lim) { acc.x += tx; acc.y += ty; }   // positions j + 8 (and, through them, j + 12)
            tx = __shfl_down_sync(0xffffffffu, acc.x, 8); ty = __shfl_down_sync(0xffffffffu, acc.y, 8);
            if (longrun && j < 4 && j + 4 < lim) { acc.x += tx; acc.y += ty; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (longrun) { cv[q] = (j < 4) && (q == qj); cx[q] = acc.x; cy[q] = acc.y; }
                else { cv[q] = active; cx[q] = g.w[q] * g0; cy[q] = g.w[q] * g1; }
            }
            __syncwarp();
        }
#if NR3D_BWD_SHORTRUN
        // Runs of two or three points in one cell: the hardware does not merge lanes that hit the same entry, so they would cost one L2
        // packet each.  Two shuffle steps towards the head of the run (the points of a run are neighbours in the warp) and only the head
        // issues the four reductions.
        const bool shortrun = active && (r == 2 || r == 3);
        if (__any_sync(0xffffffffu, shortrun)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float ux = __shfl_down_sync(0xffffffffu, cx[q], 2), uy = __shfl_down_sync(0xffffffffu, cy[q], 2);
                if (shortrun && j + 1 < r) { cx[q] += ux; cy[q] += uy; }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float ux = __shfl_down_sync(0xffffffffu, cx[q], 4), uy = __shfl_down_sync(0xffffffffu, cy[q], 4);
                if (shortrun && j + 2 < r) { cx[q] += ux; cy[q] += uy; }
                if (shortrun && j > 0) cv[q] = false;
            }
        }
#endif
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (cv[q]) PairIO<PT>::red2(grad + g.e[q], cx[q], cy[q]);
    }
}

// Forward with dy/dx (NeuS-style callers need the nablas): same walk as lotd_pair_fwd_kernel plus, per (point, feature), the three
// derivatives  dy/dx_d = sum_corners dw[d][corner] * value(corner)  (reference kernel_lod_hash_only_with_dydx, lotd_hash_only.h:
// 164-378).  dy_dx is written row-major [N, n_enc, 3] at the point's original index through a second staged tile (96 floats per
// point and 32 features = three coalesced 128-byte stores), accumulated in fp32 for either table type (INPUT_T in the reference).
constexpr int kDydxThreads = 128;
constexpr int kDChunk = 16;          // features per staged dy/dx tile (8 pseudo levels): 48 floats = 192 bytes per point and flush
constexpr int kPairDRowStride = 50;  // floats per staged dy/dx row (48 used): lane (k, side) writes bank 18k + 3 side + const, conflict free

template <typename PT>
__global__ void __launch_bounds__(kDydxThreads)
lotd_pair_fwd_dydx_kernel(const __grid_constant__ LotdTable tab, const FastIn in, PT* __restrict__ y, float* __restrict__ dydx) {
    using C = Cvt<PT>;
    const PT* params = reinterpret_cast<const PT*>(in.params);
    __shared__ float rows[kDydxThreads / 32][16 * kPairRowStride];
    __shared__ float drows[kDydxThreads / 32][16 * kPairDRowStride];
    const uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool active = p < in.N;
    const int lane = threadIdx.x & 31;
    const uint32_t side = lane & 1;
    const int k = lane >> 1;
    float* myrows = rows[threadIdx.x >> 5];
    float* mydrows = drows[threadIdx.x >> 5];
    float4 rec = make_float4(0.5f, 0.5f, 0.5f, 0.f);
    if (active) rec = __ldcs(in.xs + p);
    const float x = rec.x, yv = rec.y, z = rec.z;
    const uint64_t i = __float_as_uint(rec.w);
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const uint32_t n_enc = tab.n_enc;
    uint32_t chunk_base = 0, dchunk_base = 0;
    for (uint32_t pl = 0; pl < tab.n_pseudo; ++pl) {
        const uint32_t level = tab.map_level[pl];
        float r0 = 0.f, r1 = 0.f;
        float d0[3] = {0.f, 0.f, 0.f}, d1[3] = {0.f, 0.f, 0.f};
        if ((int32_t)level <= in.max_level) {
            Geo2 g;
            float dw[3][4];
            pair_geo_d(tab.lv[level], (uint32_t)tab.map_cnt[pl] * 2u, smooth, x, yv, z, side, g, dw);
            float2 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = PairIO<PT>::load2(params + g.e[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                r0 = C::to_f(C::add(C::from_f(r0), C::from_f(g.w[q] * v[q].x)));
                r1 = C::to_f(C::add(C::from_f(r1), C::from_f(g.w[q] * v[q].y)));
#pragma unroll
                for (int d = 0; d < 3; ++d) { d0[d] += dw[d][q] * v[q].x; d1[d] += dw[d][q] * v[q].y; }
            }
            r0 = C::to_f(C::add(C::from_f(r0), C::from_f(__shfl_xor_sync(0xffffffffu, r0, 1))));
            r1 = C::to_f(C::add(C::from_f(r1), C::from_f(__shfl_xor_sync(0xffffffffu, r1, 1))));
            // lane `side` owns feature 2 * pl + side of its point: it needs its partner's partial sums of that feature only
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float give = side ? d0[d] : d1[d];                       // the partner's feature
                const float take = __shfl_xor_sync(0xffffffffu, give, 1);
                if (side) d1[d] += take; else d0[d] += take;
            }
        }
        const uint32_t c = pl * 2u - chunk_base, dc = pl * 2u - dchunk_base;
        myrows[k * kPairRowStride + c + side] = side ? r1 : r0;
#pragma unroll
        for (int d = 0; d < 3; ++d) mydrows[k * kPairDRowStride + (dc + side) * 3 + d] = side ? d1[d] : d0[d];
        const bool last = (pl + 1 == tab.n_pseudo);
        const bool flush_y = (c + 2 == 32) || last, flush_d = (dc + 2 == kDChunk) || last;
        if (flush_d) {  // one coalesced y row (every second time) and one and a half coalesced dy/dx lines per point
            const uint32_t ywidth = c + 2, dwidth = (dc + 2) * 3;
            __syncwarp();
#pragma unroll 2
            for (int r = 0; r < 16; ++r) {
                const uint64_t ir = __shfl_sync(0xffffffffu, (uint32_t)i, 2 * r);
                const bool ok = __shfl_sync(0xffffffffu, (int)active, 2 * r);
                if (!ok) continue;
                if (flush_y && (uint32_t)lane < ywidth) st_cs(y + ir * n_enc + chunk_base + lane, C::from_f(myrows[r * kPairRowStride + lane]));
                float* drow = dydx + (ir * n_enc + dchunk_base) * 3;
                if ((uint32_t)lane < dwidth) __stcs(drow + lane, mydrows[r * kPairDRowStride + lane]);
                if ((uint32_t)lane + 32u < dwidth) __stcs(drow + 32 + lane, mydrows[r * kPairDRowStride + 32 + lane]);
            }
            __syncwarp();
            dchunk_base += kDChunk;
            if (flush_y) chunk_base += 32;
        }
    }
}

static int make_table(const nr3d_lotd_meta* m, LotdTable& tab) {
    memset(&tab, 0, sizeof(tab));
    for (uint32_t l = 0; l < m->n_levels; ++l) {
        LevelDesc& d = tab.lv[l];
        for (int k = 0; k < 4; ++k) d.res[k] = m->level_res[l][k];
        d.type = m->level_types[l]; d.n_feat = m->level_n_feats[l]; d.size = m->level_sizes[l]; d.offset = m->level_offsets[l];
    }
    for (uint32_t p = 0; p < m->n_pseudo_levels; ++p) { tab.map_level[p] = (uint8_t)m->map_levels[p]; tab.map_cnt[p] = (uint8_t)m->map_cnt[p]; }
    tab.n_levels = m->n_levels; tab.n_pseudo = m->n_pseudo_levels; tab.n_enc = m->n_encoded_dims; tab.n_params = m->n_params;
    tab.interp = m->interpolation_type; tab.fpl = m->n_feat_per_pseudo_lvl;
    return 0;
}
int make_table_public(const nr3d_lotd_meta* m, LotdTable& tab) { return make_table(m, tab); }  // for lotd_fused.cu

static int check_fast(const nr3d_lotd_meta* m, int32_t param_dtype, uint64_t N) {
    NR3D_CHECK(m != nullptr, "LoTDEncoding: null meta");
    const bool dtype_ok = param_dtype == NR3D_F32 || param_dtype == NR3D_F16;
    NR3D_CHECK(m->hash_only && m->n_dims_to_encode == 3 && m->n_feat_per_pseudo_lvl == 2 && dtype_ok,
               "LoTDEncoding: the sorted fast path needs a Dense/Hash-only meta with D=3, 2 features per pseudo level and fp32 / fp16 params");
    NR3D_CHECK(N < (1ull << 32), "LoTDEncoding: batch_size must be < 2^32");
    return 0;
}

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_lotd_sort_points(uint64_t N, const float* x, void* xs /* float4 [N] */, void* ws, uint64_t* ws_bytes, void* stream) {
    const uint32_t res = bin_res_for(N);
    const uint32_t bins = res * res * res;
    const uint32_t nb = div_up<uint32_t>(bins, kScanBlockF);
    const uint64_t need = (uint64_t)bins * 4 + (uint64_t)nb * 4 + 64 + N * 8;
    if (ws == nullptr) {
        NR3D_CHECK(ws_bytes != nullptr, "sort_points: null ws_bytes");
        *ws_bytes = need;
        return 0;
    }
    NR3D_CHECK(ws_bytes && *ws_bytes >= need, "sort_points: workspace too small");
    NR3D_CHECK(N < (1ull << 32), "sort_points: N must be < 2^32");
    if (N == 0) return 0;
    NR3D_CHECK(x && xs, "sort_points: null argument");
    NR3D_CHECK((reinterpret_cast<uintptr_t>(xs) & 15u) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15u) == 0, "sort_points: xs / ws must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* hist = reinterpret_cast<uint32_t*>(ws);
    uint32_t* bs = hist + bins;
    uint2* keyrank = reinterpret_cast<uint2*>(reinterpret_cast<char*>(ws) + (((uint64_t)bins * 4 + (uint64_t)nb * 4 + 63) / 64) * 64);
    cudaMemsetAsync(hist, 0, (size_t)bins * 4, st);
    const unsigned grid = (unsigned)div_up<uint64_t>(N, 256);
    sort_hist_kernel<<<grid, 256, 0, st>>>(N, res, x, hist, keyrank);
    NR3D_LAUNCH_CHECK("sort_hist");
    scanu_block_sums<<<nb, kScanBlockF, 0, st>>>(bins, hist, bs);
    NR3D_LAUNCH_CHECK("sort_scan1");
    scanu_of_sums<<<1, kScanBlockF, 0, st>>>(nb, bs);
    NR3D_LAUNCH_CHECK("sort_scan2");
    scanu_apply<<<nb, kScanBlockF, 0, st>>>(bins, hist, bs);
    NR3D_LAUNCH_CHECK("sort_scan3");
    sort_scatter_kernel<<<grid, 256, 0, st>>>(N, x, keyrank, hist, reinterpret_cast<float4*>(xs));
    NR3D_LAUNCH_CHECK("sort_scatter");
    return 0;
}

int nr3d_lotd_fwd_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const void* params,
                         int32_t max_level, void* y, int64_t y_stride_n, int64_t y_stride_f, void* stream) {
    if (int rc = check_fast(meta, param_dtype, N)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && params && y, "LoTDEncoding::fwd_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), params, max_level, (uint32_t)((reinterpret_cast<uintptr_t>(params) & 15u) == 0)};
    const unsigned grid = (unsigned)div_up<uint64_t>(2 * N, kFastThreads);
    if (param_dtype == NR3D_F16) lotd_pair_fwd_kernel<__half><<<grid, kFastThreads, 0, (cudaStream_t)stream>>>(tab, in, (__half*)y, y_stride_n, y_stride_f);
    else lotd_pair_fwd_kernel<float><<<grid, kFastThreads, 0, (cudaStream_t)stream>>>(tab, in, (float*)y, y_stride_n, y_stride_f);
    NR3D_LAUNCH_CHECK("lotd_fast_fwd");
    return 0;
}

int nr3d_lotd_bwd_param_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const void* dL_dy,
                               int64_t dLdy_stride_n, int64_t dLdy_stride_f, int32_t max_level, void* dL_dparam, void* stream) {
    if (int rc = check_fast(meta, param_dtype, N)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && dL_dy && dL_dparam, "LoTDEncoding::bwd_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), nullptr, max_level, 1u};
    const unsigned grid = (unsigned)div_up<uint64_t>(2 * N, kBwdThreads);
    if (param_dtype == NR3D_F16) lotd_pair_bwd_kernel<__half, false><<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(tab, in, (const __half*)dL_dy, dLdy_stride_n, dLdy_stride_f, nullptr, (__half*)dL_dparam);
    else lotd_pair_bwd_kernel<float, false><<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(tab, in, (const float*)dL_dy, dLdy_stride_n, dLdy_stride_f, nullptr, (float*)dL_dparam);
    NR3D_LAUNCH_CHECK("lotd_fast_bwd");
    return 0;
}

int nr3d_lotd_fwd_dydx_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const void* params,
                              int32_t max_level, void* y, void* dy_dx, void* stream) {
    if (int rc = check_fast(meta, param_dtype, N)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && params && y && dy_dx, "LoTDEncoding::fwd_dydx_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), params, max_level, (uint32_t)((reinterpret_cast<uintptr_t>(params) & 15u) == 0)};
    const unsigned grid = (unsigned)div_up<uint64_t>(2 * N, kDydxThreads);
    if (param_dtype == NR3D_F16) lotd_pair_fwd_dydx_kernel<__half><<<grid, kDydxThreads, 0, (cudaStream_t)stream>>>(tab, in, (__half*)y, (float*)dy_dx);
    else lotd_pair_fwd_dydx_kernel<float><<<grid, kDydxThreads, 0, (cudaStream_t)stream>>>(tab, in, (float*)y, (float*)dy_dx);
    NR3D_LAUNCH_CHECK("lotd_fast_fwd_dydx");
    return 0;
}

int nr3d_lotd_bwd_param2_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const void* dL_dy,
                                int64_t dLdy_stride_n, int64_t dLdy_stride_f, const float* dL_ddLdx, int32_t max_level, void* dL_dparam,
                                void* stream) {
    if (int rc = check_fast(meta, param_dtype, N)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && dL_dy && dL_ddLdx && dL_dparam, "LoTDEncoding::bwd_param2_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), nullptr, max_level, 1u};
    const unsigned grid = (unsigned)div_up<uint64_t>(2 * N, kBwdThreads);
    if (param_dtype == NR3D_F16) lotd_pair_bwd_kernel<__half, true><<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(tab, in, (const __half*)dL_dy, dLdy_stride_n, dLdy_stride_f, dL_ddLdx, (__half*)dL_dparam);
    else lotd_pair_bwd_kernel<float, true><<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(tab, in, (const float*)dL_dy, dLdy_stride_n, dLdy_stride_f, dL_ddLdx, (float*)dL_dparam);
    NR3D_LAUNCH_CHECK("lotd_fast_bwd2");
    return 0;
}

}  // extern "C"
