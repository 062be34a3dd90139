// lotd_fast.cu -- B200 fast path of the LoTD encoder for Dense/Hash ("hash-only") metas, D = 3, F = 2 / 4 / 8 features per pseudo level,
// fp32 or fp16 parameters, one scene or many (batch indices / batch_data_size): the workload of BASELINE.json's headline metric and of the
// reference's hash-only kernels (lotd_hash_only.h:15-695 serve the same set).
//
// What bounds this workload on B200 (measured, scripts/ubench_mem.cu, scripts/ubench_pair.cu -> profiles/r1_ubench_*.txt):
//   * the 48.5 MB parameter table is L2 resident, so DRAM only sees x, y, dL_dy (about 1.3 GB per fwd+bwd step);
//   * a warp-wide gather costs one LSU slot per DISTINCT 128-BYTE LINE it touches (287 G lanes/s when every lane has its
//     own line = 148 SMs x 1.965 GHz; 572 G lanes/s when lane pairs share a line), independent of the access width;
//   * a warp-wide red.global.add costs one L2 slot per distinct 32-byte sector / 16-byte chunk (231 G lanes/s random,
//     460 G lanes/s when lane pairs share a chunk), f32 / v2.f32 / v4.f32 cost the same, and same-address reductions
//     from different SMs serialise in the L2 slice (coarse levels: up to 4.5x slower).
// So the levers are (1) lanes of one instruction sharing lines / sectors, (2) merging same-address contributions before they
// reach L2, (3) fewer instructions.  This file implements them:
//   1. points are binned once per step by (scene, cell bin) with a counting sort (lotd_sort.cu; x-fastest bins -- brick-ordered bins lost
//      their A/B); forward and backward walk the points in that order, so coarse and middle levels hit few lines per warp;
//   2. TWO ADJACENT LANES share one point (see "pair layout" below): the x-neighbour corners of a Hash level (z-neighbours
//      of a Dense level) are fetched / scattered by the two lanes of a pair in the same instruction and coalesce in hardware;
//   3. in the backward pass, the points of a warp that fall into the same cell -- neighbours or not: groups from one match.any per level -- sum
//      their corner contributions by pointer jumping and the group's first lane issues one reduction per corner (profiles/r2_ab_merge.txt).
//      (A CTA-level stage -- shared-memory tiles that sum what is left per table entry -- is kept behind -DNR3D_BWD_TILES=1: it removes 31 % of
//      the L2 packets and is 2x slower, because fp32 adds into shared memory are CAS loops on sm_100a.)
//   4. y and dL_dy are accessed as [N, n_enc] rows staged through shared memory (one coalesced 128-byte access per point),
//      so the sort permutation costs no partial-sector traffic.
//   5. both kernels are issue-slot bound as much as memory bound, so everything that does not depend on the point is derived on the host
//      (FastLevel, lotd_pair.cuh), offsets are 32-bit and the row staging is predicated.
// Results are identical to the generic kernels up to fp32 summation order (same index functions, same weights).
#include "lotd_pair.cuh"
#include <string.h>
#include <math.h>

namespace nr3d {

// tunables (overridable with -D for A/B runs, see scripts/build_variants.py)
#ifndef NR3D_FWD_UNROLL
#define NR3D_FWD_UNROLL 1
#endif
#ifndef NR3D_FWD_THREADS
#define NR3D_FWD_THREADS 256
#endif
#ifndef NR3D_FWD_OCC          // resident threads per SM the F = 2 forward is compiled for (register cap through __launch_bounds__)
#define NR3D_FWD_OCC 1536     // A/B on B200 (profiles/r2_ab_tunables.txt): 2048 (30 registers, +21 % instructions) 0.873 ms, 1536 (39 registers) 0.843 ms, 1280 0.861 ms
#endif
#ifndef NR3D_BWD_THREADS      // 256 threads = 128 points per CTA: about one 8 x 4 x 2 brick of sort bins
#define NR3D_BWD_THREADS 256
#endif
#ifndef NR3D_FWD_TMA          // 1: A/B variant of the forward that stages the Dense levels' corner boxes with TMA bulk copies (see below)
#define NR3D_FWD_TMA 0
#endif
#ifndef NR3D_BWD_TILES        // 1: CTA-level shared-memory tiles in the backward (0: every run head scatters to L2 directly)
#define NR3D_BWD_TILES 0      // A/B on B200 (profiles/r2_ab_tiles.txt): OFF wins, 1.18 ms vs 2.45 - 3.6 ms -- see the note at smem_add2
#endif
#ifndef NR3D_TILE_FLOATS      // shared-memory floats per CTA for the backward tiles
#define NR3D_TILE_FLOATS 6144
#endif
#ifndef NR3D_BWD_OCC          // resident threads per SM the F = 2 backward is compiled for (register cap through __launch_bounds__), 0: no cap
#define NR3D_BWD_OCC 1536
#endif
#ifndef NR3D_BWD_MERGE        // how same-cell points of a warp are merged before the scatter: 0 contiguous runs only (round 1), 1 any lanes of the warp (match.any),
                              // 2 lanes linked over at most NR3D_LINK_DIST points (shuffles); 1 and 2 sum by pointer jumping
#define NR3D_BWD_MERGE 1
#endif
#ifndef NR3D_LINK_DIST        // NR3D_BWD_MERGE == 2: how many points ahead a lane looks for the next point of its cell
#define NR3D_LINK_DIST 3
#endif
#ifndef NR3D_BWD_NEIGH        // 1: on Hash levels the side-0 head of cell (X, y, z) hands its four sums to the side-1 head of cell (X - 1, y, z) (same entries)
#define NR3D_BWD_NEIGH 0
#endif
#ifndef NR3D_MERGE_DENSITY    // levels with res^3 <= NR3D_MERGE_DENSITY * (points per scene) try to merge; finer levels hold less than one point per cell
#define NR3D_MERGE_DENSITY 16 // (0: every level whose cell key fits 10 bits per axis, the round-1 rule)
#endif
constexpr int kFastThreads = NR3D_FWD_THREADS;
constexpr int kBwdThreads = NR3D_BWD_THREADS;
constexpr int kBwdWarps = kBwdThreads / 32;
constexpr int kTileFloats = NR3D_TILE_FLOATS;
constexpr int kMaxTiledLevels = 32;   // pseudo levels that can own a tile (one lane of warp 0 plans each)

// ------------------------------------------------------------------------------------------------------------
// "pair" layout.  Measured on B200 (scripts/ubench_pair.cu -> profiles/r1_ubench_pair.txt): a warp-wide gather
// costs one LSU slot per DISTINCT 128-BYTE LINE (two lanes in one line, even in different sectors: 572 G lanes/s vs 287),
// a warp-wide red.global costs one L2 slot per DISTINCT 32-BYTE SECTOR (two lanes in one sector: 460 G lanes/s vs 231).
// The two corners of a Hash level that differ in x sit at (x ^ h) and ((x+1) ^ h): the same 128-byte line 15/16 of the time
// and the same sector 3/4 of the time, for odd x as well -- which a one-thread-per-point layout can only exploit
// for even x (16-byte accesses).  So here TWO ADJACENT LANES share one point: lane side s handles the four corners with
// x + s (Hash) or z + s (Dense, z fastest), the pair's partial sums meet with one shuffle, and the hardware coalesces the
// pair's accesses.  Hash-level cost drops from 6 to ~4.25 lines (gather) and from 6 to 5 sectors (reduction) per point.
// ------------------------------------------------------------------------------------------------------------
constexpr int kPairRowStride = 34;   // floats per staged row: row k starts at bank 2k -> conflict-free pair writes and row reads

// scene (batch) of sorted record p: parameter base of the scene and whether the point takes part at all
__device__ __forceinline__ bool point_scene(const FastIn& in, uint64_t p, bool active, uint32_t& scene, uint32_t& pbase) {
    scene = 0; pbase = 0;
    if (!active) return false;
    if (in.scenes) {
        scene = __ldg(in.scenes + p);
        if (scene == 0xffffu) return false;   // batch_inds < 0: the point is skipped (zero output, no gradient)
        pbase = scene * in.n_params;
    }
    return true;
}

// Density head fused into the encoder (HEAD = true; row "glue" of DESIGN.md, the M2 workload): the stand-in decoder of the render step is
//   sigma = softplus(gain * sum_c y[c]),  alpha = 1 - exp(-sigma * delta)          (csrc/pipeline_ops.cu, nerf_ray_query.py:182)
// so the forward only needs the feature SUM of a point and the backward's dL/dy row is one scalar repeated n_enc times: the [S, n_enc]
// features and their gradient never reach HBM (4 x 128 B per sample of the M2 step).
struct HeadFwd { const float* deltas; float gain; float* sigma; float* alpha; };
struct HeadBwd { const float* d_alpha; const float* sigma; const float* alpha; const float* deltas; float gain; };

__device__ __forceinline__ float softplus_head(float v) { return v > 20.0f ? v : log1pf(expf(v)); }   // == F.softplus (beta 1, threshold 20)

template <typename PT, int F, bool HEAD = false>
__global__ void __launch_bounds__(kFastThreads, F == 2 ? NR3D_FWD_OCC / kFastThreads : 1)
lotd_pair_fwd_kernel(const __grid_constant__ LotdTable tab, const FastIn in, PT* __restrict__ y, int64_t ys_n, int64_t ys_f, const HeadFwd hd) {
    using C = Cvt<PT>;
    constexpr int H = F / 2;   // features of a pseudo level that one lane of the pair writes out
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
    uint32_t scene, pbase;
    const bool live = point_scene(in, p, active, scene, pbase);
    const float x = live ? rec.x : 0.5f, yv = live ? rec.y : 0.5f, z = live ? rec.z : 0.5f;
    const uint64_t i = __float_as_uint(rec.w);
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const bool staged = (ys_f == 1);
    const uint32_t itag = active ? (uint32_t)i : 0xffffffffu;   // (N < 2^32, so no point has index 2^32 - 1)
    uint32_t chunk_base = 0;
    float hsum = 0.f;   // HEAD: sum of the point's features, level after level
    constexpr int kFwdUnroll = NR3D_FWD_UNROLL;
#pragma unroll kFwdUnroll
    for (uint32_t pl = 0; pl < tab.n_pseudo; ++pl) {
        const uint32_t level = tab.map_level[pl];
        float r[F];  // (for fp16 tables these always hold half-representable values)
#pragma unroll
        for (int f = 0; f < F; ++f) r[f] = 0.f;
        if ((int32_t)level <= in.max_level) {
            Geo2 g;
            pair_geo(tab.lv[level], in.fl[level], (uint32_t)tab.map_cnt[pl] * F, pbase, smooth, x, yv, z, side, g);
            float v[4][F];  // all four loads are issued before the first use
#pragma unroll
            for (int q = 0; q < 4; ++q) load_feats<PT, F>(params + g.e[q], v[q]);
            if constexpr (std::is_same<PT, __half>::value) {
                // fp16 tables: every product is rounded to half and accumulated in half like the reference (linear_interpolate.cuh:118);
                // two features per packed instruction (cvt.rn.f16x2.f32 + HADD2) -- same values as the scalar chain, a third of the issue slots
                __half2 acc[F / 2];
#pragma unroll
                for (int j = 0; j < F / 2; ++j) acc[j] = __floats2half2_rn(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < F / 2; ++j) acc[j] = __hadd2(acc[j], __floats2half2_rn(g.w[q] * v[q][2 * j], g.w[q] * v[q][2 * j + 1]));
#pragma unroll
                for (int j = 0; j < F / 2; ++j) {
                    const uint32_t mine_bits = *reinterpret_cast<const uint32_t*>(&acc[j]);
                    const uint32_t other_bits = __shfl_xor_sync(0xffffffffu, mine_bits, 1);
                    const float2 t = __half22float2(__hadd2(acc[j], as_half2(other_bits)));
                    r[2 * j] = t.x; r[2 * j + 1] = t.y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int f = 0; f < F; ++f) r[f] = C::to_f(C::add(C::from_f(r[f]), C::from_f(g.w[q] * v[q][f])));
#pragma unroll
                for (int f = 0; f < F; ++f) r[f] = C::to_f(C::add(C::from_f(r[f]), C::from_f(__shfl_xor_sync(0xffffffffu, r[f], 1))));
            }
        }
        if constexpr (HEAD) {
#pragma unroll
            for (int f = 0; f < F; ++f) hsum += r[f];
        } else {
        // lane `side` owns features pl * F + side * H + [0, H) of its point
        float mine[H];
#pragma unroll
        for (int j = 0; j < H; ++j) mine[j] = live ? (side ? r[H + j] : r[j]) : 0.f;
        if (staged) {
            const uint32_t c = pl * F - chunk_base;
#pragma unroll
            for (int j = 0; j < H; ++j) myrows[k * kPairRowStride + c + side * H + j] = mine[j];
            const bool last = (pl + 1 == tab.n_pseudo);
            if (c + F == 32 || last) {  // flush the chunk: one coalesced row store per point of the warp (predicated: no branch per row)
                const bool inw = (uint32_t)lane < c + F;
                __syncwarp();
#pragma unroll 4
                for (int rr = 0; rr < 16; ++rr) {
                    const uint32_t ir = __shfl_sync(0xffffffffu, itag, 2 * rr);   // original index of point rr, 0xffffffff: no such point
                    st_cs_if(inw && ir != 0xffffffffu, y + (int64_t)ir * ys_n + chunk_base + lane, C::from_f(myrows[rr * kPairRowStride + lane]));
                }
                __syncwarp();
                chunk_base += 32;
            }
        } else if (active) {
#pragma unroll
            for (int j = 0; j < H; ++j) st_cs(y + (int64_t)i * ys_n + (int64_t)(pl * F + side * H + j) * ys_f, C::from_f(mine[j]));
        }
        }   // !HEAD
    }
    if (HEAD && active && side == 0) {   // skipped points (batch_inds < 0) have all-zero features like in the unfused path
        const float sg = softplus_head((live ? hsum : 0.f) * hd.gain);
        __stcs(hd.sigma + i, sg);
        __stcs(hd.alpha + i, 1.0f - expf(-sg * __ldcs(hd.deltas + i)));
    }
}

#if NR3D_FWD_TMA
// ------------------------------------------------------------------------------------------------------------
// A/B VARIANT (-DNR3D_FWD_TMA=1, scripts/ab_bench.py "fwd_tma*"): TMA staging of the Dense levels' corner tiles (BASELINE.json north_star:
// "TMA staging of grid-corner tiles into shared memory").  After the sort a CTA's 128 points sit on one x pencil of bins, so on a Dense
// level their corners form a small box [n0 x n1 x n2] of the z-fastest table.  Warp 0 plans one shared-memory tile per Dense level from the
// CTA's bounding box, every thread issues `cp.async.bulk.shared::cluster.global` copies (SASS UBLKCP; one per (x, y) row of the box:
// the rows are the only contiguous pieces, 16 - 48 bytes each), the CTA waits on an mbarrier and the gather loop reads those levels'
// corners from shared memory instead of L1 / L2.  Hash levels have no box to fetch (corners are scattered by the hash).
// MEASURED: see profiles/r2_ab_fwd_tma.txt -- kept out of the shipped build.
// ------------------------------------------------------------------------------------------------------------
#ifndef NR3D_FWD_TMA_FLOATS
#define NR3D_FWD_TMA_FLOATS 6144
#endif
constexpr int kTmaLevels = 8;                  // pseudo levels 0 .. 7 may own a tile (the NGP ladder has 6 Dense levels)
constexpr int kTmaFloats = NR3D_FWD_TMA_FLOATS;

struct TmaPlan {
    int32_t off[kTmaLevels];        // first float of the level's tile, -1: not staged
    uint32_t lo[kTmaLevels][3];     // smallest corner coordinate of the CTA's points on the level
    uint32_t n0[kTmaLevels], n1[kTmaLevels], rs[kTmaLevels];   // box rows along x, y; row stride in entries (covers n2 + alignment slack)
};

__global__ void __launch_bounds__(kFastThreads, 6)
lotd_pair_fwd_tma_kernel(const __grid_constant__ LotdTable tab, const FastIn in, float* __restrict__ y, int64_t ys_n) {
    constexpr int F = 2;
    const float* params = reinterpret_cast<const float*>(in.params);
    __shared__ float rows[kFastThreads / 32][16 * kPairRowStride];
    __shared__ __align__(128) float tile[kTmaFloats];
    __shared__ uint32_t s_box[6];
    __shared__ TmaPlan plan;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_total;
    const uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool active = p < in.N;
    const int lane = threadIdx.x & 31;
    const uint32_t side = lane & 1;
    const int k = lane >> 1;
    float* myrows = rows[threadIdx.x >> 5];
    float4 rec = make_float4(0.5f, 0.5f, 0.5f, 0.f);
    if (active) rec = __ldcs(in.xs + p);
    const float x = rec.x, yv = rec.y, z = rec.z;
    const uint64_t i = __float_as_uint(rec.w);
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    // ---- bounding box of the CTA's points (unit-cube coordinates; cell_pos is monotone, so it bounds the cells on every level) ----
    if (threadIdx.x < 6) s_box[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    {
        const uint32_t u[3] = {__float_as_uint(fmaxf(x, 0.f)), __float_as_uint(fmaxf(yv, 0.f)), __float_as_uint(fmaxf(z, 0.f))};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const uint32_t a = __reduce_min_sync(0xffffffffu, active ? u[d] : 0xffffffffu), b = __reduce_max_sync(0xffffffffu, active ? u[d] : 0u);
            if (lane == 0) { atomicMin(&s_box[d], a); atomicMax(&s_box[3 + d], b); }
        }
    }
    __syncthreads();
    // ---- plan: lane pl of warp 0 sizes the tile of pseudo level pl ----
    if (threadIdx.x < 32) {
        const uint32_t pl = threadIdx.x;
        uint32_t need = 0, lo[3] = {0, 0, 0}, n[3] = {0, 0, 0}, rs = 0;
        if (pl < (uint32_t)kTmaLevels && pl < tab.n_pseudo && s_box[0] != 0xffffffffu) {
            const uint32_t level = tab.map_level[pl];
            const LevelDesc& L = tab.lv[level];
            if ((int32_t)level <= in.max_level && L.type == NR3D_LOD_DENSE && L.n_feat == 2u && (L.offset & 1u) == 0u) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    lo[d] = (uint32_t)floorf(cell_pos(__uint_as_float(s_box[d]), L.res[d]));
                    n[d] = (uint32_t)floorf(cell_pos(__uint_as_float(s_box[3 + d]), L.res[d])) + 2u - lo[d];
                }
                rs = (n[2] + 2u) & ~1u;     // entries per staged row: n2 plus one entry of alignment slack, even
                const uint64_t sz = (uint64_t)n[0] * n[1] * rs * 2u;
                need = (uint32_t)min(sz, (uint64_t)kTmaFloats + 1);
            }
        }
        uint32_t incl = need;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl = min(incl + t, 2u * kTmaFloats); }
        const bool fits = need > 0 && incl <= (uint32_t)kTmaFloats;
        if (pl < (uint32_t)kTmaLevels) {
            plan.off[pl] = fits ? (int32_t)(incl - need) : -1;
#pragma unroll
            for (int d = 0; d < 3; ++d) plan.lo[pl][d] = lo[d];
            plan.n0[pl] = n[0]; plan.n1[pl] = n[1]; plan.rs[pl] = rs;
        }
        uint32_t bytes = fits ? need * 4u : 0u;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, m);
        if (lane == 0) {
            s_total = bytes;
            if (bytes) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        }
    }
    __syncthreads();
    // ---- bulk copies: one per (x, y) row of every staged box ----
    for (uint32_t pl = 0; pl < (uint32_t)kTmaLevels; ++pl) {
        const int32_t off = plan.off[pl];
        if (off < 0) continue;
        const LevelDesc& L = tab.lv[tab.map_level[pl]];
        const uint32_t n0 = plan.n0[pl], n1 = plan.n1[pl], rs = plan.rs[pl];
        for (uint32_t r = threadIdx.x; r < n0 * n1; r += kFastThreads) {
            const uint32_t a = r / n1, b = r - a * n1;
            const uint32_t ge = (L.offset >> 1) + ((plan.lo[pl][0] + a) * L.res[1] + plan.lo[pl][1] + b) * L.res[2] + plan.lo[pl][2];   // 8-byte entries
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + off + r * rs * 2u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(params + (uint64_t)(ge & ~1u) * 2u), "r"(rs * 8u), "r"(bar) : "memory");
        }
    }
    if (s_total) {
        uint32_t ok = 0;
        for (uint32_t spin = 0; !ok; ++spin) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
            if (spin > (1u << 26)) __trap();
        }
    }
    // ---- gather loop: staged Dense levels read their corners from the tile ----
    uint32_t chunk_base = 0;
    for (uint32_t pl = 0; pl < tab.n_pseudo; ++pl) {
        const uint32_t level = tab.map_level[pl];
        float r[F] = {0.f, 0.f};
        if ((int32_t)level <= in.max_level) {
            const LevelDesc& L = tab.lv[level];
            Geo2 g;
            pair_geo(L, in.fl[level], (uint32_t)tab.map_cnt[pl] * F, 0u, smooth, x, yv, z, side, g);
            float2 v[4];
            const int32_t off = pl < (uint32_t)kTmaLevels ? plan.off[pl] : -1;
            if (off >= 0) {
                const uint32_t n1 = plan.n1[pl], rs = plan.rs[pl];
#pragma unroll
                for (int q = 0; q < 4; ++q) {      // corner (c0 + (q & 1), c1 + (q >> 1), c2 + side), see pair_geo
                    const uint32_t cx = g.c[0] + (q & 1), cy = g.c[1] + (q >> 1);
                    const uint32_t row = (cx - plan.lo[pl][0]) * n1 + (cy - plan.lo[pl][1]);
                    const uint32_t row_ge = (L.offset >> 1) + (cx * L.res[1] + cy) * L.res[2] + plan.lo[pl][2];
                    const uint32_t col = g.c[2] + side - plan.lo[pl][2] + (row_ge & 1u);
                    v[q] = *reinterpret_cast<const float2*>(tile + off + (row * rs + col) * 2u);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float2*>(params + g.e[q]));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { r[0] += g.w[q] * v[q].x; r[1] += g.w[q] * v[q].y; }
#pragma unroll
            for (int f = 0; f < F; ++f) r[f] += __shfl_xor_sync(0xffffffffu, r[f], 1);
        }
        const float mine = active ? (side ? r[1] : r[0]) : 0.f;
        const uint32_t c = pl * F - chunk_base;
        myrows[k * kPairRowStride + c + side] = mine;
        const bool last = (pl + 1 == tab.n_pseudo);
        if (c + F == 32 || last) {
            const uint32_t width = c + F;
            __syncwarp();
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
                const uint64_t ir = __shfl_sync(0xffffffffu, (uint32_t)i, 2 * rr);
                const bool ok = __shfl_sync(0xffffffffu, (int)active, 2 * rr);
                if (ok && (uint32_t)lane < width) __stcs(y + (int64_t)ir * ys_n + chunk_base + lane, myrows[rr * kPairRowStride + lane]);
            }
            __syncwarp();
            chunk_base += 32;
        }
    }
}
#endif  // NR3D_FWD_TMA

// ------------------------------------------------------------------------------------------------------------
// backward: dL/dparam scatter
// ------------------------------------------------------------------------------------------------------------
// fp32 pair added to a shared-memory tile entry: ONE 64-bit compare-and-swap loop (fp32 shared-memory atomics are CAS loops on sm_100a
// anyway -- SASS ATOMS.CAST.SPIN -- so the pair costs the same as a single float).
// MEASURED (profiles/r2_ab_tiles.txt, 4 Mi points): the tiles remove 31 % of the L2 reduction packets as simulated (scripts/sim_tiles.py)
// and still LOSE by 2x: the backward goes from 1.18 ms to 2.45 ms (x-fastest bins) / 2.9 - 3.6 ms (bricks, 3072 - 12288 tile floats).
// ATOMS.CAS.64 retires about one lane every two clocks per SM (9 tiled levels x ~100 lane operations x 1771 warps per SM = 2 ms), i.e.
// a shared-memory float add costs ~3x what the L2 reduction unit charges for a whole packet (231 G packets/s = 1.26 clocks per SM).
// The only cheap fp32 adder with conflict resolution on this chip is the L2 reduction unit; the code stays for the record (-DNR3D_BWD_TILES=1).
__device__ __forceinline__ void smem_add2(float* addr, float a, float b) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(p), assumed;
    do {
        assumed = old;
        const float lo = __uint_as_float((uint32_t)assumed) + a, hi = __uint_as_float((uint32_t)(assumed >> 32)) + b;
        old = atomicCAS(p, assumed, ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo));
    } while (old != assumed);
}

struct TilePlan {                       // one entry per pseudo level, written by warp 0
    int32_t off[kMaxTiledLevels];       // first float of the level's tile inside the CTA tile memory, -1: not tiled
    uint32_t lo[kMaxTiledLevels][3];    // smallest corner coordinate of the CTA's points on the level
    uint32_t n[kMaxTiledLevels][3];     // corners per axis
};

// SECOND = false: dL/dparam.  SECOND = true: d(dL/dx)/dparam . dL_ddLdx (second-order backward of NeuS-style eikonal terms,
// reference kernel_lod_hashonly_backward_input_backward_grid, lotd_hash_only.h:472-695): the same scatter with the corner weights
// replaced by sum_d ddx[d] * dw[d][corner] (pair_geo_d), ddx = dL_ddLdx [N,3] read at the point's original index.
// HEAD = true: dL/dy[i, :] = g_i for every feature, g_i = dL/dalpha_i * delta_i * (1 - alpha_i) * gain * sigmoid(gain * s_i) (HeadBwd).
template <typename PT, int F, bool SECOND, bool HEAD = false>
__global__ void __launch_bounds__(kBwdThreads, (F == 2 && !SECOND && NR3D_BWD_OCC) ? NR3D_BWD_OCC / kBwdThreads : 1)
lotd_pair_bwd_kernel(const __grid_constant__ LotdTable tab, const FastIn in, const PT* __restrict__ dLdy, int64_t gs_n, int64_t gs_f,
                     const float* __restrict__ ddx, PT* __restrict__ grad, const HeadBwd hd) {
    using C = Cvt<PT>;
    extern __shared__ __align__(16) float smem[];
#if NR3D_BWD_TILES
    float* ctile = smem;                                                    // [kTileFloats] CTA tiles (NR3D_BWD_TILES)
    __shared__ uint32_t s_box[8];   // min x, y, z bits | max x, y, z bits | min scene | max scene over the live points of the CTA
    __shared__ TilePlan plan;
#endif
    float* myrows = smem + (NR3D_BWD_TILES ? kTileFloats : 0) + (threadIdx.x >> 5) * (16 * kPairRowStride);
    const uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool active = p < in.N;
    const int lane = threadIdx.x & 31;
    const uint32_t side = lane & 1;
    const int k = lane >> 1;
    float4 rec = make_float4(0.5f, 0.5f, 0.5f, 0.f);
    if (active) rec = __ldcs(in.xs + p);
    uint32_t scene, pbase;
    const bool live = point_scene(in, p, active, scene, pbase);
    const float x = live ? rec.x : 0.5f, yv = live ? rec.y : 0.5f, z = live ? rec.z : 0.5f;
    const uint64_t i = __float_as_uint(rec.w);
    const PT* grow = dLdy + (int64_t)i * gs_n;
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const bool staged = !HEAD && (gs_f == 1);
    const uint32_t ltag = live ? (uint32_t)i : 0xffffffffu;   // (N < 2^32, so no point has index 2^32 - 1)
    float ghead = 0.f;
    if (HEAD && live) {
        // alpha = 1 - exp(-sigma * delta): d alpha / d sigma = delta * (1 - alpha);  sigma = softplus(gain * s): d sigma / d s = gain * (1 - exp(-sigma))
        const float sg = __ldcs(hd.sigma + i);
        const float g_sigma = __ldcs(hd.d_alpha + i) * __ldcs(hd.deltas + i) * (1.0f - __ldcs(hd.alpha + i));
        ghead = C::to_f(C::from_f(g_sigma * (sg > 20.0f ? 1.0f : (1.0f - expf(-sg))) * hd.gain));
    }
    float gx[3] = {0.f, 0.f, 0.f};
    if (SECOND && live) { gx[0] = __ldg(ddx + i * 3); gx[1] = __ldg(ddx + i * 3 + 1); gx[2] = __ldg(ddx + i * 3 + 2); }
    // a point starts a new run on every level when its scene differs from the previous point's
#if !NR3D_BWD_MERGE
    const uint32_t scene_prev = __shfl_up_sync(0xffffffffu, scene, 2);
    const bool scene_break = scene != scene_prev;
#endif

#if NR3D_BWD_TILES
    // ---- CTA bounding box of the live points (in unit-cube coordinates: 6 reductions instead of 6 per level) and the tile plan ----
    if (threadIdx.x < 8) s_box[threadIdx.x] = (threadIdx.x < 3 || threadIdx.x == 6) ? 0xffffffffu : 0u;
    for (int t = threadIdx.x; t < kTileFloats / 4; t += kBwdThreads) reinterpret_cast<float4*>(ctile)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    {
        // non-negative floats order like their bit patterns; coordinates below 0 land in cell 0 like 0 does
        const uint32_t ux = __float_as_uint(fmaxf(x, 0.f)), uy = __float_as_uint(fmaxf(yv, 0.f)), uz = __float_as_uint(fmaxf(z, 0.f));
        const uint32_t mn[4] = {live ? ux : 0xffffffffu, live ? uy : 0xffffffffu, live ? uz : 0xffffffffu, live ? scene : 0xffffffffu};
        const uint32_t mx[4] = {live ? ux : 0u, live ? uy : 0u, live ? uz : 0u, live ? scene : 0u};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const uint32_t a = __reduce_min_sync(0xffffffffu, mn[d]), b = __reduce_max_sync(0xffffffffu, mx[d]);
            if (lane == 0) { atomicMin(&s_box[d == 3 ? 6 : d], a); atomicMax(&s_box[d == 3 ? 7 : 3 + d], b); }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t pl = threadIdx.x;
        uint32_t need = 0;   // floats of this level's tile (0: level not processed at all)
        bool fits = false;
        uint32_t lo[3] = {0, 0, 0}, n[3] = {0, 0, 0};
        const bool one_scene = s_box[6] != 0xffffffffu && s_box[6] == s_box[7];
        if (pl >= in.pl_begin && pl < in.pl_end && pl < kMaxTiledLevels && one_scene) {
            const uint32_t level = tab.map_level[pl];
            if ((int32_t)level <= in.max_level) {
                const LevelDesc& L = tab.lv[level];
                uint64_t size = 1;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    lo[d] = (uint32_t)floorf(cell_pos(__uint_as_float(s_box[d]), L.res[d]));
                    const uint32_t hi = (uint32_t)floorf(cell_pos(__uint_as_float(s_box[3 + d]), L.res[d])) + 1u;
                    n[d] = min(hi - lo[d] + 1u, 4096u);
                    size *= n[d];
                }
                need = (uint32_t)min(size * F, (uint64_t)kTileFloats + 1);
                fits = true;
            }
        }
        uint32_t incl = need;   // inclusive prefix over the pseudo levels: tiles are handed out in level order while the memory lasts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl = min(incl + t, 2u * kTileFloats); }
        fits = fits && incl <= (uint32_t)kTileFloats;
        plan.off[pl] = fits ? (int32_t)(incl - need) : -1;
#pragma unroll
        for (int d = 0; d < 3; ++d) { plan.lo[pl][d] = lo[d]; plan.n[pl][d] = n[d]; }
    }
    __syncthreads();
#endif

    uint32_t chunk_base = 0xffffffffu;   // first feature of the staged chunk (none yet)
    for (uint32_t pl = in.pl_begin; pl < in.pl_end; ++pl) {
        const uint32_t level = tab.map_level[pl];
        float gv[F];
#pragma unroll
        for (int f = 0; f < F; ++f) gv[f] = HEAD ? ghead : 0.f;
        if (HEAD) {
        } else if (staged) {
            const uint32_t want_base = (pl * (uint32_t)F) / 32u * 32u;
            if (want_base != chunk_base) {  // stage the next (up to) 32 features of the warp's 16 rows: one coalesced row read per point
                const uint32_t width = min(32u, tab.n_enc - want_base);
                // (a level-group launch only touches the sectors of its own features)
                const bool inw = (uint32_t)lane < width && want_base + lane >= in.pl_begin * F && want_base + lane < in.pl_end * F;
                __syncwarp();
#pragma unroll 4
                for (int rr = 0; rr < 16; ++rr) {   // predicated loads (0 for absent points / columns, never used): no branch per row
                    const uint32_t ir = __shfl_sync(0xffffffffu, ltag, 2 * rr);
                    myrows[rr * kPairRowStride + lane] = ldcs_f_if<PT>(inw && ir != 0xffffffffu, dLdy + (int64_t)ir * gs_n + want_base + lane);
                }
                __syncwarp();
            }
            chunk_base = want_base;
            if (live) {
#pragma unroll
                for (int f = 0; f < F; ++f) gv[f] = myrows[k * kPairRowStride + pl * F - chunk_base + f];
            }
        } else if (live) {
#pragma unroll
            for (int f = 0; f < F; ++f) gv[f] = C::to_f(grow[(int64_t)(pl * F + f) * gs_f]);
        }
        if ((int32_t)level > in.max_level) continue;  // uniform
        const LevelDesc& L = tab.lv[level];
        Geo2 g;
        if (SECOND) {
            float dw[3][4];
            pair_geo_d(L, in.fl[level], (uint32_t)tab.map_cnt[pl] * F, pbase, smooth, x, yv, z, side, g, dw);
#pragma unroll
            for (int q = 0; q < 4; ++q) g.w[q] = gx[0] * dw[0][q] + gx[1] * dw[1][q] + gx[2] * dw[2][q];
        } else {
            pair_geo(L, in.fl[level], (uint32_t)tab.map_cnt[pl] * F, pbase, smooth, x, yv, z, side, g);
        }
        // every lane holds four contributions (entry g.e[q], F values)
        float cx[4][F];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int f = 0; f < F; ++f) cx[q][f] = g.w[q] * gv[f];
        // Points of the warp in the same cell of this level merge their contributions before they leave the warp: the hardware merges different
        // entries of a sector, not lanes that hit the same entry.  The cell key holds 10 bits per axis; levels finer than in.merge_res hold less
        // than one point per cell and skip the attempt (it costs issue slots the kernel does not have: 76 % busy in the round-2 ncu capture).
        bool issue = live;   // this lane issues its four contributions
        const bool can_key = max(L.res[0], max(L.res[1], L.res[2])) <= in.merge_res;
#if NR3D_BWD_MERGE
        if (can_key) {
            // ANY lanes of the warp with the same (cell, side, scene) form a group -- not only neighbours: the sort orders the points by bin, and a
            // cell boundary that cuts a bin leaves its points interleaved (A B A B: four "runs", two cells).  scripts/sim_sectors.py: 55.6 -> 51.2
            // packets per point.  Every lane learns its successor in the group (`next`, 32: none) and whether it is the group's first lane; the sums
            // then travel towards the first lane by pointer jumping (after step t a lane holds the sum over itself and its 2^t - 1 successors, `next`
            // points 2^t members ahead).
            int next = 32;
            bool first = true;
#if NR3D_BWD_MERGE == 1   // groups from match.any
            // (match.any costs about as much as it finds DISTINCT values -- measured, profiles/r2_ab_merge.txt -- so the key leaves the side out: the
            // two lanes of a point fall into one group that the parity mask splits again, and all idle lanes share one sentinel)
            const uint32_t mkey = live ? g.key : 0x80000000u;
            const uint32_t smask = in.scenes ? __match_any_sync(0xffffffffu, scene) : 0xffffffffu;   // lanes of my scene (recomputed per level: one register less)
            const uint32_t same_cell = __match_any_sync(0xffffffffu, mkey);   // (every lane of the warp executes the match)
            const uint32_t grp = live ? (same_cell & smask & (0x55555555u << side)) : (1u << lane);
            {
                const uint32_t above = grp & ~((2u << lane) - 1u);
                if (above) next = __ffs(above) - 1;
                first = (grp & ((1u << lane) - 1u)) == 0u;
            }
#else                     // links to the nearest same-cell point at most NR3D_LINK_DIST points away (three shuffles instead of a match)
            const uint32_t mkey = live ? (g.key | (side << 30)) : (0x80000000u | (uint32_t)lane);
#pragma unroll
            for (int d = NR3D_LINK_DIST; d >= 1; --d) {
                const uint32_t kd = __shfl_down_sync(0xffffffffu, mkey, 2 * d);
                bool same = lane + 2 * d < 32 && kd == mkey;
                if (in.scenes) same = same && __shfl_down_sync(0xffffffffu, scene, 2 * d) == scene;
                if (same) next = lane + 2 * d;   // the nearest one wins (d runs downwards)
                const uint32_t b = __ballot_sync(0xffffffffu, same);   // bit l: lane l has a same-cell successor d points on, so lane l + 2 d is not a first lane
                if (lane >= 2 * d && ((b >> (lane - 2 * d)) & 1u)) first = false;
            }
#endif
            if (__any_sync(0xffffffffu, next < 32)) {
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) {
                    if (d > 1 && !__any_sync(0xffffffffu, next < 32)) break;   // warp uniform
                    const bool take = next < 32;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int f = 0; f < F; ++f) {
                            const float u = __shfl_sync(0xffffffffu, cx[q][f], next);
                            if (take) cx[q][f] += u;
                        }
                    const int nn = __shfl_sync(0xffffffffu, next, next);
                    next = take ? nn : 32;
                }
                issue = live && first;
            }
#if NR3D_BWD_NEIGH
            if (L.type != NR3D_LOD_DENSE) {   // warp uniform
                // The side-1 head of cell (X - 1, y, z) and the side-0 head of cell (X, y, z) address the same four entries (X ^ h(y + dy, z + dz)):
                // the lower lane of such a pair takes the other's sums.
                const uint32_t nkey = issue ? (g.key + side) : 0x80000000u;   // x coordinate of my corners (<= res - 1: no carry into y)
#if NR3D_BWD_MERGE != 1
                const uint32_t smask = in.scenes ? __match_any_sync(0xffffffffu, scene) : 0xffffffffu;
#endif
                const uint32_t pairm = __match_any_sync(0xffffffffu, nkey) & smask;
                const uint32_t other = (issue && __popc(pairm) == 2) ? (pairm ^ (1u << lane)) : 0u;   // (link merging can leave several heads per cell: those keep their sums)
                if (__any_sync(0xffffffffu, other != 0u)) {
                    const int ol = other ? (__ffs(other) - 1) : lane;
                    const bool recv = other != 0u && lane < ol;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int f = 0; f < F; ++f) {
                            const float u = __shfl_sync(0xffffffffu, cx[q][f], ol);
                            if (recv) cx[q][f] += u;
                        }
                    if (other != 0u && !recv) issue = false;
                }
            }
#endif
        }
#else
        if (can_key) {
            // (round-1 rule) runs = CONSECUTIVE points of the warp in the same cell
            const uint32_t key = live ? g.key : (0xffffffffu - (uint32_t)k);
            const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 2);
            const uint32_t hmask = __ballot_sync(0xffffffffu, k == 0 || key != prev || scene_break) & 0x55555555u;  // bit 2k set: point k starts a run
            if (hmask != 0x55555555u) {
                // run of my point: points [s0, e0), length r, my position j
                const uint32_t le = hmask & (0xffffffffu >> (31 - 2 * k));
                const int s0 = (31 - __clz(le)) >> 1;
                const uint32_t above = hmask & (0xffffffffu << (2 * k + 1));
                const int e0 = above ? ((__ffs(above) - 1) >> 1) : 16;
                const int r = e0 - s0, j = k - s0;
                const int rmax = __reduce_max_sync(0xffffffffu, r);
                // segmented reduction towards the head of every run: after the step with distance d, position j holds the sum over
                // positions [j, min(j + 2d, r)) -- the points of a run are neighbours in the warp, 2 lanes apart
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) {
                    if (d >= rmax) break;   // warp uniform
                    const bool take = j + d < r;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int f = 0; f < F; ++f) {
                            const float u = __shfl_down_sync(0xffffffffu, cx[q][f], 2 * d);
                            if (take) cx[q][f] += u;
                        }
                }
                issue = live && j == 0;
            }
        }
#endif
#if NR3D_BWD_TILES
        const int32_t toff = pl < kMaxTiledLevels ? plan.off[pl] : -1;   // CTA uniform
        if (toff >= 0) {
            // tile entry of corner (a, b, c): z fastest for Dense levels (as in the table), x fastest for Hash levels (x ^ h(y, z): the x
            // neighbours of a cell share a sector 3 times out of 4), so that the flush below leaves in few packets
            const uint32_t n0 = plan.n[pl][0], n1 = plan.n[pl][1], n2 = plan.n[pl][2];
            const bool dense = L.type == NR3D_LOD_DENSE;
            const uint32_t st0 = dense ? n1 * n2 : 1u, st1 = dense ? n2 : n0, st2 = dense ? 1u : n0 * n1;
            const uint32_t l0 = (g.c[0] - plan.lo[pl][0]) * st0 + (g.c[1] - plan.lo[pl][1]) * st1 + (g.c[2] - plan.lo[pl][2]) * st2;
            if (issue) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t lq = l0 + (dense ? (q & 1) * st0 + (q >> 1) * st1 + side * st2 : side * st0 + (q & 1) * st1 + (q >> 1) * st2);
                    float* t = ctile + toff + lq * F;
#pragma unroll
                    for (int f = 0; f < F; f += 2) smem_add2(t + f, cx[q][f], cx[q][f + 1]);
                }
            }
            continue;
        }
#endif
        if (issue) {
#pragma unroll
            for (int q = 0; q < 4; ++q) red_feats<PT, F>(grad + g.e[q], cx[q]);
        }
    }

#if NR3D_BWD_TILES
    // ---- flush the tiles: ONE reduction per touched entry and CTA, consecutive lanes on consecutive entries ----
    __syncthreads();
    const uint32_t tbase = s_box[6] == 0xffffffffu ? 0u : s_box[6] * in.n_params;   // the tiled CTA lives in one scene
    const uint32_t npl = min(in.pl_end, (uint32_t)kMaxTiledLevels);
    for (uint32_t pl = in.pl_begin; pl < npl; ++pl) {
        const int32_t toff = plan.off[pl];
        if (toff < 0) continue;
        const LevelDesc& L = tab.lv[tab.map_level[pl]];
        const uint32_t n0 = plan.n[pl][0], n1 = plan.n[pl][1], n2 = plan.n[pl][2];
        const uint32_t size = n0 * n1 * n2;
        const bool dense = L.type == NR3D_LOD_DENSE;
        const bool pow2 = (L.size & (L.size - 1u)) == 0;
        const uint32_t base = tbase + L.offset + (uint32_t)tab.map_cnt[pl] * F;
        for (uint32_t e = threadIdx.x; e < size; e += kBwdThreads) {
            float v[F];
            bool any = false;
#pragma unroll
            for (int f = 0; f < F; f += 2) {
                const float2 t = *reinterpret_cast<const float2*>(ctile + toff + e * F + f);
                v[f] = t.x; v[f + 1] = t.y;
                any = any || t.x != 0.f || t.y != 0.f;
            }
            if (!any) continue;   // untouched corner of the bounding box (or a sum that is exactly zero)
            uint32_t idx;
            if (dense) {
                const uint32_t c = e % n2, ab = e / n2, b = ab % n1, a = ab / n1;
                idx = ((plan.lo[pl][0] + a) * L.res[1] + plan.lo[pl][1] + b) * L.res[2] + plan.lo[pl][2] + c;
            } else {
                const uint32_t a = e % n0, bc = e / n0, b = bc % n1, c = bc / n1;
                const uint32_t h = (plan.lo[pl][0] + a) ^ ((plan.lo[pl][1] + b) * 2654435761u) ^ ((plan.lo[pl][2] + c) * 805459861u);
                idx = pow2 ? (h & (L.size - 1u)) : (h % L.size);
            }
            red_feats<PT, F>(grad + base + idx * L.n_feat, v);
        }
    }
#endif
}

// Forward with dy/dx (NeuS-style callers need the nablas): same walk as lotd_pair_fwd_kernel plus, per (point, feature), the three
// derivatives  dy/dx_d = sum_corners dw[d][corner] * value(corner)  (reference kernel_lod_hash_only_with_dydx, lotd_hash_only.h:
// 164-378).  dy_dx is written row-major [N, n_enc, 3] at the point's original index through a second staged tile (96 floats per
// point and 32 features = three coalesced 128-byte stores), accumulated in fp32 for either table type (INPUT_T in the reference).
constexpr int kDydxThreads = 128;
constexpr int kDChunk = 16;          // features per staged dy/dx tile: 48 floats = 192 bytes per point and flush
constexpr int kPairDRowStride = 50;  // floats per staged dy/dx row (48 used): lane (k, side) writes bank 18k + 3 side + const, conflict free for F = 2

template <typename PT, int F>
__global__ void __launch_bounds__(kDydxThreads)
lotd_pair_fwd_dydx_kernel(const __grid_constant__ LotdTable tab, const FastIn in, PT* __restrict__ y, float* __restrict__ dydx) {
    using C = Cvt<PT>;
    constexpr int H = F / 2;
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
    uint32_t scene, pbase;
    const bool live = point_scene(in, p, active, scene, pbase);
    const float x = live ? rec.x : 0.5f, yv = live ? rec.y : 0.5f, z = live ? rec.z : 0.5f;
    const uint64_t i = __float_as_uint(rec.w);
    const bool smooth = tab.interp == NR3D_INTERP_SMOOTHSTEP;
    const uint32_t n_enc = tab.n_enc;
    uint32_t chunk_base = 0, dchunk_base = 0;
    for (uint32_t pl = 0; pl < tab.n_pseudo; ++pl) {
        const uint32_t level = tab.map_level[pl];
        float r[F], dd[F][3];
#pragma unroll
        for (int f = 0; f < F; ++f) { r[f] = 0.f; dd[f][0] = dd[f][1] = dd[f][2] = 0.f; }
        if ((int32_t)level <= in.max_level) {
            Geo2 g;
            float dw[3][4];
            pair_geo_d(tab.lv[level], in.fl[level], (uint32_t)tab.map_cnt[pl] * F, pbase, smooth, x, yv, z, side, g, dw);
            float v[4][F];
#pragma unroll
            for (int q = 0; q < 4; ++q) load_feats<PT, F>(params + g.e[q], v[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    r[f] = C::to_f(C::add(C::from_f(r[f]), C::from_f(g.w[q] * v[q][f])));
#pragma unroll
                    for (int d = 0; d < 3; ++d) dd[f][d] += dw[d][q] * v[q][f];
                }
#pragma unroll
            for (int f = 0; f < F; ++f) r[f] = C::to_f(C::add(C::from_f(r[f]), C::from_f(__shfl_xor_sync(0xffffffffu, r[f], 1))));
        }
        // lane `side` owns features pl * F + side * H + [0, H) of its point: it needs its partner's partial sums of those features only
        const uint32_t c = pl * F - chunk_base, dc = pl * F - dchunk_base;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            myrows[k * kPairRowStride + c + side * H + j] = live ? (side ? r[H + j] : r[j]) : 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float give = side ? dd[j][d] : dd[H + j][d];                       // the partner's feature
                const float take = __shfl_xor_sync(0xffffffffu, give, 1);
                const float own = (side ? dd[H + j][d] : dd[j][d]) + take;
                mydrows[k * kPairDRowStride + (dc + side * H + j) * 3 + d] = live ? own : 0.f;
            }
        }
        const bool last = (pl + 1 == tab.n_pseudo);
        const bool flush_y = (c + F == 32) || last, flush_d = (dc + F == kDChunk) || last;
        if (flush_d) {  // one coalesced y row (every second time) and one and a half coalesced dy/dx lines per point
            const uint32_t ywidth = c + F, dwidth = (dc + F) * 3;
            __syncwarp();
#pragma unroll 2
            for (int rr = 0; rr < 16; ++rr) {
                const uint64_t ir = __shfl_sync(0xffffffffu, (uint32_t)i, 2 * rr);
                const bool ok = __shfl_sync(0xffffffffu, (int)active, 2 * rr);
                if (!ok) continue;
                if (flush_y && (uint32_t)lane < ywidth) st_cs(y + ir * n_enc + chunk_base + lane, C::from_f(myrows[rr * kPairRowStride + lane]));
                float* drow = dydx + (ir * n_enc + dchunk_base) * 3;
                if ((uint32_t)lane < dwidth) __stcs(drow + lane, mydrows[rr * kPairDRowStride + lane]);
                if ((uint32_t)lane + 32u < dwidth) __stcs(drow + 32 + lane, mydrows[rr * kPairDRowStride + 32 + lane]);
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

static int check_fast(const nr3d_lotd_meta* m, int32_t param_dtype, uint64_t N, uint32_t n_scenes, const void* table) {
    NR3D_CHECK(m != nullptr, "LoTDEncoding: null meta");
    const bool dtype_ok = param_dtype == NR3D_F32 || param_dtype == NR3D_F16;
    const uint32_t F = m->n_feat_per_pseudo_lvl;
    NR3D_CHECK(m->hash_only && m->n_dims_to_encode == 3 && (F == 2 || F == 4 || F == 8) && dtype_ok,
               "LoTDEncoding: the sorted fast path needs a Dense/Hash-only meta with D=3, 2 / 4 / 8 features per pseudo level and fp32 / fp16 params");
    NR3D_CHECK(N < (1ull << 32), "LoTDEncoding: batch_size must be < 2^32");
    NR3D_CHECK((uint64_t)(n_scenes ? n_scenes : 1) * m->n_params < (1ull << 32), "LoTDEncoding: the sorted fast path indexes the tables with 32 bits (n_scenes * n_params < 2^32)");
    NR3D_CHECK((reinterpret_cast<uintptr_t>(table) & 15u) == 0, "LoTDEncoding: the sorted fast path needs 16-byte aligned tables");
    return 0;
}

// Finest resolution whose cells still hold about one point each (res^3 <= NR3D_MERGE_DENSITY * points per scene): the backward tries to merge
// same-cell points up to there.  The cell key has 10 bits per axis, so never beyond 1024.
static uint32_t merge_res_for(uint64_t N, uint32_t n_scenes) {
    if (NR3D_MERGE_DENSITY == 0) return 1024u;
    const double cells = (double)NR3D_MERGE_DENSITY * (double)N / (double)(n_scenes ? n_scenes : 1);
    const double r = cbrt(cells);
    return r >= 1024.0 ? 1024u : (uint32_t)r;
}

// shared memory of the backward kernel: CTA tiles + one staged [16, 32] row tile per warp
constexpr size_t kBwdSmem = sizeof(float) * ((NR3D_BWD_TILES ? kTileFloats : 0) + kBwdWarps * 16 * kPairRowStride);

template <typename PT, int F, bool SECOND, bool HEAD = false>
static int launch_bwd(const LotdTable& tab, const FastIn& in, const void* dL_dy, int64_t gs_n, int64_t gs_f, const float* ddx, void* grad, cudaStream_t st,
                      const HeadBwd hd = HeadBwd{}) {
    auto kern = lotd_pair_bwd_kernel<PT, F, SECOND, HEAD>;
    static bool configured = false;   // per instantiation; racing threads would set the same value
    if (!configured) {
        NR3D_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem) == cudaSuccess, "lotd_fast_bwd: cannot reserve %d bytes of shared memory", (int)kBwdSmem);
        configured = true;
    }
    const unsigned grid = (unsigned)div_up<uint64_t>(2 * in.N, kBwdThreads);
    kern<<<grid, kBwdThreads, kBwdSmem, st>>>(tab, in, (const PT*)dL_dy, gs_n, gs_f, ddx, (PT*)grad, hd);
    NR3D_LAUNCH_CHECK(SECOND ? "lotd_fast_bwd2" : "lotd_fast_bwd");
    return 0;
}

template <typename PT, int F>
static int launch_fwd(const LotdTable& tab, const FastIn& in, void* y, int64_t ys_n, int64_t ys_f, cudaStream_t st) {
    const unsigned grid = (unsigned)div_up<uint64_t>(2 * in.N, kFastThreads);
#if NR3D_FWD_TMA
    if constexpr (std::is_same<PT, float>::value && F == 2) {
        if (in.scenes == nullptr && ys_f == 1) {
            lotd_pair_fwd_tma_kernel<<<grid, kFastThreads, 0, st>>>(tab, in, (float*)y, ys_n);
            NR3D_LAUNCH_CHECK("lotd_fast_fwd_tma");
            return 0;
        }
    }
#endif
    lotd_pair_fwd_kernel<PT, F><<<grid, kFastThreads, 0, st>>>(tab, in, (PT*)y, ys_n, ys_f, HeadFwd{});
    NR3D_LAUNCH_CHECK("lotd_fast_fwd");
    return 0;
}

template <typename PT, int F>
static int launch_fwd_head(const LotdTable& tab, const FastIn& in, const HeadFwd& hd, cudaStream_t st) {
    const unsigned grid = (unsigned)div_up<uint64_t>(2 * in.N, kFastThreads);
    lotd_pair_fwd_kernel<PT, F, true><<<grid, kFastThreads, 0, st>>>(tab, in, (PT*)nullptr, 0, 0, hd);
    NR3D_LAUNCH_CHECK("lotd_fast_fwd_head");
    return 0;
}

template <typename PT, int F>
static int launch_fwd_dydx(const LotdTable& tab, const FastIn& in, void* y, float* dy_dx, cudaStream_t st) {
    const unsigned grid = (unsigned)div_up<uint64_t>(2 * in.N, kDydxThreads);
    lotd_pair_fwd_dydx_kernel<PT, F><<<grid, kDydxThreads, 0, st>>>(tab, in, (PT*)y, dy_dx);
    NR3D_LAUNCH_CHECK("lotd_fast_fwd_dydx");
    return 0;
}

#define NR3D_DISPATCH_PT_F(PDT, FPL, CALL)                                                    \
    do {                                                                                      \
        if ((PDT) == NR3D_F16) {                                                              \
            if ((FPL) == 2) { using PT = __half; constexpr int F = 2; CALL; }                 \
            else if ((FPL) == 4) { using PT = __half; constexpr int F = 4; CALL; }            \
            else { using PT = __half; constexpr int F = 8; CALL; }                            \
        } else {                                                                              \
            if ((FPL) == 2) { using PT = float; constexpr int F = 2; CALL; }                  \
            else if ((FPL) == 4) { using PT = float; constexpr int F = 4; CALL; }             \
            else { using PT = float; constexpr int F = 8; CALL; }                             \
        }                                                                                     \
    } while (0)

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_lotd_fwd_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                         const void* params, int32_t max_level, void* y, int64_t y_stride_n, int64_t y_stride_f, void* stream) {
    if (int rc = check_fast(meta, param_dtype, N, n_scenes, params)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && params && y, "LoTDEncoding::fwd_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), scenes, params, max_level, meta->n_params, 0u, meta->n_pseudo_levels};
    fast_levels(meta, in);
    int rc = 0;
    NR3D_DISPATCH_PT_F(param_dtype, meta->n_feat_per_pseudo_lvl, (rc = launch_fwd<PT, F>(tab, in, y, y_stride_n, y_stride_f, (cudaStream_t)stream)));
    return rc;
}

int nr3d_lotd_bwd_param_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                               const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f, int32_t max_level, uint32_t pl_begin, uint32_t pl_end,
                               void* dL_dparam, void* stream) {
    if (int rc = check_fast(meta, param_dtype, N, n_scenes, dL_dparam)) return rc;
    if (pl_begin >= pl_end) return 0;
    if (N == 0) return 0;
    NR3D_CHECK(xs && dL_dy && dL_dparam, "LoTDEncoding::bwd_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    const uint32_t pl_stop = (pl_end < meta->n_pseudo_levels) ? pl_end : meta->n_pseudo_levels;
    FastIn in{N, reinterpret_cast<const float4*>(xs), scenes, nullptr, max_level, meta->n_params, pl_begin, pl_stop, merge_res_for(N, n_scenes)};
    fast_levels(meta, in);
    int rc = 0;
    NR3D_DISPATCH_PT_F(param_dtype, meta->n_feat_per_pseudo_lvl,
                       (rc = launch_bwd<PT, F, false>(tab, in, dL_dy, dLdy_stride_n, dLdy_stride_f, nullptr, dL_dparam, (cudaStream_t)stream)));
    return rc;
}

int nr3d_lotd_density_head_fwd_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                                      const void* params, int32_t max_level, const float* deltas, float gain, float* sigma, float* alpha, void* stream) {
    if (int rc = check_fast(meta, param_dtype, N, n_scenes, params)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && params && deltas && sigma && alpha, "LoTDEncoding::density_head_fwd_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), scenes, params, max_level, meta->n_params, 0u, meta->n_pseudo_levels};
    fast_levels(meta, in);
    const HeadFwd hd{deltas, gain, sigma, alpha};
    int rc = 0;
    NR3D_DISPATCH_PT_F(param_dtype, meta->n_feat_per_pseudo_lvl, (rc = launch_fwd_head<PT, F>(tab, in, hd, (cudaStream_t)stream)));
    return rc;
}

int nr3d_lotd_density_head_bwd_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                                      const float* d_alpha, const float* sigma, const float* alpha, const float* deltas, float gain, int32_t max_level,
                                      void* dL_dparam, void* stream) {
    if (int rc = check_fast(meta, param_dtype, N, n_scenes, dL_dparam)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && d_alpha && sigma && alpha && deltas && dL_dparam, "LoTDEncoding::density_head_bwd_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), scenes, nullptr, max_level, meta->n_params, 0u, meta->n_pseudo_levels, merge_res_for(N, n_scenes)};
    fast_levels(meta, in);
    const HeadBwd hd{d_alpha, sigma, alpha, deltas, gain};
    int rc = 0;
    NR3D_DISPATCH_PT_F(param_dtype, meta->n_feat_per_pseudo_lvl,
                       (rc = launch_bwd<PT, F, false, true>(tab, in, nullptr, 0, 0, nullptr, dL_dparam, (cudaStream_t)stream, hd)));
    return rc;
}

int nr3d_lotd_fwd_dydx_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                              const void* params, int32_t max_level, void* y, void* dy_dx, void* stream) {
    if (int rc = check_fast(meta, param_dtype, N, n_scenes, params)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && params && y && dy_dx, "LoTDEncoding::fwd_dydx_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    FastIn in{N, reinterpret_cast<const float4*>(xs), scenes, params, max_level, meta->n_params, 0u, meta->n_pseudo_levels};
    fast_levels(meta, in);
    int rc = 0;
    NR3D_DISPATCH_PT_F(param_dtype, meta->n_feat_per_pseudo_lvl, (rc = launch_fwd_dydx<PT, F>(tab, in, y, (float*)dy_dx, (cudaStream_t)stream)));
    return rc;
}

int nr3d_lotd_bwd_param2_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                                const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f, const float* dL_ddLdx, int32_t max_level,
                                void* dL_dparam, void* stream) {
    const uint32_t pl_begin = 0, pl_end = meta ? meta->n_pseudo_levels : 0;
    if (int rc = check_fast(meta, param_dtype, N, n_scenes, dL_dparam)) return rc;
    if (N == 0) return 0;
    NR3D_CHECK(xs && dL_dy && dL_ddLdx && dL_dparam, "LoTDEncoding::bwd_param2_sorted: null argument");
    LotdTable tab;
    make_table(meta, tab);
    const uint32_t pl_stop = (pl_end < meta->n_pseudo_levels) ? pl_end : meta->n_pseudo_levels;
    FastIn in{N, reinterpret_cast<const float4*>(xs), scenes, nullptr, max_level, meta->n_params, pl_begin, pl_stop, merge_res_for(N, n_scenes)};
    fast_levels(meta, in);
    int rc = 0;
    NR3D_DISPATCH_PT_F(param_dtype, meta->n_feat_per_pseudo_lvl,
                       (rc = launch_bwd<PT, F, true>(tab, in, dL_dy, dLdy_stride_n, dLdy_stride_f, dL_ddLdx, dL_dparam, (cudaStream_t)stream)));
    return rc;
}

}  // extern "C"
