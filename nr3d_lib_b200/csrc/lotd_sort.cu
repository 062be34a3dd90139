// lotd_sort.cu -- counting sort of the query points by (scene, cell bin) for the cell-sorted LoTD fast path (lotd_fast.cu), with a
// device-side check of whether the records of an earlier call are still current.
//
// Why the check lives on the device: lod_fwd and lod_bwd of one training step see the same points, so the sort should run once per step.
// The reference's autograd wrappers (nr3d_lib/models/grid_encodings/lotd/lotd.py:60-119) hand the same tensor to both calls but offer no
// way to pass our sorted records along.  A host-side cache keyed on the tensor's address / version counter is blind to writes that do
// not bump the counter (`x.data.mul_()`, raw kernels, dlpack aliases).  So every call fingerprints the points it was given (one pass over
// x, 12 bytes per point) and compares with the fingerprint stored next to the records; the sort kernels return at once when it matches.
// Nothing is read back to the host; everything is ordered on the caller's stream.
//
//   verify   1 launch   64-bit sum + xor of a per-point hash of (x, y, z, index, scene); last block: compare, set `skip`, clear the scan state
//                       (skipped for `force` calls -- forward passes, whose points are new: the histogram pass takes the fingerprint instead)
//   hist     1 launch   bin key -> rank inside the bin (atomic counter), 4 bytes per point kept
//   scan     1 launch   decoupled look-back exclusive scan of the bin counters; zeroes the counters for the next call
//   scatter  1 launch   16-byte records (x, y, z, original index) [+ uint16 scene] written at offsets[key] + rank
#include "lotd_pair.cuh"
#include <string.h>

namespace nr3d {

#ifndef NR3D_BIN_RES        // 0: chosen per call from the number of points (about two points per bin), else fixed (A/B runs)
#define NR3D_BIN_RES 0
#endif

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;                        // per thread
constexpr int kScanTile = kScanThreads * kScanItems;   // counters per tile
constexpr uint32_t kMaxTiles = 8192;                 // 2^25 counters
constexpr uint64_t kHeaderBytes = 256;

struct SortHeader {
    unsigned long long fp_sum, fp_xor, fp_n;      // fingerprint (and point count / scene configuration) the records in `xs` were built from
    unsigned long long acc_sum, acc_xor;          // accumulators of the running verify pass (zero between calls)
    uint32_t blocks_done;                         // last-block detection (zero between calls)
    uint32_t skip;                                // verdict of the last verify pass: 1 = records are current
    uint32_t tile_counter;                        // scan: dynamic tile ids (zero at scan start)
    uint32_t pad;
};

// bins per axis: about two points per bin (A/B on B200: 4 Mi uniform points 128^3 > 64^3, 256^3; 30 Mi ray samples 256^3 > 192^3 > 128^3,
// profiles/r1_ab_tunables.txt); with several scenes every scene gets its own bin grid.
static inline uint32_t bin_res_for(uint64_t N, uint32_t n_scenes) {
    if (NR3D_BIN_RES) return NR3D_BIN_RES;
    const uint64_t per = N / (n_scenes ? n_scenes : 1);
    if (per >= (12ull << 20)) return 256u;
    if (per >= (1536ull << 10)) return 128u;
    if (per >= (192ull << 10)) return 64u;
    return 32u;
}

// Optional map applied to every coordinate before binning and before it is stored in the record: x' = clamp(fma(x, scale, shift)), so that
// callers holding points in another box (ray samples in [-1, 1]^3: scale = shift = 0.5) need no separate passes over the [N, 3] array
// (reference: LoTDEncoding.forward normalises, lotd_encoding.py:162, LoTD.forward clamps, lotd.py:211).  Identity: scale 1, shift 0, no clamp.
struct SortMap {
    float scale, shift;
    int32_t clamp01;   // 1: clamp to [1e-6, 1 - 1e-6] like `x.clamp(1e-6, 1 - 1e-6)` (same float32 constants)
    __device__ __forceinline__ float operator()(float v) const {
        v = fmaf(v, scale, shift);
        return clamp01 ? fminf(fmaxf(v, 1.0e-6f), (float)(1.0 - 1.0e-6)) : v;
    }
};

__device__ __forceinline__ uint32_t bin_key(float x, float y, float z, uint32_t res) {
    const uint32_t bx = min(res - 1, (uint32_t)fmaxf(x * (float)res, 0.f));
    const uint32_t by = min(res - 1, (uint32_t)fmaxf(y * (float)res, 0.f));
    const uint32_t bz = min(res - 1, (uint32_t)fmaxf(z * (float)res, 0.f));
    return bin_order(bx, by, bz, res);
}

// scene of point i: 0 for single-scene calls, 0xffff for skipped points (batch_inds < 0 or out of range)
__device__ __forceinline__ uint32_t scene_of(uint64_t i, const int64_t* __restrict__ batch_inds, uint32_t bds, uint32_t n_scenes) {
    if (batch_inds) {
        const int64_t b = __ldg(batch_inds + i);
        return (b < 0 || b >= (int64_t)n_scenes) ? 0xffffu : (uint32_t)b;
    }
    if (bds) {
        const uint64_t b = i / bds;
        return b >= n_scenes ? 0xffffu : (uint32_t)b;
    }
    return 0u;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned long long point_hash(const float* __restrict__ x, uint64_t i, uint32_t sc) {
    const uint32_t a = __float_as_uint(__ldcs(x + i * 3)), b = __float_as_uint(__ldcs(x + i * 3 + 1)), c = __float_as_uint(__ldcs(x + i * 3 + 2));
    return mix64((((unsigned long long)a << 32) | b) ^ mix64((((unsigned long long)c << 32) | (uint32_t)i) + ((unsigned long long)sc << 48) + 0x9e3779b97f4a7c15ull));
}

// Block-level end of a fingerprint pass: adds the block's partial sums to the header; the LAST block compares with the stored fingerprint,
// stores the new one, sets `skip` and re-arms the scan state.  Returns (to every thread of the last block) whether it was the last block.
__device__ __forceinline__ bool fingerprint_finish(unsigned long long sum, unsigned long long xr, uint64_t N, uint32_t n_scenes, uint32_t res, int force,
                                                   SortHeader* __restrict__ hdr, unsigned long long* __restrict__ status, uint32_t n_tiles, uint32_t map_tag) {
    __shared__ unsigned long long s_sum, s_xor;
    __shared__ bool s_last, s_same;
    if (threadIdx.x == 0) { s_sum = 0; s_xor = 0; }
    __syncthreads();
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, m);
        xr ^= __shfl_xor_sync(0xffffffffu, xr, m);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_sum, sum); atomicXor(&s_xor, xr); }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&hdr->acc_sum, s_sum);
        atomicXor(&hdr->acc_xor, s_xor);
        __threadfence();
        s_last = atomicAdd(&hdr->blocks_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (threadIdx.x == 0) {
        const unsigned long long fs = atomicAdd(&hdr->acc_sum, 0ull), fx = atomicAdd(&hdr->acc_xor, 0ull);
        // (the records also depend on the coordinate map they were built with)
        const unsigned long long fn = N ^ ((unsigned long long)n_scenes << 40) ^ ((unsigned long long)res << 52) ^ mix64((unsigned long long)map_tag);
        const bool same = !force && fs == hdr->fp_sum && fx == hdr->fp_xor && fn == hdr->fp_n;
        hdr->fp_sum = fs; hdr->fp_xor = fx; hdr->fp_n = fn;
        hdr->skip = same ? 1u : 0u;
        hdr->acc_sum = 0; hdr->acc_xor = 0; hdr->blocks_done = 0; hdr->tile_counter = 0;
        s_same = same;
    }
    __syncthreads();
    if (!s_same)
        for (uint32_t t = threadIdx.x; t < n_tiles; t += blockDim.x) status[t] = 0ull;
    return true;
}

__global__ void __launch_bounds__(256) sort_verify_kernel(uint64_t N, const float* __restrict__ x, const int64_t* __restrict__ batch_inds, uint32_t bds,
                                                          uint32_t n_scenes, uint32_t res, int force, SortHeader* __restrict__ hdr,
                                                          unsigned long long* __restrict__ status, uint32_t n_tiles, uint32_t map_tag) {
    unsigned long long sum = 0, xr = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long h = point_hash(x, i, scene_of(i, batch_inds, bds, n_scenes));
        sum += h; xr ^= h;
    }
    fingerprint_finish(sum, xr, N, n_scenes, res, force, hdr, status, n_tiles, map_tag);
}

// pass 1: rank of the point inside its bin (the rank makes the scatter pass atomic-free)
// FP = true ("the points are new", forward calls): no verify pass ran; this kernel sorts unconditionally and records the fingerprint of the
// points on the way (same pass over x), so that a later verify pass -- the backward of the same step -- finds the records current.
template <bool FP>
__global__ void __launch_bounds__(256) sort_hist_kernel(uint64_t N, uint32_t res, uint32_t n_scenes, const float* __restrict__ x,
                                                        const int64_t* __restrict__ batch_inds, uint32_t bds, SortHeader* __restrict__ hdr,
                                                        uint32_t* __restrict__ hist, uint32_t* __restrict__ rank,
                                                        unsigned long long* __restrict__ status, uint32_t n_tiles, const SortMap map, uint32_t map_tag) {
    if (!FP && hdr->skip) return;
    const uint32_t bins = res * res * res;
    unsigned long long sum = 0, xr = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t sc = scene_of(i, batch_inds, bds, n_scenes);
        const uint32_t k = sc == 0xffffu ? n_scenes * bins : sc * bins + bin_key(map(x[i * 3]), map(x[i * 3 + 1]), map(x[i * 3 + 2]), res);
        rank[i] = atomicAdd(hist + k, 1u);
        if (FP) { const unsigned long long h = point_hash(x, i, sc); sum += h; xr ^= h; }
    }
    if (FP) fingerprint_finish(sum, xr, N, n_scenes, res, /*force=*/1, hdr, status, n_tiles, map_tag);
}

// single-pass exclusive scan (decoupled look-back): hist -> offsets; the counters are zeroed on the way for the next call
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kFlagMask = 3ull << 62;
__global__ void __launch_bounds__(kScanThreads) sort_scan_kernel(uint32_t n, SortHeader* __restrict__ hdr, uint32_t* __restrict__ hist,
                                                                 uint32_t* __restrict__ offsets, unsigned long long* __restrict__ status) {
    if (hdr->skip) return;
    __shared__ uint32_t s_tile, s_prefix, ws[32];
    if (threadIdx.x == 0) s_tile = atomicAdd(&hdr->tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    if (base + kScanItems <= n) {
        const uint4 t = *reinterpret_cast<const uint4*>(hist + base);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        *reinterpret_cast<uint4*>(hist + base) = make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = base + k < n ? hist[base + k] : 0u;
            if (base + k < n) hist[base + k] = 0u;
        }
    }
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t s = mine;
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
    const uint32_t excl = ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0u) + s - mine;  // exclusive prefix of this thread inside the tile
    if (threadIdx.x < 32) {   // warp 0: publish the tile aggregate, then look back 32 predecessors at a time until an inclusive prefix shows up
        const uint32_t lane = threadIdx.x;
        const uint32_t total = ws[31];
        uint32_t run = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(status, kFlagPrefix | total);
        } else {
            if (lane == 0) atomicExch(status + tile, kFlagAgg | total);
            for (int j0 = (int)tile - 1; ; j0 -= 32) {
                const int j = j0 - (int)lane;
                unsigned long long st = kFlagPrefix;             // before tile 0: an (empty) inclusive prefix
                if (j >= 0) {
                    do { st = *reinterpret_cast<volatile unsigned long long*>(status + j); } while ((st & kFlagMask) == 0ull);
                }
                const uint32_t pmask = __ballot_sync(0xffffffffu, (st & kFlagMask) == kFlagPrefix);
                const int first = pmask ? (__ffs(pmask) - 1) : 31;   // nearest predecessor that already holds an inclusive prefix
                uint32_t v = ((int)lane <= first) ? (uint32_t)st : 0u;
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                run += v;
                if (pmask) break;
            }
            if (lane == 0) atomicExch(status + tile, kFlagPrefix | (unsigned long long)(run + total));
        }
        if (lane == 0) s_prefix = run;
    }
    __syncthreads();
    uint32_t o = s_prefix + excl;
    if (base + kScanItems <= n) {
        uint4 t;
        t.x = o; o += v[0]; t.y = o; o += v[1]; t.z = o; o += v[2]; t.w = o;
        *reinterpret_cast<uint4*>(offsets + base) = t;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) { if (base + k < n) offsets[base + k] = o; o += v[k]; }
    }
}

// pass 2: sorted record = (x, y, z, original index) written with ONE 16-byte store per point (+ the scene for batched calls)
__global__ void __launch_bounds__(256) sort_scatter_kernel(uint64_t N, uint32_t res, uint32_t n_scenes, const float* __restrict__ x,
                                                           const int64_t* __restrict__ batch_inds, uint32_t bds, const SortHeader* __restrict__ hdr,
                                                           const uint32_t* __restrict__ rank, const uint32_t* __restrict__ offsets,
                                                           float4* __restrict__ xs, uint16_t* __restrict__ scenes, const SortMap map) {
    if (hdr->skip) return;
    const uint32_t bins = res * res * res;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const float px = map(x[i * 3]), py = map(x[i * 3 + 1]), pz = map(x[i * 3 + 2]);
        const uint32_t sc = scene_of(i, batch_inds, bds, n_scenes);
        const uint32_t k = sc == 0xffffu ? n_scenes * bins : sc * bins + bin_key(px, py, pz, res);
        const uint32_t pos = __ldg(offsets + k) + __ldcs(rank + i);
        xs[pos] = make_float4(px, py, pz, __uint_as_float((uint32_t)i));
        if (scenes) scenes[pos] = (uint16_t)sc;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Two-level variant for large single-scene calls (>= kTwoLevelMin points: the ray-sample chunks of the M2 workload, 30 Mi samples over 256^3
// bins).  There the one-level sort is bound by RANDOM traffic: the histogram's atomics and the scatter's offset gathers miss L2 in a 64 MB
// counter table, and the 16-byte records land on random sectors of a 480 MB array (partial-sector writes: DRAM read-modify-write).
// Level 1 partitions the points into 16^3 coarse buckets with CTA-local staging -- a tile of 4096 points counts its buckets in shared
// memory, reserves ONE range per (tile, bucket) with a single global atomic and writes each bucket's records back to back (ray samples:
// runs of ~12 records) --, level 2 runs the fine sort over that coarse-sorted copy, where consecutive records share their counter lines and
// their output segments, so atomics, gathers and record writes all combine in L2.  The final order is the same x-fastest bin order.
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t kCoarse = 16, kCoarseN = kCoarse * kCoarse * kCoarse;
constexpr int kCoarseTile = 4096;            // points per tile (256 threads x 16)
static uint64_t g_two_level_min = 12ull << 20;

__device__ __forceinline__ uint32_t coarse_key(float x, float y, float z, uint32_t res) {   // (x, y, z) already mapped; res is a multiple of 16
    const uint32_t per = res / kCoarse;
    const uint32_t bx = min(res - 1, (uint32_t)fmaxf(x * (float)res, 0.f)) / per;
    const uint32_t by = min(res - 1, (uint32_t)fmaxf(y * (float)res, 0.f)) / per;
    const uint32_t bz = min(res - 1, (uint32_t)fmaxf(z * (float)res, 0.f)) / per;
    return (bz * kCoarse + by) * kCoarse + bx;
}

template <bool FP>
__global__ void __launch_bounds__(256) sort_coarse_count_kernel(uint64_t N, uint32_t res, const float* __restrict__ x, SortHeader* __restrict__ hdr,
                                                                uint32_t* __restrict__ coarse_cnt, unsigned long long* __restrict__ status, uint32_t n_tiles,
                                                                const SortMap map, uint32_t map_tag) {
    if (!FP && hdr->skip) return;
    __shared__ uint32_t sh[kCoarseN];
    unsigned long long sum = 0, xr = 0;
    const uint64_t tiles = (N + kCoarseTile - 1) / kCoarseTile;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < kCoarseN; b += 256) sh[b] = 0;
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kCoarseTile / 256; ++j) {
            const uint64_t i = tile * kCoarseTile + (uint64_t)j * 256 + threadIdx.x;
            if (i < N) {
                atomicAdd(&sh[coarse_key(map(x[i * 3]), map(x[i * 3 + 1]), map(x[i * 3 + 2]), res)], 1u);
                if (FP) { const unsigned long long h = point_hash(x, i, 0u); sum += h; xr ^= h; }
            }
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < kCoarseN; b += 256) { const uint32_t c = sh[b]; if (c) atomicAdd(coarse_cnt + b, c); }
        __syncthreads();
    }
    if (FP) fingerprint_finish(sum, xr, N, 1u, res, /*force=*/1, hdr, status, n_tiles, map_tag);
}

// exclusive scan of the 4096 bucket sizes -> first record of every bucket (also the running cursor of the partition pass); re-zeroes the sizes
__global__ void __launch_bounds__(1024) sort_coarse_scan_kernel(const SortHeader* __restrict__ hdr, uint32_t* __restrict__ coarse_cnt, uint32_t* __restrict__ cursor) {
    if (hdr->skip) return;
    __shared__ uint32_t ws[32];
    const uint32_t t = threadIdx.x;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = coarse_cnt[t * 4 + k]; coarse_cnt[t * 4 + k] = 0; }
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t s = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, s, d); if ((t & 31) >= (uint32_t)d) s += u; }
    if ((t & 31) == 31) ws[t >> 5] = s;
    __syncthreads();
    if (t < 32) {
        uint32_t w = ws[t];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, w, d); if (t >= (uint32_t)d) w += u; }
        ws[t] = w;
    }
    __syncthreads();
    uint32_t o = ((t >> 5) ? ws[(t >> 5) - 1] : 0u) + s - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) { cursor[t * 4 + k] = o; o += v[k]; }
}

__global__ void __launch_bounds__(256) sort_coarse_partition_kernel(uint64_t N, uint32_t res, const float* __restrict__ x, const SortHeader* __restrict__ hdr,
                                                                    uint32_t* __restrict__ cursor, float4* __restrict__ tmp, const SortMap map) {
    if (hdr->skip) return;
    __shared__ uint32_t cnt[kCoarseN], base[kCoarseN];
    const uint64_t tiles = (N + kCoarseTile - 1) / kCoarseTile;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < kCoarseN; b += 256) cnt[b] = 0;
        __syncthreads();
        uint32_t kr[kCoarseTile / 256];      // (bucket << 16) | rank inside (tile, bucket)   -- a tile holds 4096 points: 12 bits suffice for both
#pragma unroll
        for (int j = 0; j < kCoarseTile / 256; ++j) {
            const uint64_t i = tile * kCoarseTile + (uint64_t)j * 256 + threadIdx.x;
            kr[j] = 0xffffffffu;
            if (i < N) {
                const uint32_t ck = coarse_key(map(x[i * 3]), map(x[i * 3 + 1]), map(x[i * 3 + 2]), res);
                kr[j] = (ck << 16) | atomicAdd(&cnt[ck], 1u);
            }
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < kCoarseN; b += 256) { const uint32_t c = cnt[b]; if (c) base[b] = atomicAdd(cursor + b, c); }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kCoarseTile / 256; ++j) {
            const uint64_t i = tile * kCoarseTile + (uint64_t)j * 256 + threadIdx.x;
            if (kr[j] != 0xffffffffu)
                tmp[base[kr[j] >> 16] + (kr[j] & 0xffffu)] = make_float4(map(x[i * 3]), map(x[i * 3 + 1]), map(x[i * 3 + 2]), __uint_as_float((uint32_t)i));
        }
        __syncthreads();
    }
}

// level 2 over the coarse-sorted records (coordinates already mapped)
__global__ void __launch_bounds__(256) sort_hist_rec_kernel(uint64_t N, uint32_t res, const float4* __restrict__ tmp, const SortHeader* __restrict__ hdr,
                                                            uint32_t* __restrict__ hist, uint32_t* __restrict__ rank) {
    if (hdr->skip) return;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 r = tmp[i];
        rank[i] = atomicAdd(hist + bin_key(r.x, r.y, r.z, res), 1u);
    }
}
__global__ void __launch_bounds__(256) sort_scatter_rec_kernel(uint64_t N, uint32_t res, const float4* __restrict__ tmp, const SortHeader* __restrict__ hdr,
                                                               const uint32_t* __restrict__ rank, const uint32_t* __restrict__ offsets, float4* __restrict__ xs) {
    if (hdr->skip) return;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 r = __ldcs(tmp + i);
        xs[__ldg(offsets + bin_key(r.x, r.y, r.z, res)) + __ldcs(rank + i)] = r;
    }
}

// Workspace layout of a call: [header | scan status | hist | offsets | rank | (two-level: tmp records | coarse sizes | coarse cursors)].
// The header, the histogram counters and the coarse bucket sizes are STATE: every call expects them zero and leaves them zero.
struct SortLayout {
    uint32_t res, n_cnt, n_tiles;
    uint64_t cnt_bytes, rank_bytes, need;
    bool two_level;
};

static int sort_layout(uint64_t N, bool has_batch_inds, uint32_t batch_data_size, uint32_t n_scenes, SortLayout& l) {
    NR3D_CHECK(n_scenes < 0xffffu, "sort_points: at most 65534 scenes");
    l.res = bin_res_for(N, n_scenes);
    const uint64_t n_cnt64 = (uint64_t)n_scenes * l.res * l.res * l.res + 1;   // one extra bucket for skipped points
    NR3D_CHECK(n_cnt64 <= (uint64_t)kMaxTiles * kScanTile, "sort_points: too many scenes for the bin grid (%llu counters)", (unsigned long long)n_cnt64);
    l.n_cnt = (uint32_t)n_cnt64;
    l.n_tiles = div_up<uint32_t>(l.n_cnt, kScanTile);
    l.cnt_bytes = div_up<uint64_t>((uint64_t)l.n_cnt * 4, 256) * 256;
    l.two_level = N >= g_two_level_min && n_scenes == 1 && !has_batch_inds && batch_data_size == 0 && (l.res % kCoarse) == 0;
    l.rank_bytes = div_up<uint64_t>(N * 4, 256) * 256;
    l.need = kHeaderBytes + (uint64_t)kMaxTiles * 8 + 2 * l.cnt_bytes + l.rank_bytes + (l.two_level ? N * 16 + 2 * kCoarseN * 4 : 0);
    return 0;
}

}  // namespace nr3d

using namespace nr3d;

extern "C" {

// Puts the stateful parts of a workspace into their initial (all-zero) state for calls with this (N, batching) configuration, on `stream`:
// what `zero-fill before the first use` asks for, without touching the hundreds of megabytes of rank / record scratch behind them.  Callers
// that keep ONE buffer for point counts that change from call to call (ray samples: every step has its own count) call it whenever the
// configuration changes instead of allocating and zero-filling a new workspace.
int nr3d_lotd_sort_ws_reset(uint64_t N, int32_t has_batch_inds, uint32_t batch_data_size, uint32_t n_scenes, void* ws, uint64_t ws_bytes, void* stream) {
    if (n_scenes == 0) n_scenes = 1;
    SortLayout l;
    if (int rc = sort_layout(N, has_batch_inds != 0, batch_data_size, n_scenes, l)) return rc;
    NR3D_CHECK(ws != nullptr && ws_bytes >= l.need, "sort_points: workspace too small");
    NR3D_CHECK((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "sort_points: ws must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char* w = reinterpret_cast<char*>(ws);
    NR3D_CHECK(cudaMemsetAsync(w, 0, kHeaderBytes + (uint64_t)kMaxTiles * 8 + l.cnt_bytes, st) == cudaSuccess, "sort_points: memset failed");
    if (l.two_level) {
        char* coarse = w + kHeaderBytes + (uint64_t)kMaxTiles * 8 + 2 * l.cnt_bytes + l.rank_bytes + N * 16;
        NR3D_CHECK(cudaMemsetAsync(coarse, 0, 2 * kCoarseN * 4, st) == cudaSuccess, "sort_points: memset failed");
    }
    return 0;
}

int nr3d_lotd_sort_points_mapped(uint64_t N, const float* x, const int64_t* batch_inds, uint32_t batch_data_size, uint32_t n_scenes, int32_t force,
                                 float scale, float shift, int32_t clamp01, void* xs /* float4 [N] */, uint16_t* scenes /* [N] or NULL */, void* ws,
                                 uint64_t* ws_bytes, void* stream) {
    const SortMap map{scale, shift, clamp01};
    uint32_t su, hu;
    memcpy(&su, &scale, 4); memcpy(&hu, &shift, 4);
    const uint32_t map_tag = (su * 0x9E3779B1u) ^ (hu * 0x85EBCA77u) ^ (clamp01 ? 0x27D4EB2Fu : 0u);
    if (n_scenes == 0) n_scenes = 1;
    SortLayout l;
    if (int rc = sort_layout(N, batch_inds != nullptr, batch_data_size, n_scenes, l)) return rc;
    const uint32_t res = l.res, n_cnt = l.n_cnt, n_tiles = l.n_tiles;
    const uint64_t cnt_bytes = l.cnt_bytes, rank_bytes = l.rank_bytes, need = l.need;
    const bool two_level = l.two_level;
    if (ws == nullptr) {
        NR3D_CHECK(ws_bytes != nullptr, "sort_points: null ws_bytes");
        *ws_bytes = need;
        return 0;
    }
    NR3D_CHECK(ws_bytes && *ws_bytes >= need, "sort_points: workspace too small");
    NR3D_CHECK(N < (1ull << 32), "sort_points: N must be < 2^32");
    if (N == 0) return 0;
    NR3D_CHECK(x && xs, "sort_points: null argument");
    NR3D_CHECK((batch_inds == nullptr && batch_data_size == 0 && n_scenes == 1) || scenes != nullptr, "sort_points: batched calls need the `scenes` output");
    NR3D_CHECK((reinterpret_cast<uintptr_t>(xs) & 15u) == 0 && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "sort_points: xs must be 16-byte, ws 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char* w = reinterpret_cast<char*>(ws);
    SortHeader* hdr = reinterpret_cast<SortHeader*>(w);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(w + kHeaderBytes);
    uint32_t* hist = reinterpret_cast<uint32_t*>(w + kHeaderBytes + (uint64_t)kMaxTiles * 8);
    uint32_t* offsets = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(hist) + cnt_bytes);
    uint32_t* rank = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(offsets) + cnt_bytes);
    // a few CTAs per SM with grid-stride loops: when the records are current, the three sort kernels cost an (almost) empty launch each
    const uint64_t gwant = div_up<uint64_t>(N, 256);
    const unsigned grid = (unsigned)(gwant < (uint64_t)kSMs * 16 ? gwant : (uint64_t)kSMs * 16);
    const uint64_t vwant = div_up<uint64_t>(N, 256 * 8);
    const unsigned vgrid = (unsigned)(vwant < (uint64_t)kSMs * 8 ? vwant : (uint64_t)kSMs * 8);
    if (two_level) {
        float4* tmp = reinterpret_cast<float4*>(reinterpret_cast<char*>(rank) + rank_bytes);
        uint32_t* coarse_cnt = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(tmp) + N * 16);
        uint32_t* cursor = coarse_cnt + kCoarseN;
        const uint64_t twant = div_up<uint64_t>(N, kCoarseTile);
        const unsigned tgrid = (unsigned)(twant < (uint64_t)kSMs * 6 ? twant : (uint64_t)kSMs * 6);
        if (force) {
            sort_coarse_count_kernel<true><<<tgrid, 256, 0, st>>>(N, res, x, hdr, coarse_cnt, status, n_tiles, map, map_tag);
        } else {
            sort_verify_kernel<<<vgrid, 256, 0, st>>>(N, x, nullptr, 0, 1u, res, 0, hdr, status, n_tiles, map_tag);
            NR3D_LAUNCH_CHECK("sort_verify");
            sort_coarse_count_kernel<false><<<tgrid, 256, 0, st>>>(N, res, x, hdr, coarse_cnt, status, n_tiles, map, map_tag);
        }
        NR3D_LAUNCH_CHECK("sort_coarse_count");
        sort_coarse_scan_kernel<<<1, 1024, 0, st>>>(hdr, coarse_cnt, cursor);
        NR3D_LAUNCH_CHECK("sort_coarse_scan");
        sort_coarse_partition_kernel<<<tgrid, 256, 0, st>>>(N, res, x, hdr, cursor, tmp, map);
        NR3D_LAUNCH_CHECK("sort_coarse_partition");
        sort_hist_rec_kernel<<<grid, 256, 0, st>>>(N, res, tmp, hdr, hist, rank);
        NR3D_LAUNCH_CHECK("sort_hist_rec");
        sort_scan_kernel<<<n_tiles, kScanThreads, 0, st>>>(n_cnt, hdr, hist, offsets, status);
        NR3D_LAUNCH_CHECK("sort_scan");
        sort_scatter_rec_kernel<<<grid, 256, 0, st>>>(N, res, tmp, hdr, rank, offsets, reinterpret_cast<float4*>(xs));
        NR3D_LAUNCH_CHECK("sort_scatter_rec");
        return 0;
    }
    if (force) {   // new points (forward calls, first use of the buffers): sort unconditionally, the fingerprint is taken inside the histogram pass
        sort_hist_kernel<true><<<grid, 256, 0, st>>>(N, res, n_scenes, x, batch_inds, batch_data_size, hdr, hist, rank, status, n_tiles, map, map_tag);
    } else {       // probably the points of the previous call (the backward of a step): fingerprint first, the sort kernels return at once on a match
        sort_verify_kernel<<<vgrid, 256, 0, st>>>(N, x, batch_inds, batch_data_size, n_scenes, res, 0, hdr, status, n_tiles, map_tag);
        NR3D_LAUNCH_CHECK("sort_verify");
        sort_hist_kernel<false><<<grid, 256, 0, st>>>(N, res, n_scenes, x, batch_inds, batch_data_size, hdr, hist, rank, status, n_tiles, map, map_tag);
    }
    NR3D_LAUNCH_CHECK("sort_hist");
    sort_scan_kernel<<<n_tiles, kScanThreads, 0, st>>>(n_cnt, hdr, hist, offsets, status);
    NR3D_LAUNCH_CHECK("sort_scan");
    sort_scatter_kernel<<<grid, 256, 0, st>>>(N, res, n_scenes, x, batch_inds, batch_data_size, hdr, rank, offsets, reinterpret_cast<float4*>(xs), scenes, map);
    NR3D_LAUNCH_CHECK("sort_scatter");
    return 0;
}

// test / A-B knob: point count from which single-scene calls take the two-level sort (0 restores the default of 12 Mi).  Changing it changes the
// workspace size of a given N: drop cached workspaces afterwards (bindings._lotd.clear_sort_cache()).
int nr3d_lotd_sort_set_two_level_min(uint64_t n_points) {
    g_two_level_min = n_points ? n_points : (12ull << 20);
    return 0;
}

int nr3d_lotd_sort_points(uint64_t N, const float* x, const int64_t* batch_inds, uint32_t batch_data_size, uint32_t n_scenes, int32_t force,
                          void* xs /* float4 [N] */, uint16_t* scenes /* [N] or NULL */, void* ws, uint64_t* ws_bytes, void* stream) {
    return nr3d_lotd_sort_points_mapped(N, x, batch_inds, batch_data_size, n_scenes, force, 1.0f, 0.0f, 0, xs, scenes, ws, ws_bytes, stream);
}

}  // extern "C"
