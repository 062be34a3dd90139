// pack_next.cu -- the hierarchical-sampling pack ops ("next" row n2 of SURVEY.md section 8f): searchsorted, invert_cdf,
// merge of two sorted packs, per-pack sort, per-pack matmul.  They sit between the march and the composite in the
// reference's NeuS / StreetSurf samplers (nr3d_lib/graphics/neus/neus_ray_query.py, graphics/raysample.py).
//
// The reference gives every pack to ONE thread (csrc/pack_ops/pack_ops_cuda.cu:1374-1407, 1505-1571, 1633-1681, 2634-2763,
// 2060-2085).  Every query of searchsorted / invert_cdf is independent, so here a warp owns a pack and its lanes take
// the queries; the merge is three warp-parallel phases (lower bounds, counts + scan, ranks); the sort is a block-wide
// direction-free bitonic network in shared memory (global memory for packs that do not fit).
#include "common.cuh"

namespace nr3d {

constexpr int kNextThreads = 256;
constexpr int kNextWarps = kNextThreads / 32;

template <typename T> __device__ __forceinline__ T clamp_next(T v, T lo, T hi) { return v < lo ? lo : (hi < v ? hi : v); }
constexpr unsigned kFullMask = 0xffffffffu;

template <typename T> struct Num { static __device__ __forceinline__ float f(T v) { return (float)v; } };

// first index with data[idx] >= val  (== binary_search_unsafe, pack_ops_cuda.cu:1336-1362)
template <typename T>
__device__ __forceinline__ uint32_t lower_bound_dev(T val, const T* __restrict__ data, uint32_t length) {
    uint32_t first = 0, count = length;
    while (count > 0) {
        const uint32_t step = count >> 1;
        const uint32_t it = first + step;
        if (data[it] < val) { first = it + 1; count -= step + 1; }
        else count = step;
    }
    return first;
}

struct PR { uint64_t begin; uint32_t len; };
__device__ __forceinline__ PR pack_of(const int64_t* __restrict__ pi, uint64_t p) {
    PR r;
    r.begin = (uint64_t)pi[2 * p];
    const int64_t n = pi[2 * p + 1];
    r.len = n > 0 ? (uint32_t)n : 0u;
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kNextThreads)
searchsorted_kernel(uint64_t P, const T* __restrict__ bins, const int64_t* __restrict__ pack_infos, const T* __restrict__ vals,
                    uint32_t num_to_search, const int64_t* __restrict__ val_pack_infos, int64_t* __restrict__ pidx) {
    const uint64_t p = (uint64_t)blockIdx.x * kNextWarps + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = threadIdx.x & 31;
    const PR b = pack_of(pack_infos, p);
    uint64_t out_begin = p * num_to_search;
    uint32_t nq = num_to_search;
    if (val_pack_infos) { const PR q = pack_of(val_pack_infos, p); out_begin = q.begin; nq = q.len; }
    for (uint32_t i = lane; i < nq; i += 32) {
        uint32_t pos = 0;
        if (b.len) pos = min(lower_bound_dev<T>(vals[out_begin + i], bins + b.begin, b.len), b.len - 1);
        pidx[out_begin + i] = (int64_t)(b.begin + pos);
    }
}

template <typename T>
__global__ void __launch_bounds__(kNextThreads)
invert_cdf_kernel(uint64_t P, const T* __restrict__ bins, const T* __restrict__ cdfs, const int64_t* __restrict__ pack_infos,
                  const T* __restrict__ u_vals, uint32_t num_to_sample, T* __restrict__ samples, int64_t* __restrict__ bin_idx) {
    const uint64_t p = (uint64_t)blockIdx.x * kNextWarps + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = threadIdx.x & 31;
    const PR b = pack_of(pack_infos, p);
    const T eps = (T)1.0e-5f;
    const T* bn = bins + b.begin;
    const T* cd = cdfs + b.begin;
    for (uint32_t i = lane; i < num_to_sample; i += 32) {
        const uint64_t o = p * num_to_sample + i;
        const T u = u_vals[o];
        uint32_t pos = 0;
        if (b.len) pos = min(lower_bound_dev<T>(u, cd, b.len), b.len - 1);
        bin_idx[o] = (int64_t)(pos + b.begin);
        if (pos == 0) {
            samples[o] = bn[0];
        } else {
            const T pmf = cd[pos] - cd[pos - 1];
            samples[o] = pmf < eps ? bn[pos - 1] : (bn[pos - 1] + ((u - cd[pos - 1]) / pmf) * (bn[pos] - bn[pos - 1]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// try_merge_two_packs_sorted_aligned (pack_ops_cuda.cu:1505-1571): positions of a's and b's elements in the merged
// sorted pack.  pidx_a must be zero-filled.  Semantics kept: a[i] goes after every b whose lower bound in a is <= i;
// b's with equal lower bound keep their order when they are consecutive.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kNextThreads)
merge_sorted_kernel(uint64_t P, const T* __restrict__ vals_a, const int64_t* __restrict__ pi_a, const T* __restrict__ vals_b,
                    const int64_t* __restrict__ pi_b, const int64_t* __restrict__ pi_m, int64_t* __restrict__ pidx_a,
                    int64_t* __restrict__ pidx_b) {
    const uint64_t p = (uint64_t)blockIdx.x * kNextWarps + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = threadIdx.x & 31;
    const PR a = pack_of(pi_a, p), b = pack_of(pi_b, p);
    const int64_t out_begin = pi_m[2 * p];
    const T* va = vals_a + a.begin;
    const T* vb = vals_b + b.begin;
    int64_t* pa = pidx_a + a.begin;
    int64_t* pb = pidx_b + b.begin;
    // phase 1: lower bound of every b in a; count per a-slot
    for (uint32_t j = lane; j < b.len; j += 32) {
        const uint32_t i = lower_bound_dev<T>(vb[j], va, a.len);
        pb[j] = (int64_t)i;
        if (i < a.len) atomicAdd(reinterpret_cast<unsigned long long*>(pa + i), 1ull);
    }
    __syncwarp();
    __threadfence_block();
    // phase 2: pidx_a[i] = out_begin + i + inclusive_cumsum(count)[i]
    int64_t carry = 0;
    for (uint32_t base = 0; base < a.len; base += 32) {
        const uint32_t i = base + lane;
        const int64_t c = i < a.len ? __ldcg(reinterpret_cast<const long long*>(pa + i)) : 0;
        int64_t s = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(kFullMask, s, d); if (lane >= d) s += t; }
        if (i < a.len) pa[i] = out_begin + (int64_t)i + carry + s;
        carry += __shfl_sync(kFullMask, s, 31);
    }
    __syncwarp();
    __threadfence_block();
    // phase 3: rank of b inside its run of equal lower bounds + slot after a[i-1]
    int64_t prev_last = -1;          // lower bound of the last b of the previous chunk
    int64_t run_start_carry = 0;     // index j where the run that reaches into this chunk started
    for (uint32_t base = 0; base < b.len; base += 32) {
        const uint32_t j = base + lane;
        const bool ok = j < b.len;
        const int64_t i = ok ? pb[j] : -2;
        int64_t prev = __shfl_up_sync(kFullMask, i, 1);
        if (lane == 0) prev = prev_last;
        const bool start = ok && (i != prev);
        int64_t st = start ? (int64_t)j : -1;   // max-scan of run starts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(kFullMask, st, d); if (lane >= d) st = max(st, t); }
        if (st < 0) st = run_start_carry;
        if (ok) {
            const int64_t acc = (int64_t)j - st;
            pb[j] = acc + ((i == 0) ? out_begin : ((int64_t)__ldcg(reinterpret_cast<const long long*>(pa + i - 1)) + 1));
        }
        const int last_lane = (int)min(31u, b.len - 1 - base);
        prev_last = __shfl_sync(kFullMask, i, last_lane);
        run_start_carry = __shfl_sync(kFullMask, st, last_lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-pack sort (packed_sort_qsort / packed_sort_thrust, pack_ops_cuda.cu:2556-2763): ascending, in place, optional
// permutation of the GLOBAL indices.  Direction-free bitonic network: every compare-exchange puts the smaller key at
// the lower index, so virtual +inf padding beyond the pack needs no storage.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void bitonic_network(T* keys, int64_t* ids, uint32_t n) {
    uint32_t Pw = 1;
    while (Pw < n) Pw <<= 1;
    for (uint32_t k = 2; k <= Pw; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const bool first = (j == (k >> 1));
            for (uint32_t i = threadIdx.x; i < Pw; i += blockDim.x) {
                const uint32_t l = first ? (i ^ (k - 1)) : (i ^ j);
                if (l > i && l < n) {
                    const T a = keys[i], b = keys[l];
                    if (b < a) {
                        keys[i] = b; keys[l] = a;
                        if (ids) { const int64_t t = ids[i]; ids[i] = ids[l]; ids[l] = t; }
                    }
                }
            }
            __syncthreads();
        }
    }
}

constexpr uint32_t kSortSmemElems = 2048;

template <typename T>
__global__ void __launch_bounds__(kNextThreads)
pack_sort_kernel(uint64_t P, T* __restrict__ vals, const int64_t* __restrict__ pack_infos, int64_t* __restrict__ idx) {
    __shared__ T skeys[kSortSmemElems];
    __shared__ int64_t sids[kSortSmemElems];
    for (uint64_t p = blockIdx.x; p < P; p += gridDim.x) {
        const PR r = pack_of(pack_infos, p);
        if (r.len < 2) continue;   // uniform per block
        T* v = vals + r.begin;
        int64_t* id = idx ? idx + r.begin : nullptr;
        if (r.len <= kSortSmemElems) {
            for (uint32_t i = threadIdx.x; i < r.len; i += blockDim.x) { skeys[i] = v[i]; if (id) sids[i] = id[i]; }
            __syncthreads();
            bitonic_network<T>(skeys, id ? sids : nullptr, r.len);
            for (uint32_t i = threadIdx.x; i < r.len; i += blockDim.x) { v[i] = skeys[i]; if (id) id[i] = sids[i]; }
            __syncthreads();
        } else {
            __syncthreads();
            bitonic_network<T>(v, id, r.len);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// packed_matmul (pack_ops_cuda.cu:2060-2085): out[i, o] = sum_k feats[i, k] * other[p, o, k]
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kNextThreads)
pack_matmul_kernel(uint64_t P, uint32_t C, uint32_t Co, const T* __restrict__ feats, const T* __restrict__ other,
                   const int64_t* __restrict__ pack_infos, T* __restrict__ out) {
    const uint64_t p = (uint64_t)blockIdx.x * kNextWarps + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = threadIdx.x & 31;
    const PR r = pack_of(pack_infos, p);
    const T* w = other + p * (uint64_t)Co * C;
    const uint64_t n = (uint64_t)r.len * Co;
    for (uint64_t t = lane; t < n; t += 32) {
        const uint64_t i = r.begin + t / Co;
        const uint32_t o = (uint32_t)(t % Co);
        T acc = (T)0;
        for (uint32_t k = 0; k < C; ++k) acc += feats[i * C + k] * w[(uint64_t)o * C + k];
        out[i * Co + o] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// interleave_sample_step_wrt_depth_in_packed_segments (pack_ops_cuda.cu:606-795): depth-proportional stepping restricted to
// the [entry, exit] segments of every ray.  Stepping is a serial recurrence per ray -> one thread per ray, two passes.
// Quirks kept: the walk to a segment entry advances by min_step at least once per segment; max_steps bounds the ray total.
// Fill pass with coalesced output (same idea as march_fill_staged_kernel in march.cu): every lane parks up to kSegStage samples of its
// ray in a shared-memory row, then the warp writes the rows out ray by ray (consecutive lanes -> consecutive addresses).  The segment
// walk of seg_sample_kernel is kept statement for statement, only made resumable (state: segment, t, step, inside-a-segment flag).
constexpr int kSegStage = 8;
constexpr int kSegStride = 9;
constexpr int kSegThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kSegThreads)
seg_sample_fill_staged_kernel(uint64_t P, T dt_gamma, T min_step, T max_step, const T* __restrict__ nears, const T* __restrict__ fars,
                              const T* __restrict__ entries, const T* __restrict__ exits, const int64_t* __restrict__ seg_pack_infos,
                              const int64_t* __restrict__ pack_infos, T* __restrict__ t_samples, T* __restrict__ deltas,
                              int64_t* __restrict__ nidx, int64_t* __restrict__ sidx) {
    __shared__ T s_t[kSegThreads / 32][32 * kSegStride];
    __shared__ T s_d[kSegThreads / 32][32 * kSegStride];
    __shared__ int64_t s_s[kSegThreads / 32][32 * kSegStride];
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool done = p >= P;
    const uint64_t pc = done ? 0 : p;
    const T near = nears[pc], far = fars[pc];
    uint64_t i = (uint64_t)seg_pack_infos[2 * pc];
    const uint64_t seg_end = i + (uint64_t)seg_pack_infos[2 * pc + 1];
    const uint64_t begin = (uint64_t)pack_infos[2 * pc];
    const uint32_t limit = done ? 0u : (uint32_t)pack_infos[2 * pc + 1];
    if (limit == 0) done = true;
    T t = near, exit_ = (T)0;
    uint32_t step = 0, flushed = 0;
    bool in_seg = false;
    while (true) {
        int pending = 0;
        while (!done && pending < kSegStage) {
            if (!in_seg) {
                if (i >= seg_end) { done = true; break; }
                const T entry = entries[i];
                exit_ = exits[i];
                if (entry >= far || exit_ <= near) { done = true; break; }
                do { t += min_step; } while (t < entry);
                in_seg = true;
            }
            if (t <= exit_ && t <= far && step < limit) {
                const T dt = clamp_next<T>(t * dt_gamma, min_step, max_step);
                s_t[wid][lane * kSegStride + pending] = t;
                s_d[wid][lane * kSegStride + pending] = dt;
                s_s[wid][lane * kSegStride + pending] = (int64_t)i;
                ++pending;
                t += dt;
                step++;
            } else {
                in_seg = false;
                ++i;
            }
        }
        __syncwarp();
        const uint32_t any = __ballot_sync(0xffffffffu, pending > 0);
        for (uint32_t m = any; m; m &= m - 1) {
            const int l = __ffs(m) - 1;
            const int n = __shfl_sync(0xffffffffu, pending, l);
            const uint64_t b = __shfl_sync(0xffffffffu, (unsigned long long)(begin + flushed), l);
            const int64_t ray = (int64_t)__shfl_sync(0xffffffffu, (unsigned long long)p, l);
            if (lane < n) {
                const int src = l * kSegStride + lane;
                t_samples[b + lane] = s_t[wid][src];
                deltas[b + lane] = s_d[wid][src];
                nidx[b + lane] = ray;
                sidx[b + lane] = s_s[wid][src];
            }
        }
        flushed += (uint32_t)pending;
        __syncwarp();
        if (__all_sync(0xffffffffu, done)) break;
    }
}

// ------------------------------------------------------------------------------------------------------------------

template <typename T, bool FILL>
__global__ void __launch_bounds__(kNextThreads)
seg_sample_kernel(uint64_t P, uint32_t max_steps, T dt_gamma, T min_step, T max_step, const T* __restrict__ nears, const T* __restrict__ fars,
                  const T* __restrict__ entries, const T* __restrict__ exits, const int64_t* __restrict__ seg_pack_infos,
                  const int64_t* __restrict__ pack_infos, int64_t* __restrict__ n_per_pack, T* __restrict__ t_samples, T* __restrict__ deltas,
                  int64_t* __restrict__ nidx, int64_t* __restrict__ sidx) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const T near = nears[p], far = fars[p];
    const uint64_t seg_begin = (uint64_t)seg_pack_infos[2 * p], seg_end = seg_begin + (uint64_t)seg_pack_infos[2 * p + 1];
    uint64_t begin = 0;
    uint32_t limit = max_steps;
    if (FILL) { begin = (uint64_t)pack_infos[2 * p]; limit = (uint32_t)pack_infos[2 * p + 1]; }
    T t = near;
    uint32_t step = 0;
    for (uint64_t i = seg_begin; i < seg_end; ++i) {
        const T entry = entries[i], exit_ = exits[i];
        if (entry >= far || exit_ <= near) break;
        do { t += min_step; } while (t < entry);
        while (t <= exit_ && t <= far && step < limit) {
            const T dt = clamp_next<T>(t * dt_gamma, min_step, max_step);
            if (FILL) {
                t_samples[begin + step] = t;
                deltas[begin + step] = dt;
                nidx[begin + step] = (int64_t)p;
                sidx[begin + step] = (int64_t)i;
            }
            t += dt;
            step++;
        }
    }
    if (!FILL) n_per_pack[p] = step;
}

// octree_mark_consecutive_segments (pack_ops_cuda.cu:2807-2841).  `offset_fix == 0` reproduces the reference, which indexes
// `point_indices` from 0 for every pack (only the first pack sees its own nuggets); `offset_fix != 0` indexes from the pack's begin.
__global__ void __launch_bounds__(kNextThreads)
mark_consecutive_kernel(uint64_t P, const int64_t* __restrict__ pack_infos, const int32_t* __restrict__ pidx, const int16_t* __restrict__ points,
                        int32_t offset_fix, uint8_t* __restrict__ mark_start, uint8_t* __restrict__ mark_end) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint64_t begin = (uint64_t)pack_infos[2 * p], len = (uint64_t)pack_infos[2 * p + 1];
    if (len == 0) return;
    const int32_t* pi = offset_fix ? pidx + begin : pidx;
    mark_start[begin] = 1;
    int32_t q = pi[0];
    int px = points[3 * q], py = points[3 * q + 1], pz = points[3 * q + 2];
    for (uint64_t j = 1; j < len; ++j) {
        q = pi[j];
        const int nx = points[3 * q], ny = points[3 * q + 1], nz = points[3 * q + 2];
        if (abs(nx - px) + abs(ny - py) + abs(nz - pz) > 1) { mark_end[begin + j - 1] = 1; mark_start[begin + j] = 1; }
        px = nx; py = ny; pz = nz;
    }
    mark_end[begin + len - 1] = 1;
}

static inline unsigned wgrid(uint64_t warps) { return (unsigned)div_up<uint64_t>(warps, kNextWarps); }

#define NR3D_NEXT_DISPATCH(dtype, NAME, ...)                                                         \
    switch (dtype) {                                                                                 \
    case NR3D_F32: { using T = float; __VA_ARGS__; } break;                                          \
    case NR3D_F64: { using T = double; __VA_ARGS__; } break;                                         \
    case NR3D_I32: { using T = int32_t; __VA_ARGS__; } break;                                        \
    case NR3D_I64: { using T = int64_t; __VA_ARGS__; } break;                                        \
    default: return fail(NAME ": unsupported dtype code %d (supported: f32, f64, i32, i64)", (int)dtype); }
#define NR3D_NEXT_DISPATCH_FLOAT(dtype, NAME, ...)                                                   \
    switch (dtype) {                                                                                 \
    case NR3D_F32: { using T = float; __VA_ARGS__; } break;                                          \
    case NR3D_F64: { using T = double; __VA_ARGS__; } break;                                         \
    default: return fail(NAME ": expected f32 / f64, got dtype code %d", (int)dtype); }

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_pack_searchsorted(int32_t dtype, uint64_t P, const void* bins, const int64_t* pack_infos, const void* vals,
                           uint32_t num_to_search, const int64_t* val_pack_infos, int64_t* pidx, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(bins && pack_infos && vals && pidx, "packed_searchsorted: null argument");
    NR3D_NEXT_DISPATCH(dtype, "packed_searchsorted",
        (searchsorted_kernel<T><<<wgrid(P), kNextThreads, 0, (cudaStream_t)stream>>>(P, (const T*)bins, pack_infos, (const T*)vals, num_to_search, val_pack_infos, pidx)));
    NR3D_LAUNCH_CHECK("packed_searchsorted");
    return 0;
}

int nr3d_pack_invert_cdf(int32_t dtype, uint64_t P, const void* bins, const void* cdfs, const int64_t* pack_infos, const void* u_vals,
                         uint32_t num_to_sample, void* samples, int64_t* bin_idx, void* stream) {
    if (P == 0 || num_to_sample == 0) return 0;
    NR3D_CHECK(bins && cdfs && pack_infos && u_vals && samples && bin_idx, "packed_invert_cdf: null argument");
    NR3D_NEXT_DISPATCH_FLOAT(dtype, "packed_invert_cdf",
        (invert_cdf_kernel<T><<<wgrid(P), kNextThreads, 0, (cudaStream_t)stream>>>(P, (const T*)bins, (const T*)cdfs, pack_infos, (const T*)u_vals, num_to_sample, (T*)samples, bin_idx)));
    NR3D_LAUNCH_CHECK("packed_invert_cdf");
    return 0;
}

int nr3d_pack_merge_sorted_aligned(int32_t dtype, uint64_t P, const void* vals_a, const int64_t* pack_infos_a, const void* vals_b,
                                   const int64_t* pack_infos_b, const int64_t* pack_infos_merged, int64_t* pidx_a, int64_t* pidx_b,
                                   void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(vals_a && pack_infos_a && vals_b && pack_infos_b && pack_infos_merged && pidx_a && pidx_b, "try_merge_two_packs_sorted_aligned: null argument");
    NR3D_NEXT_DISPATCH(dtype, "try_merge_two_packs_sorted_aligned",
        (merge_sorted_kernel<T><<<wgrid(P), kNextThreads, 0, (cudaStream_t)stream>>>(P, (const T*)vals_a, pack_infos_a, (const T*)vals_b, pack_infos_b, pack_infos_merged, pidx_a, pidx_b)));
    NR3D_LAUNCH_CHECK("try_merge_two_packs_sorted_aligned");
    return 0;
}

int nr3d_pack_sort(int32_t dtype, uint64_t P, void* vals, const int64_t* pack_infos, int64_t* idx, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(vals && pack_infos, "packed_sort: null argument");
    const unsigned grid = (unsigned)(P < 148ull * 32 ? P : 148ull * 32);
    NR3D_NEXT_DISPATCH(dtype, "packed_sort",
        (pack_sort_kernel<T><<<grid, kNextThreads, 0, (cudaStream_t)stream>>>(P, (T*)vals, pack_infos, idx)));
    NR3D_LAUNCH_CHECK("packed_sort");
    return 0;
}

int nr3d_pack_matmul(int32_t dtype, uint64_t P, uint32_t C, uint32_t C_out, const void* feats, const void* other, const int64_t* pack_infos,
                     void* out, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(feats && other && pack_infos && out && C > 0 && C_out > 0, "packed_matmul: null argument");
    NR3D_NEXT_DISPATCH_FLOAT(dtype, "packed_matmul",
        (pack_matmul_kernel<T><<<wgrid(P), kNextThreads, 0, (cudaStream_t)stream>>>(P, C, C_out, (const T*)feats, (const T*)other, pack_infos, (T*)out)));
    NR3D_LAUNCH_CHECK("packed_matmul");
    return 0;
}

int nr3d_pack_seg_sample_count(int32_t dtype, uint64_t P, const void* nears, const void* fars, const void* entries, const void* exits,
                               const int64_t* seg_pack_infos, uint32_t max_steps, double dt_gamma, double min_step, double max_step,
                               int64_t* n_per_pack, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(nears && fars && seg_pack_infos && n_per_pack, "interleave_sample_step_wrt_depth_in_packed_segments: null argument");
    const unsigned grid = (unsigned)div_up<uint64_t>(P, kNextThreads);
    NR3D_NEXT_DISPATCH_FLOAT(dtype, "interleave_sample_step_wrt_depth_in_packed_segments",
        (seg_sample_kernel<T, false><<<grid, kNextThreads, 0, (cudaStream_t)stream>>>(P, max_steps, (T)dt_gamma, (T)min_step, (T)max_step, (const T*)nears,
            (const T*)fars, (const T*)entries, (const T*)exits, seg_pack_infos, nullptr, n_per_pack, nullptr, nullptr, nullptr, nullptr)));
    NR3D_LAUNCH_CHECK("seg_sample_count");
    return 0;
}

int nr3d_pack_seg_sample_fill(int32_t dtype, uint64_t P, const void* nears, const void* fars, const void* entries, const void* exits,
                              const int64_t* seg_pack_infos, const int64_t* pack_infos, double dt_gamma, double min_step, double max_step,
                              void* t_samples, void* deltas, int64_t* nidx, int64_t* sidx, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(nears && fars && seg_pack_infos && pack_infos, "interleave_sample_step_wrt_depth_in_packed_segments: null argument");
    const unsigned grid = (unsigned)div_up<uint64_t>(P, kSegThreads);
    NR3D_NEXT_DISPATCH_FLOAT(dtype, "interleave_sample_step_wrt_depth_in_packed_segments",
        (seg_sample_fill_staged_kernel<T><<<grid, kSegThreads, 0, (cudaStream_t)stream>>>(P, (T)dt_gamma, (T)min_step, (T)max_step, (const T*)nears,
            (const T*)fars, (const T*)entries, (const T*)exits, seg_pack_infos, pack_infos, (T*)t_samples, (T*)deltas, nidx, sidx)));
    NR3D_LAUNCH_CHECK("seg_sample_fill");
    return 0;
}

int nr3d_pack_mark_consecutive_segments(uint64_t P, const int64_t* pack_infos, const int32_t* pidx, const int16_t* point_hierarchies,
                                        int32_t offset_fix, uint8_t* mark_start, uint8_t* mark_end, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(pack_infos && pidx && point_hierarchies && mark_start && mark_end, "octree_mark_consecutive_segments: null argument");
    mark_consecutive_kernel<<<(unsigned)div_up<uint64_t>(P, kNextThreads), kNextThreads, 0, (cudaStream_t)stream>>>(P, pack_infos, pidx,
        point_hierarchies, offset_fix, mark_start, mark_end);
    NR3D_LAUNCH_CHECK("octree_mark_consecutive_segments");
    return 0;
}

}  // extern "C"
