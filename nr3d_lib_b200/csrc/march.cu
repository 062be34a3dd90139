// march.cu -- occupancy-grid ray marching (single + batched) for sm_100a.
//
// Behavioural contract: csrc/occ_grid/src/ray_marching.cu:17-134 and batched_marching.cu:18-148 with
// include/occ_grid/helpers_march.h:11-77 and helpers_contraction.h:10-125.  The float expressions below keep the
// reference's operation order (and rely on the same nvcc a*b+c contraction) so that sample counts and
// pack offsets are bit-exact.  Differences in structure, not in results:
//   * one templated kernel serves both passes and both the single / batched variants;
//   * the per-ray counts are turned into packed_info by an on-device scan (pack_ops.cu) -- no cumsum/stack
//     tensor ops and a single host read of the total;
//   * rays with batch_inds < 0 get count 0 (the reference leaves `num_steps` uninitialised for them).
#include "common.cuh"

namespace nr3d {

template <typename TI, typename TO>
int scan_counts(uint64_t n, const TI* counts, TO* pack_infos, int64_t* total, void* ws, uint64_t* ws_bytes, cudaStream_t stream);

struct F3 { float x, y, z; };

__device__ __forceinline__ float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
__device__ __forceinline__ float calc_dt(float t, float dt_gamma, float dt_min, float dt_max) { return clampf(t * dt_gamma, dt_min, dt_max); }

__device__ __forceinline__ F3 roi_to_unit(F3 p, F3 lo, F3 hi) {
    return F3{(p.x - lo.x) / (hi.x - lo.x), (p.y - lo.y) / (hi.y - lo.y), (p.z - lo.z) / (hi.z - lo.z)};
}

__device__ __forceinline__ F3 apply_contraction(F3 p, F3 lo, F3 hi, int type) {
    F3 u = roi_to_unit(p, lo, hi);
    if (type == NR3D_CONTRACTION_TANH) {  // helpers_contraction.h:24-35
        u = F3{u.x - 0.5f, u.y - 0.5f, u.z - 0.5f};
        return F3{tanhf(u.x) * 0.5f + 0.5f, tanhf(u.y) * 0.5f + 0.5f, tanhf(u.z) * 0.5f + 0.5f};
    }
    if (type == NR3D_CONTRACTION_SPHERE) {  // helpers_contraction.h:58-76
        u = F3{u.x * 2.0f - 1.0f, u.y * 2.0f - 1.0f, u.z * 2.0f - 1.0f};
        const float norm_sq = u.x * u.x + u.y * u.y + u.z * u.z;
        const float norm = sqrtf(norm_sq);
        if (norm > 1.0f) {
            const float s = 2.0f - 1.0f / norm;
            u = F3{s * (u.x / norm), s * (u.y / norm), s * (u.z / norm)};
        }
        return F3{u.x * 0.25f + 0.5f, u.y * 0.25f + 0.5f, u.z * 0.25f + 0.5f};
    }
    return u;
}

__device__ __forceinline__ bool grid_occupied_at(F3 p, F3 lo, F3 hi, int type, int3 res, const uint8_t* __restrict__ grid, int* idx_out) {
    if (type == NR3D_CONTRACTION_AABB &&
        (p.x < lo.x || p.x > hi.x || p.y < lo.y || p.y > hi.y || p.z < lo.z || p.z > hi.z)) return false;
    const F3 u = apply_contraction(p, lo, hi, type);
    int ix = (int)(u.x * (float)res.x), iy = (int)(u.y * (float)res.y), iz = (int)(u.z * (float)res.z);
    ix = max(0, min(ix, res.x - 1));
    iy = max(0, min(iy, res.y - 1));
    iz = max(0, min(iz, res.z - 1));
    const int idx = ix * (res.y * res.z) + iy * res.z + iz;
    *idx_out = idx;
    return grid[idx] != 0;
}

__device__ __forceinline__ float distance_to_next_voxel(F3 p, F3 dir, F3 inv_dir, F3 lo, F3 hi, int3 res) {
    const F3 r = F3{(float)res.x, (float)res.y, (float)res.z};
    const F3 u = roi_to_unit(p, lo, hi);
    const F3 q = F3{u.x * r.x, u.y * r.y, u.z * r.z};
    const float tx = ((floorf(q.x + 0.5f + 0.5f * copysignf(1.0f, dir.x)) - q.x) * inv_dir.x) / r.x * (hi.x - lo.x);
    const float ty = ((floorf(q.y + 0.5f + 0.5f * copysignf(1.0f, dir.y)) - q.y) * inv_dir.y) / r.y * (hi.y - lo.y);
    const float tz = ((floorf(q.z + 0.5f + 0.5f * copysignf(1.0f, dir.z)) - q.z) * inv_dir.z) / r.z * (hi.z - lo.z);
    return fmaxf(fminf(fminf(tx, ty), tz), 0.0f);
}

__device__ __forceinline__ float advance_to_next_voxel(float t, float dt_min, F3 p, F3 dir, F3 inv_dir, F3 lo, F3 hi, int3 res) {
    const float t_target = t + distance_to_next_voxel(p, dir, inv_dir, lo, hi, res);
    float t_ = t;
    do { t_ += dt_min; } while (t_ < t_target);
    return t_;
}

struct MarchArgs {
    uint64_t n_rays;
    const float *rays_o, *rays_d, *t_min, *t_max;
    const int32_t* batch_inds;
    uint32_t batch_data_size;
    uint32_t n_batches;   // rays whose batch index is outside [0, n_batches) are skipped (count 0) instead of reading grid / roi out of bounds
    const float* roi;
    const uint8_t* grid;
    int3 res;
    int type;
    float step_size, max_step_size, dt_gamma;
    uint32_t max_steps;
};

// MODE 0: count pass (num_steps).  MODE 1: fill pass with thread-per-ray stores (A/B only, the staged fill kernel below is used).
// MODE 2: RECORD pass of the single-pass marcher: counts like MODE 0 and parks every sample as one 16-byte record (t_start, t_end, voxel id)
// in rec[j * n_rays + ray] -- sample-major, so the 32 rays of a warp fill one 512-byte row per sample index and the rows are complete by the
// time L2 writes them back.  march_compact_kernel then moves the records to their packed positions; the ray is marched ONCE instead of twice.
template <int MODE>
__global__ void __launch_bounds__(256)
march_kernel(const MarchArgs a, const int32_t* __restrict__ packed_info, int32_t* __restrict__ num_steps, float* __restrict__ t_starts,
             float* __restrict__ t_ends, int32_t* __restrict__ ridx_out, int32_t* __restrict__ bidx_out, int32_t* __restrict__ gidx_out,
             float4* __restrict__ rec) {
    constexpr bool FILL = MODE == 1;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_rays) return;
    uint32_t batch_ind = 0;
    if (a.batch_inds) {
        const int32_t b = a.batch_inds[i];
        if (b < 0 || (uint32_t)b >= a.n_batches) {
            if (!FILL) num_steps[i] = 0;
            return;
        }
        batch_ind = (uint32_t)b;
    } else if (a.batch_data_size) {
        batch_ind = (uint32_t)(i / a.batch_data_size);
        if (batch_ind >= a.n_batches) {
            if (!FILL) num_steps[i] = 0;
            return;
        }
    }
    const uint32_t cells = (uint32_t)(a.res.x * a.res.y * a.res.z);
    const uint32_t grid_offset = batch_ind * cells;
    const uint8_t* __restrict__ grid = a.grid + grid_offset;
    const float* roi = a.roi + (uint64_t)batch_ind * 6;

    uint32_t max_steps = a.max_steps;
    uint64_t base = 0;
    if (FILL) {
        base = (uint32_t)packed_info[i * 2 + 0];
        max_steps = (uint32_t)packed_info[i * 2 + 1];
        if (max_steps == 0) return;
    }
    const F3 origin = F3{a.rays_o[i * 3 + 0], a.rays_o[i * 3 + 1], a.rays_o[i * 3 + 2]};
    const F3 dir = F3{a.rays_d[i * 3 + 0], a.rays_d[i * 3 + 1], a.rays_d[i * 3 + 2]};
    const F3 inv_dir = F3{1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
    const float near = a.t_min[i], far = a.t_max[i];
    const F3 lo = F3{roi[0], roi[1], roi[2]};
    const F3 hi = F3{roi[3], roi[4], roi[5]};
    const float dt_min = a.step_size, dt_max = a.max_step_size;

    uint32_t j = 0;
    float t0 = near;
    float dt = calc_dt(t0, a.dt_gamma, dt_min, dt_max);
    float t1 = t0 + dt;
    float t_mid = (t0 + t1) * 0.5f;
    while ((t_mid < far) && (j < max_steps)) {
        const F3 p = F3{origin.x + t_mid * dir.x, origin.y + t_mid * dir.y, origin.z + t_mid * dir.z};
        int grid_idx = -1;
        if (grid_occupied_at(p, lo, hi, a.type, a.res, grid, &grid_idx)) {
            if (FILL) {
                t_starts[base + j] = t0;
                t_ends[base + j] = t1;
                ridx_out[base + j] = (int32_t)i;
                if (bidx_out) bidx_out[base + j] = (int32_t)batch_ind;
                if (gidx_out) gidx_out[base + j] = grid_idx + (int32_t)grid_offset;
            }
            if (MODE == 2) __stcg(rec + (uint64_t)j * a.n_rays + i, make_float4(t0, t1, __int_as_float(grid_idx + (int32_t)grid_offset), 0.f));
            ++j;
            t0 = t1;
            t1 = t0 + calc_dt(t0, a.dt_gamma, dt_min, dt_max);
            t_mid = (t0 + t1) * 0.5f;
        } else if (a.type == NR3D_CONTRACTION_AABB) {
            t_mid = advance_to_next_voxel(t_mid, dt_min, p, dir, inv_dir, lo, hi, a.res);
            dt = calc_dt(t_mid, a.dt_gamma, dt_min, dt_max);
            t0 = t_mid - dt * 0.5f;
            t1 = t_mid + dt * 0.5f;
        } else {
            t0 = t1;
            t1 = t0 + calc_dt(t0, a.dt_gamma, dt_min, dt_max);
            t_mid = (t0 + t1) * 0.5f;
        }
    }
    if (!FILL) num_steps[i] = (int32_t)j;
}

// Second half of the single-pass marcher: records rec[j * n_rays + ray] (written by march_kernel<2>) -> packed outputs.  One warp owns 32
// consecutive rays, whose samples form ONE contiguous span of the outputs.  Per round of 32 sample indices: 32 coalesced 512-byte row reads
// into a padded shared-memory tile (conflict-free both ways), then ray after ray the warp writes 32 consecutive samples of that ray
// (consecutive lanes -> consecutive addresses).  Rows beyond a ray's own count hold stale bytes that are read and never used.
constexpr int kCompactThreads = 64;
__global__ void __launch_bounds__(kCompactThreads)
march_compact_kernel(uint64_t n_rays, const int32_t* __restrict__ batch_inds, uint32_t batch_data_size, const float4* __restrict__ rec,
                     const int32_t* __restrict__ packed_info, float* __restrict__ t_starts, float* __restrict__ t_ends,
                     int32_t* __restrict__ ridx_out, int32_t* __restrict__ bidx_out, int32_t* __restrict__ gidx_out) {
    __shared__ float s_t0[kCompactThreads / 32][32 * 33];
    __shared__ float s_t1[kCompactThreads / 32][32 * 33];
    __shared__ int32_t s_g[kCompactThreads / 32][32 * 33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t i0 = ((uint64_t)blockIdx.x * (kCompactThreads / 32) + wid) * 32;
    if (i0 >= n_rays) return;   // whole warp
    const uint64_t i = i0 + lane;
    const bool have = i < n_rays;
    const int32_t base = have ? packed_info[i * 2] : 0, cnt = have ? packed_info[i * 2 + 1] : 0;
    int32_t bat = 0;
    if (have && bidx_out) bat = batch_inds ? batch_inds[i] : (batch_data_size ? (int32_t)(i / batch_data_size) : 0);
    const int32_t maxc = __reduce_max_sync(0xffffffffu, cnt);
    float* t0s = s_t0[wid]; float* t1s = s_t1[wid]; int32_t* gs = s_g[wid];
    for (int32_t j0 = 0; j0 < maxc; j0 += 32) {
        const int rows = min(32, maxc - j0);   // warp uniform
        if (have) {
#pragma unroll 8
            for (int jj = 0; jj < rows; ++jj) {
                const float4 v = __ldcs(rec + (uint64_t)(j0 + jj) * n_rays + i);
                t0s[jj * 33 + lane] = v.x; t1s[jj * 33 + lane] = v.y; gs[jj * 33 + lane] = __float_as_int(v.z);
            }
        }
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const int32_t c = __shfl_sync(0xffffffffu, cnt, r) - j0;     // samples of ray r left from this round on
            if (c <= 0) continue;                                        // warp uniform
            const int64_t b = (int64_t)__shfl_sync(0xffffffffu, base, r) + j0;
            const int32_t br = __shfl_sync(0xffffffffu, bat, r);
            if (lane < c) {
                t_starts[b + lane] = t0s[lane * 33 + r];
                t_ends[b + lane] = t1s[lane * 33 + r];
                ridx_out[b + lane] = (int32_t)(i0 + r);
                if (bidx_out) bidx_out[b + lane] = br;
                if (gidx_out) gidx_out[b + lane] = gs[lane * 33 + r];
            }
        }
        __syncwarp();
    }
}

// Fill pass with coalesced output.  One thread per ray writes t_starts[base + j] etc. with a different base per lane: every store
// instruction of a warp touches 32 sectors for 128 useful bytes.  Here each lane parks up to kStage samples in a shared-memory row,
// then the warp writes the rows out one ray at a time (consecutive lanes -> consecutive addresses, full sectors).  The march itself
// is the loop of march_kernel, statement for statement (bit-exact outputs), only chunked.
constexpr int kStage = 16;            // samples parked per ray and round
constexpr int kStageStride = 17;      // row stride (floats): conflict-free when the lanes of a warp write the same column
constexpr int kFillThreads = 128;

__global__ void __launch_bounds__(kFillThreads)
march_fill_staged_kernel(const MarchArgs a, const int32_t* __restrict__ packed_info, float* __restrict__ t_starts, float* __restrict__ t_ends,
                         int32_t* __restrict__ ridx_out, int32_t* __restrict__ bidx_out, int32_t* __restrict__ gidx_out) {
    __shared__ float s_t0[kFillThreads / 32][32 * kStageStride];
    __shared__ float s_t1[kFillThreads / 32][32 * kStageStride];
    __shared__ int32_t s_g[kFillThreads / 32][32 * kStageStride];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* my_t0 = s_t0[wid] + lane * kStageStride;
    float* my_t1 = s_t1[wid] + lane * kStageStride;
    int32_t* my_g = s_g[wid] + lane * kStageStride;
    bool done = i >= a.n_rays;
    uint32_t batch_ind = 0;
    if (!done) {
        if (a.batch_inds) {
            const int32_t b = a.batch_inds[i];
            if (b < 0 || (uint32_t)b >= a.n_batches) done = true; else batch_ind = (uint32_t)b;
        } else if (a.batch_data_size) {
            batch_ind = (uint32_t)(i / a.batch_data_size);
            if (batch_ind >= a.n_batches) { done = true; batch_ind = 0; }
        }
    }
    const uint32_t cells = (uint32_t)(a.res.x * a.res.y * a.res.z);
    const uint32_t grid_offset = batch_ind * cells;
    const uint8_t* __restrict__ grid = a.grid + grid_offset;
    const float* roi = a.roi + (uint64_t)batch_ind * 6;
    uint32_t max_steps = 0;
    uint64_t base = 0;
    if (!done) {
        base = (uint32_t)packed_info[i * 2 + 0];
        max_steps = (uint32_t)packed_info[i * 2 + 1];
        if (max_steps == 0) done = true;
    }
    const uint64_t ic = done ? 0 : i;   // finished / padding lanes read ray 0's data and never use it
    const F3 origin = F3{a.rays_o[ic * 3 + 0], a.rays_o[ic * 3 + 1], a.rays_o[ic * 3 + 2]};
    const F3 dir = F3{a.rays_d[ic * 3 + 0], a.rays_d[ic * 3 + 1], a.rays_d[ic * 3 + 2]};
    const F3 inv_dir = F3{1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
    const float near = a.t_min[ic], far = a.t_max[ic];
    const F3 lo = F3{roi[0], roi[1], roi[2]};
    const F3 hi = F3{roi[3], roi[4], roi[5]};
    const float dt_min = a.step_size, dt_max = a.max_step_size;

    uint32_t j = 0, flushed = 0;
    float t0 = near;
    float dt = calc_dt(t0, a.dt_gamma, dt_min, dt_max);
    float t1 = t0 + dt;
    float t_mid = (t0 + t1) * 0.5f;
    while (true) {
        int pending = 0;
        while (!done && pending < kStage) {
            if (!((t_mid < far) && (j < max_steps))) { done = true; break; }
            const F3 p = F3{origin.x + t_mid * dir.x, origin.y + t_mid * dir.y, origin.z + t_mid * dir.z};
            int grid_idx = -1;
            if (grid_occupied_at(p, lo, hi, a.type, a.res, grid, &grid_idx)) {
                my_t0[pending] = t0;
                my_t1[pending] = t1;
                my_g[pending] = grid_idx + (int32_t)grid_offset;
                ++pending;
                ++j;
                t0 = t1;
                t1 = t0 + calc_dt(t0, a.dt_gamma, dt_min, dt_max);
                t_mid = (t0 + t1) * 0.5f;
            } else if (a.type == NR3D_CONTRACTION_AABB) {
                t_mid = advance_to_next_voxel(t_mid, dt_min, p, dir, inv_dir, lo, hi, a.res);
                dt = calc_dt(t_mid, a.dt_gamma, dt_min, dt_max);
                t0 = t_mid - dt * 0.5f;
                t1 = t_mid + dt * 0.5f;
            } else {
                t0 = t1;
                t1 = t0 + calc_dt(t0, a.dt_gamma, dt_min, dt_max);
                t_mid = (t0 + t1) * 0.5f;
            }
        }
        __syncwarp();
        // write the parked rows: ray by ray, lanes = consecutive samples of that ray
        const uint32_t any = __ballot_sync(0xffffffffu, pending > 0);
        for (uint32_t m = any; m; m &= m - 1) {
            const int l = __ffs(m) - 1;
            const int n = __shfl_sync(0xffffffffu, pending, l);
            const uint64_t b = __shfl_sync(0xffffffffu, (unsigned long long)(base + flushed), l);
            const int32_t ray = (int32_t)__shfl_sync(0xffffffffu, (unsigned long long)i, l);
            const int32_t bat = (int32_t)__shfl_sync(0xffffffffu, batch_ind, l);
            if (lane < n) {
                const int src = l * kStageStride + lane;
                t_starts[b + lane] = s_t0[wid][src];
                t_ends[b + lane] = s_t1[wid][src];
                ridx_out[b + lane] = ray;
                if (bidx_out) bidx_out[b + lane] = bat;
                if (gidx_out) gidx_out[b + lane] = s_g[wid][src];
            }
        }
        flushed += (uint32_t)pending;
        __syncwarp();
        if (__all_sync(0xffffffffu, done)) break;
    }
}

static int fill_args(MarchArgs& a, uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                     const int32_t* batch_inds, uint32_t batch_data_size, int32_t n_batches, const float* roi, const uint8_t* grid,
                     int32_t rx, int32_t ry, int32_t rz, int32_t contraction, float step_size, float max_step_size, float dt_gamma,
                     uint32_t max_steps) {
    NR3D_CHECK(rays_o && rays_d && t_min && t_max && roi && grid, "ray_marching: null argument");
    NR3D_CHECK(rx > 0 && ry > 0 && rz > 0 && n_batches > 0, "ray_marching: invalid grid shape");
    NR3D_CHECK(contraction >= 0 && contraction <= 2, "ray_marching: invalid contraction type %d", contraction);
    NR3D_CHECK(batch_data_size == 0 || n_rays % batch_data_size == 0,
               "batched_ray_marching: Expect nonzero `batch_data_size`=%u to be a divisor of `n_rays`=%llu", batch_data_size,
               (unsigned long long)n_rays);
    NR3D_CHECK(n_rays < (1ull << 31), "ray_marching: n_rays must be < 2^31");
    a.n_rays = n_rays; a.rays_o = rays_o; a.rays_d = rays_d; a.t_min = t_min; a.t_max = t_max;
    a.batch_inds = batch_inds; a.batch_data_size = batch_data_size; a.n_batches = (uint32_t)n_batches; a.roi = roi; a.grid = grid;
    a.res = make_int3(rx, ry, rz); a.type = contraction;
    a.step_size = step_size; a.max_step_size = max_step_size; a.dt_gamma = dt_gamma; a.max_steps = max_steps;
    return 0;
}


// ------------------------------------------------------------------------------------------------------------
// forest (multi-block) marcher: SURVEY.md 8f row n4, csrc/occ_grid/src/forest_marching.cu:27-150.
// Each ray carries a pack of block segments (block index, entry depth, exit depth) produced by the octree ray trace; the
// march walks them front to back with ONE running step counter / depth, using block `b`'s occupancy grid inside
// [world_origin + k_b * world_block_size, + world_block_size].  Unlike the single-grid marcher there is no AABB test
// (the segment bounds play that role) and the voxel index is clamped.  Same float op order as the reference.
// ------------------------------------------------------------------------------------------------------------
struct ForestMarchArgs {
    uint64_t n_rays;
    const float *rays_o, *rays_d, *t_min, *t_max;
    const int32_t* seg_block_inds;
    const float *seg_entries, *seg_exits;
    const int32_t* seg_pack_infos;
    const int16_t* block_ks;  // [n_trees, 3]
    F3 world_origin, world_block_size;
    const uint8_t* grid;      // [n_trees, rx, ry, rz]
    int3 res;
    float step_size, max_step_size, dt_gamma;
    uint32_t max_steps;
};

template <bool FILL>
__global__ void __launch_bounds__(256)
forest_march_kernel(const ForestMarchArgs a, const int32_t* __restrict__ packed_info, int32_t* __restrict__ num_steps,
                    float* __restrict__ t_starts, float* __restrict__ t_ends, int32_t* __restrict__ ridx_out,
                    int32_t* __restrict__ blidx_out, int32_t* __restrict__ gidx_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_rays) return;
    const uint32_t seg_begin = (uint32_t)a.seg_pack_infos[i * 2], seg_length = (uint32_t)a.seg_pack_infos[i * 2 + 1];
    const uint32_t cells = (uint32_t)(a.res.x * a.res.y * a.res.z);
    uint32_t max_steps = a.max_steps;
    uint64_t base = 0;
    if (FILL) {
        base = (uint32_t)packed_info[i * 2 + 0];
        max_steps = (uint32_t)packed_info[i * 2 + 1];
        if (max_steps == 0) return;
    }
    const F3 origin = F3{a.rays_o[i * 3 + 0], a.rays_o[i * 3 + 1], a.rays_o[i * 3 + 2]};
    const F3 dir = F3{a.rays_d[i * 3 + 0], a.rays_d[i * 3 + 1], a.rays_d[i * 3 + 2]};
    const F3 inv_dir = F3{1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
    const float near = a.t_min[i], far = a.t_max[i];
    const float dt_min = a.step_size, dt_max = a.max_step_size;

    uint32_t j = 0;
    float t0 = near;
    float dt = calc_dt(t0, a.dt_gamma, dt_min, dt_max);
    float t1 = t0 + dt;
    float t_mid = (t0 + t1) * 0.5f;
    for (uint32_t s = 0; s < seg_length; ++s) {
        const float cur_entry = a.seg_entries[seg_begin + s], cur_exit = a.seg_exits[seg_begin + s];
        const uint32_t block_ind = (uint32_t)a.seg_block_inds[seg_begin + s];
        const int16_t* k = a.block_ks + (uint64_t)block_ind * 3;
        const F3 lo = F3{a.world_origin.x + (float)k[0] * a.world_block_size.x, a.world_origin.y + (float)k[1] * a.world_block_size.y,
                         a.world_origin.z + (float)k[2] * a.world_block_size.z};
        const F3 hi = F3{lo.x + a.world_block_size.x, lo.y + a.world_block_size.y, lo.z + a.world_block_size.z};
        const uint32_t grid_offset = block_ind * cells;
        const uint8_t* __restrict__ grid = a.grid + grid_offset;
        if (cur_entry >= far || cur_exit <= near) break;
        do { t_mid += a.step_size; } while (t_mid < cur_entry);   // march to the entry of this block segment
        dt = calc_dt(t_mid, a.dt_gamma, dt_min, dt_max);
        t0 = t_mid - dt * 0.5f;
        t1 = t_mid + dt * 0.5f;
        while ((t_mid <= cur_exit) && (t_mid <= far) && (j < max_steps)) {
            const F3 p = F3{origin.x + t_mid * dir.x, origin.y + t_mid * dir.y, origin.z + t_mid * dir.z};
            const F3 u = roi_to_unit(p, lo, hi);
            int ix = (int)(u.x * (float)a.res.x), iy = (int)(u.y * (float)a.res.y), iz = (int)(u.z * (float)a.res.z);
            ix = max(0, min(ix, a.res.x - 1));
            iy = max(0, min(iy, a.res.y - 1));
            iz = max(0, min(iz, a.res.z - 1));
            const int grid_idx = ix * (a.res.y * a.res.z) + iy * a.res.z + iz;
            if (grid[grid_idx] != 0) {
                if (FILL) {
                    t_starts[base + j] = t0;
                    t_ends[base + j] = t1;
                    ridx_out[base + j] = (int32_t)i;
                    blidx_out[base + j] = (int32_t)block_ind;
                    if (gidx_out) gidx_out[base + j] = grid_idx + (int32_t)grid_offset;
                }
                ++j;
                t0 = t1;
                t1 = t0 + calc_dt(t0, a.dt_gamma, dt_min, dt_max);
                t_mid = (t0 + t1) * 0.5f;
            } else {
                t_mid = advance_to_next_voxel(t_mid, dt_min, p, dir, inv_dir, lo, hi, a.res);
                dt = calc_dt(t_mid, a.dt_gamma, dt_min, dt_max);
                t0 = t_mid - dt * 0.5f;
                t1 = t_mid + dt * 0.5f;
            }
        }
    }
    if (!FILL) num_steps[i] = (int32_t)j;
}

static int fill_forest_args(ForestMarchArgs& a, uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                            const int32_t* seg_block_inds, const float* seg_entries, const float* seg_exits, const int32_t* seg_pack_infos,
                            const int16_t* block_ks, const float* world_origin, const float* world_block_size, const uint8_t* grid,
                            int32_t rx, int32_t ry, int32_t rz, float step_size, float max_step_size, float dt_gamma, uint32_t max_steps) {
    NR3D_CHECK(rays_o && rays_d && t_min && t_max && seg_pack_infos && block_ks && world_origin && world_block_size && grid,
               "forest_ray_marching: null argument");
    NR3D_CHECK(rx > 0 && ry > 0 && rz > 0, "forest_ray_marching: invalid grid shape");
    NR3D_CHECK(n_rays < (1ull << 31), "forest_ray_marching: n_rays must be < 2^31");
    a.n_rays = n_rays; a.rays_o = rays_o; a.rays_d = rays_d; a.t_min = t_min; a.t_max = t_max;
    a.seg_block_inds = seg_block_inds; a.seg_entries = seg_entries; a.seg_exits = seg_exits; a.seg_pack_infos = seg_pack_infos;
    a.block_ks = block_ks;
    a.world_origin = F3{world_origin[0], world_origin[1], world_origin[2]};
    a.world_block_size = F3{world_block_size[0], world_block_size[1], world_block_size[2]};
    a.grid = grid; a.res = make_int3(rx, ry, rz);
    a.step_size = step_size; a.max_step_size = max_step_size; a.dt_gamma = dt_gamma; a.max_steps = max_steps;
    return 0;
}

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_march_count(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                     const int32_t* batch_inds, uint32_t batch_data_size, int32_t n_batches, const float* roi, const uint8_t* grid,
                     int32_t rx, int32_t ry, int32_t rz, int32_t contraction, float step_size, float max_step_size, float dt_gamma,
                     uint32_t max_steps, int32_t* num_steps, void* stream) {
    if (n_rays == 0) return 0;
    MarchArgs a;
    if (int rc = fill_args(a, n_rays, rays_o, rays_d, t_min, t_max, batch_inds, batch_data_size, n_batches, roi, grid, rx, ry, rz,
                           contraction, step_size, max_step_size, dt_gamma, max_steps)) return rc;
    NR3D_CHECK(num_steps != nullptr, "ray_marching: null num_steps");
    march_kernel<0><<<(unsigned)div_up<uint64_t>(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(a, nullptr, num_steps, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    NR3D_LAUNCH_CHECK("ray_marching(count)");
    return 0;
}

int nr3d_march_record(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                      const int32_t* batch_inds, uint32_t batch_data_size, int32_t n_batches, const float* roi, const uint8_t* grid,
                      int32_t rx, int32_t ry, int32_t rz, int32_t contraction, float step_size, float max_step_size, float dt_gamma,
                      uint32_t max_steps, int32_t* num_steps, void* records, uint64_t records_bytes, void* stream) {
    if (n_rays == 0) return 0;
    MarchArgs a;
    if (int rc = fill_args(a, n_rays, rays_o, rays_d, t_min, t_max, batch_inds, batch_data_size, n_batches, roi, grid, rx, ry, rz,
                           contraction, step_size, max_step_size, dt_gamma, max_steps)) return rc;
    NR3D_CHECK(num_steps != nullptr && records != nullptr, "ray_marching: null num_steps / records");
    NR3D_CHECK((reinterpret_cast<uintptr_t>(records) & 15u) == 0, "ray_marching: records must be 16-byte aligned");
    NR3D_CHECK(records_bytes / 16 / n_rays >= (uint64_t)max_steps, "ray_marching: records buffer too small (%llu bytes for %llu rays x %u steps)",
               (unsigned long long)records_bytes, (unsigned long long)n_rays, max_steps);
    march_kernel<2><<<(unsigned)div_up<uint64_t>(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(a, nullptr, num_steps, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                                             reinterpret_cast<float4*>(records));
    NR3D_LAUNCH_CHECK("ray_marching(record)");
    return 0;
}

int nr3d_march_compact(uint64_t n_rays, const int32_t* batch_inds, uint32_t batch_data_size, const void* records, const int32_t* packed_info,
                       float* t_starts, float* t_ends, int32_t* ridx, int32_t* bidx, int32_t* gidx, void* stream) {
    if (n_rays == 0) return 0;
    NR3D_CHECK(records && packed_info && t_starts && t_ends && ridx, "ray_marching(compact): null argument");
    NR3D_CHECK(n_rays < (1ull << 31), "ray_marching: n_rays must be < 2^31");
    constexpr int rays_per_cta = kCompactThreads;   // 32 rays per warp
    march_compact_kernel<<<(unsigned)div_up<uint64_t>(n_rays, rays_per_cta), kCompactThreads, 0, (cudaStream_t)stream>>>(
        n_rays, batch_inds, batch_data_size, reinterpret_cast<const float4*>(records), packed_info, t_starts, t_ends, ridx, bidx, gidx);
    NR3D_LAUNCH_CHECK("ray_marching(compact)");
    return 0;
}

int nr3d_march_pack(uint64_t n_rays, const int32_t* num_steps, int32_t* packed_info, int64_t* total, void* ws, uint64_t* ws_bytes, void* stream) {
    return scan_counts<int32_t, int32_t>(n_rays, num_steps, packed_info, total, ws, ws_bytes, (cudaStream_t)stream);
}

int nr3d_march_fill(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                    const int32_t* batch_inds, uint32_t batch_data_size, int32_t n_batches, const float* roi, const uint8_t* grid,
                    int32_t rx, int32_t ry, int32_t rz, int32_t contraction, float step_size, float max_step_size, float dt_gamma,
                    uint32_t max_steps, const int32_t* packed_info, float* t_starts, float* t_ends, int32_t* ridx, int32_t* bidx,
                    int32_t* gidx, void* stream) {
    if (n_rays == 0) return 0;
    MarchArgs a;
    if (int rc = fill_args(a, n_rays, rays_o, rays_d, t_min, t_max, batch_inds, batch_data_size, n_batches, roi, grid, rx, ry, rz,
                           contraction, step_size, max_step_size, dt_gamma, max_steps)) return rc;
    NR3D_CHECK(packed_info && t_starts && t_ends && ridx, "ray_marching: null output");
#ifdef NR3D_MARCH_FILL_UNSTAGED   // thread-per-ray stores (kept for A/B runs)
    march_kernel<1><<<(unsigned)div_up<uint64_t>(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(a, packed_info, nullptr, t_starts, t_ends, ridx, bidx, gidx, nullptr);
#else
    march_fill_staged_kernel<<<(unsigned)div_up<uint64_t>(n_rays, kFillThreads), kFillThreads, 0, (cudaStream_t)stream>>>(a, packed_info, t_starts, t_ends, ridx, bidx, gidx);
#endif
    NR3D_LAUNCH_CHECK("ray_marching(fill)");
    return 0;
}

int nr3d_forest_march_count(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                            const int32_t* seg_block_inds, const float* seg_entries, const float* seg_exits, const int32_t* seg_pack_infos,
                            const int16_t* block_ks, const float* world_origin, const float* world_block_size, const uint8_t* grid,
                            int32_t rx, int32_t ry, int32_t rz, float step_size, float max_step_size, float dt_gamma, uint32_t max_steps,
                            int32_t* num_steps, void* stream) {
    if (n_rays == 0) return 0;
    ForestMarchArgs a;
    if (int rc = fill_forest_args(a, n_rays, rays_o, rays_d, t_min, t_max, seg_block_inds, seg_entries, seg_exits, seg_pack_infos, block_ks,
                                  world_origin, world_block_size, grid, rx, ry, rz, step_size, max_step_size, dt_gamma, max_steps)) return rc;
    NR3D_CHECK(num_steps != nullptr, "forest_ray_marching: null num_steps");
    forest_march_kernel<false><<<(unsigned)div_up<uint64_t>(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(a, nullptr, num_steps, nullptr, nullptr, nullptr, nullptr, nullptr);
    NR3D_LAUNCH_CHECK("forest_ray_marching(count)");
    return 0;
}

int nr3d_forest_march_fill(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                           const int32_t* seg_block_inds, const float* seg_entries, const float* seg_exits, const int32_t* seg_pack_infos,
                           const int16_t* block_ks, const float* world_origin, const float* world_block_size, const uint8_t* grid,
                           int32_t rx, int32_t ry, int32_t rz, float step_size, float max_step_size, float dt_gamma, uint32_t max_steps,
                           const int32_t* packed_info, float* t_starts, float* t_ends, int32_t* ridx, int32_t* blidx, int32_t* gidx,
                           void* stream) {
    if (n_rays == 0) return 0;
    ForestMarchArgs a;
    if (int rc = fill_forest_args(a, n_rays, rays_o, rays_d, t_min, t_max, seg_block_inds, seg_entries, seg_exits, seg_pack_infos, block_ks,
                                  world_origin, world_block_size, grid, rx, ry, rz, step_size, max_step_size, dt_gamma, max_steps)) return rc;
    NR3D_CHECK(packed_info && t_starts && t_ends && ridx && blidx, "forest_ray_marching: null output");
    forest_march_kernel<true><<<(unsigned)div_up<uint64_t>(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(a, packed_info, nullptr, t_starts, t_ends, ridx, blidx, gidx);
    NR3D_LAUNCH_CHECK("forest_ray_marching(fill)");
    return 0;
}

}  // extern "C"
