/* march_oracle.c -- CPU oracle of the occupancy-grid ray marcher.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, float32, operation-order-exact restatement of
 *   csrc/occ_grid/src/ray_marching.cu:17-134 and batched_marching.cu:18-148   (the per-ray loop)
 *   csrc/occ_grid/include/occ_grid/helpers_march.h:11-77                      (calc_dt, grid lookup, DDA skip)
 *   csrc/occ_grid/include/occ_grid/helpers_contraction.h:10-125               (roi / tanh / sphere contraction)
 *   csrc/occ_grid/include/occ_grid/helpers_math.h:1167-1172,1471-1475          (clamp, sign = copysignf)
 * The only place where nvcc's default a*b+c contraction changes a result is `origin + t_mid * dir`
 * (ray_marching.cu:88); it is written with fmaf() here.  Every other candidate multiplies by 0.5f (exact), so the
 * fused and unfused results coincide.  Compile with -ffp-contract=off so that gcc adds no contractions of its own.
 *
 * Pinned against the reference's own CUDA build on a B200 (tests/golden/march_*.npz, tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

typedef struct { float x, y, z; } f3;

static float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
static float calc_dt(float t, float dt_gamma, float dt_min, float dt_max) { return clampf(t * dt_gamma, dt_min, dt_max); }

static f3 roi_to_unit(f3 p, f3 lo, f3 hi) {
    f3 r = {(p.x - lo.x) / (hi.x - lo.x), (p.y - lo.y) / (hi.y - lo.y), (p.z - lo.z) / (hi.z - lo.z)};
    return r;
}

static f3 apply_contraction(f3 p, f3 lo, f3 hi, int type) {
    f3 u = roi_to_unit(p, lo, hi);
    if (type == 1) { /* UN_BOUNDED_TANH, helpers_contraction.h:24-35 */
        f3 r = {fmaf(tanhf(u.x - 0.5f), 0.5f, 0.5f), fmaf(tanhf(u.y - 0.5f), 0.5f, 0.5f), fmaf(tanhf(u.z - 0.5f), 0.5f, 0.5f)};
        return r;
    }
    if (type == 2) { /* UN_BOUNDED_SPHERE, helpers_contraction.h:58-76 */
        f3 v = {fmaf(u.x, 2.0f, -1.0f), fmaf(u.y, 2.0f, -1.0f), fmaf(u.z, 2.0f, -1.0f)};
        float norm_sq = fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x));
        float norm = sqrtf(norm_sq);
        if (norm > 1.0f) {
            float s = 2.0f - 1.0f / norm;
            v.x = s * (v.x / norm); v.y = s * (v.y / norm); v.z = s * (v.z / norm);
        }
        f3 r = {fmaf(v.x, 0.25f, 0.5f), fmaf(v.y, 0.25f, 0.5f), fmaf(v.z, 0.25f, 0.5f)};
        return r;
    }
    return u;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

static int grid_occupied_at(f3 p, f3 lo, f3 hi, int type, const int res[3], const uint8_t* grid, int* idx_out) {
    if (type == 0 && (p.x < lo.x || p.x > hi.x || p.y < lo.y || p.y > hi.y || p.z < lo.z || p.z > hi.z)) return 0;
    f3 u = apply_contraction(p, lo, hi, type);
    int ix = (int)(u.x * (float)res[0]), iy = (int)(u.y * (float)res[1]), iz = (int)(u.z * (float)res[2]);
    ix = imax(0, imin(ix, res[0] - 1));
    iy = imax(0, imin(iy, res[1] - 1));
    iz = imax(0, imin(iz, res[2] - 1));
    int idx = ix * (res[1] * res[2]) + iy * res[2] + iz;
    *idx_out = idx;
    return grid[idx] != 0;
}

static float dist_axis(float q, float d, float inv_d, float r, float extent) {
    return ((floorf(q + 0.5f + 0.5f * copysignf(1.0f, d)) - q) * inv_d) / r * extent;
}

static float advance_to_next_voxel(float t, float dt_min, f3 p, f3 dir, f3 inv_dir, f3 lo, f3 hi, const int res[3]) {
    f3 u = roi_to_unit(p, lo, hi);
    float rx = (float)res[0], ry = (float)res[1], rz = (float)res[2];
    float tx = dist_axis(u.x * rx, dir.x, inv_dir.x, rx, hi.x - lo.x);
    float ty = dist_axis(u.y * ry, dir.y, inv_dir.y, ry, hi.y - lo.y);
    float tz = dist_axis(u.z * rz, dir.z, inv_dir.z, rz, hi.z - lo.z);
    float t_target = t + fmaxf(fminf(fminf(tx, ty), tz), 0.0f);
    float t_ = t;
    do { t_ += dt_min; } while (t_ < t_target);
    return t_;
}

/* One ray.  If t_starts == NULL only counts.  Returns the number of samples written/counted. */
static uint32_t march_one(const float* o, const float* d, float near, float far, const float* roi, const int res[3],
                          const uint8_t* grid, int type, float step_size, float max_step_size, float dt_gamma, uint32_t max_steps,
                          float* t_starts, float* t_ends, int32_t* gidx, int32_t grid_offset) {
    f3 origin = {o[0], o[1], o[2]}, dir = {d[0], d[1], d[2]};
    f3 inv_dir = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
    f3 lo = {roi[0], roi[1], roi[2]}, hi = {roi[3], roi[4], roi[5]};
    float dt_min = step_size, dt_max = max_step_size;
    uint32_t j = 0;
    float t0 = near;
    float dt = calc_dt(t0, dt_gamma, dt_min, dt_max);
    float t1 = t0 + dt;
    float t_mid = (t0 + t1) * 0.5f;
    while ((t_mid < far) && (j < max_steps)) {
        f3 p = {fmaf(t_mid, dir.x, origin.x), fmaf(t_mid, dir.y, origin.y), fmaf(t_mid, dir.z, origin.z)};
        int gi = -1;
        if (grid_occupied_at(p, lo, hi, type, res, grid, &gi)) {
            if (t_starts) { t_starts[j] = t0; t_ends[j] = t1; if (gidx) gidx[j] = gi + grid_offset; }
            ++j;
            t0 = t1;
            t1 = t0 + calc_dt(t0, dt_gamma, dt_min, dt_max);
            t_mid = (t0 + t1) * 0.5f;
        } else if (type == 0) {
            t_mid = advance_to_next_voxel(t_mid, dt_min, p, dir, inv_dir, lo, hi, res);
            dt = calc_dt(t_mid, dt_gamma, dt_min, dt_max);
            t0 = t_mid - dt * 0.5f;
            t1 = t_mid + dt * 0.5f;
        } else {
            t0 = t1;
            t1 = t0 + calc_dt(t0, dt_gamma, dt_min, dt_max);
            t_mid = (t0 + t1) * 0.5f;
        }
    }
    return j;
}

/* Pass 1: num_steps[R]. batch_inds may be NULL; batch_data_size 0 = unused. Rays with batch_inds<0 count 0. */
void march_oracle_count(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                        const int32_t* batch_inds, uint32_t batch_data_size, const float* roi, const uint8_t* grid,
                        int rx, int ry, int rz, int type, float step_size, float max_step_size, float dt_gamma,
                        uint32_t max_steps, int32_t* num_steps) {
    const int res[3] = {rx, ry, rz};
    const uint32_t cells = (uint32_t)(rx * ry * rz);
    for (uint64_t i = 0; i < n_rays; ++i) {
        uint32_t b = 0;
        if (batch_inds) { if (batch_inds[i] < 0) { num_steps[i] = 0; continue; } b = (uint32_t)batch_inds[i]; }
        else if (batch_data_size) b = (uint32_t)(i / batch_data_size);
        num_steps[i] = (int32_t)march_one(rays_o + 3 * i, rays_d + 3 * i, t_min[i], t_max[i], roi + 6 * (size_t)b, res,
                                          grid + (size_t)b * cells, type, step_size, max_step_size, dt_gamma, max_steps,
                                          NULL, NULL, NULL, 0);
    }
}

/* Pass 2 with packed_info[R,2] = (offset, count) as produced from pass 1. */
void march_oracle_fill(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                       const int32_t* batch_inds, uint32_t batch_data_size, const float* roi, const uint8_t* grid,
                       int rx, int ry, int rz, int type, float step_size, float max_step_size, float dt_gamma,
                       const int32_t* packed_info, float* t_starts, float* t_ends, int32_t* ridx, int32_t* bidx, int32_t* gidx) {
    const int res[3] = {rx, ry, rz};
    const uint32_t cells = (uint32_t)(rx * ry * rz);
    for (uint64_t i = 0; i < n_rays; ++i) {
        uint32_t b = 0;
        if (batch_inds) { if (batch_inds[i] < 0) continue; b = (uint32_t)batch_inds[i]; }
        else if (batch_data_size) b = (uint32_t)(i / batch_data_size);
        const int32_t base = packed_info[2 * i], cnt = packed_info[2 * i + 1];
        if (cnt <= 0) continue;
        uint32_t n = march_one(rays_o + 3 * i, rays_d + 3 * i, t_min[i], t_max[i], roi + 6 * (size_t)b, res,
                               grid + (size_t)b * cells, type, step_size, max_step_size, dt_gamma, (uint32_t)cnt,
                               t_starts + base, t_ends + base, gidx ? gidx + base : NULL, (int32_t)(b * cells));
        for (uint32_t j = 0; j < n; ++j) { ridx[base + j] = (int32_t)i; if (bidx) bidx[base + j] = (int32_t)b; }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * forest (multi-block) marcher: csrc/occ_grid/src/forest_marching.cu:27-150.  One running (j, t0, t1, t_mid) per ray
 * across its block segments; block b spans [world_origin + k_b * world_block_size, + world_block_size] (the product is
 * contracted into an FMA by nvcc, hence fmaf) and is looked up WITHOUT the AABB test, with a clamped voxel index
 * (forest_marching.cu:16-25).  If t_starts == NULL only counts.
 * ------------------------------------------------------------------------------------------------------------- */
static uint32_t forest_march_one(const float* o, const float* d, float near, float far, uint32_t seg_length, const int32_t* seg_block_inds,
                                 const float* seg_entries, const float* seg_exits, const int16_t* block_ks, const float* world_origin,
                                 const float* world_block_size, const int res[3], const uint8_t* grid_all, float step_size,
                                 float max_step_size, float dt_gamma, uint32_t max_steps, float* t_starts, float* t_ends, int32_t* blidx,
                                 int32_t* gidx) {
    f3 origin = {o[0], o[1], o[2]}, dir = {d[0], d[1], d[2]};
    f3 inv_dir = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
    const uint32_t cells = (uint32_t)(res[0] * res[1] * res[2]);
    float dt_min = step_size, dt_max = max_step_size;
    uint32_t j = 0;
    float t0 = near;
    float dt = calc_dt(t0, dt_gamma, dt_min, dt_max);
    float t1 = t0 + dt;
    float t_mid = (t0 + t1) * 0.5f;
    for (uint32_t s = 0; s < seg_length; ++s) {
        const float cur_entry = seg_entries[s], cur_exit = seg_exits[s];
        const uint32_t b = (uint32_t)seg_block_inds[s];
        const int16_t* k = block_ks + 3 * (size_t)b;
        f3 lo = {fmaf((float)k[0], world_block_size[0], world_origin[0]), fmaf((float)k[1], world_block_size[1], world_origin[1]),
                 fmaf((float)k[2], world_block_size[2], world_origin[2])};
        f3 hi = {lo.x + world_block_size[0], lo.y + world_block_size[1], lo.z + world_block_size[2]};
        const uint32_t grid_offset = b * cells;
        const uint8_t* grid = grid_all + grid_offset;
        if (cur_entry >= far || cur_exit <= near) break;
        do { t_mid += step_size; } while (t_mid < cur_entry);
        dt = calc_dt(t_mid, dt_gamma, dt_min, dt_max);
        t0 = t_mid - dt * 0.5f;
        t1 = t_mid + dt * 0.5f;
        while ((t_mid <= cur_exit) && (t_mid <= far) && (j < max_steps)) {
            f3 p = {fmaf(t_mid, dir.x, origin.x), fmaf(t_mid, dir.y, origin.y), fmaf(t_mid, dir.z, origin.z)};
            f3 u = roi_to_unit(p, lo, hi);
            int ix = (int)(u.x * (float)res[0]), iy = (int)(u.y * (float)res[1]), iz = (int)(u.z * (float)res[2]);
            ix = imax(0, imin(ix, res[0] - 1));
            iy = imax(0, imin(iy, res[1] - 1));
            iz = imax(0, imin(iz, res[2] - 1));
            const int gi = ix * (res[1] * res[2]) + iy * res[2] + iz;
            if (grid[gi] != 0) {
                if (t_starts) { t_starts[j] = t0; t_ends[j] = t1; blidx[j] = (int32_t)b; if (gidx) gidx[j] = gi + (int32_t)grid_offset; }
                ++j;
                t0 = t1;
                t1 = t0 + calc_dt(t0, dt_gamma, dt_min, dt_max);
                t_mid = (t0 + t1) * 0.5f;
            } else {
                t_mid = advance_to_next_voxel(t_mid, dt_min, p, dir, inv_dir, lo, hi, res);
                dt = calc_dt(t_mid, dt_gamma, dt_min, dt_max);
                t0 = t_mid - dt * 0.5f;
                t1 = t_mid + dt * 0.5f;
            }
        }
    }
    return j;
}

void forest_march_oracle_count(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                               const int32_t* seg_block_inds, const float* seg_entries, const float* seg_exits, const int32_t* seg_pack_infos,
                               const int16_t* block_ks, const float* world_origin, const float* world_block_size, const uint8_t* grid,
                               int rx, int ry, int rz, float step_size, float max_step_size, float dt_gamma, uint32_t max_steps,
                               int32_t* num_steps) {
    const int res[3] = {rx, ry, rz};
    for (uint64_t i = 0; i < n_rays; ++i) {
        const int32_t sb = seg_pack_infos[2 * i], sl = seg_pack_infos[2 * i + 1];
        num_steps[i] = (int32_t)forest_march_one(rays_o + 3 * i, rays_d + 3 * i, t_min[i], t_max[i], (uint32_t)sl, seg_block_inds + sb,
                                                 seg_entries + sb, seg_exits + sb, block_ks, world_origin, world_block_size, res, grid,
                                                 step_size, max_step_size, dt_gamma, max_steps, NULL, NULL, NULL, NULL);
    }
}

void forest_march_oracle_fill(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                              const int32_t* seg_block_inds, const float* seg_entries, const float* seg_exits, const int32_t* seg_pack_infos,
                              const int16_t* block_ks, const float* world_origin, const float* world_block_size, const uint8_t* grid,
                              int rx, int ry, int rz, float step_size, float max_step_size, float dt_gamma, const int32_t* packed_info,
                              float* t_starts, float* t_ends, int32_t* ridx, int32_t* blidx, int32_t* gidx) {
    const int res[3] = {rx, ry, rz};
    for (uint64_t i = 0; i < n_rays; ++i) {
        const int32_t sb = seg_pack_infos[2 * i], sl = seg_pack_infos[2 * i + 1];
        const int32_t base = packed_info[2 * i], cnt = packed_info[2 * i + 1];
        if (cnt <= 0) continue;
        uint32_t n = forest_march_one(rays_o + 3 * i, rays_d + 3 * i, t_min[i], t_max[i], (uint32_t)sl, seg_block_inds + sb, seg_entries + sb,
                                      seg_exits + sb, block_ks, world_origin, world_block_size, res, grid, step_size, max_step_size, dt_gamma,
                                      (uint32_t)cnt, t_starts + base, t_ends + base, blidx + base, gidx ? gidx + base : NULL);
        for (uint32_t j = 0; j < n; ++j) ridx[base + j] = (int32_t)i;
    }
}
