/* lotd_port.c -- plain-C, float32, multi-threaded (OpenMP) CPU port of the reference's Dense/Hash ("hash-only") LoTD kernels.
 * TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/ (second oracle, op-order faithful) and by bench.py's cpu_baseline and
 * `--impl reference` legs.  Nothing under nr3d_lib_b200/ links or calls it.
 *
 * Restates, per (point, level), what one CUDA thread of the reference does:
 *   forward   kernel_lod_hash_only                  csrc/lotd/include/lotd/lotd_hash_only.h:15-162
 *   dL/dparam kernel_lod_hashonly_backward_grid     csrc/lotd/include/lotd/lotd_hash_only.h:380-470
 *   pos_fract (scale = res - 2, + 0.5, floor)       csrc/lotd/include/lotd/lotd_cuda.h:959-1077
 *   indices   grid_index_dense (last dim fastest) / grid_index_hash (primes 1, 2654435761, 805459861)   lotd_cuda.h:92-160
 *   n-linear  corner idx bit d <-> dimension d, weight multiplied in dimension order   linear_interpolate.cuh:92-121
 * The reference has no CPU implementation of this path (every op checks for CUDA tensors), hence "port".
 * Pinned by tests/test_oracle_cpu.py against the golden vectors recorded from the reference's own CUDA build
 * (tests/golden/lotd_ngp8_f32.npz, lotd_ngp_smooth_f32.npz, lotd_hash_f4_f32.npz) and against the float64 oracle.
 *
 * Layout: x [N,3] f32, params flat f32 (level tables at `offsets`), y / dL_dy row-major [N, n_enc], one level = n_feats[l]
 * consecutive output features.  Cubic levels only (res[l] per axis).  Gradients are accumulated with `omp atomic`.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { PORT_DENSE = 0, PORT_HASH = 7 };   /* values of the reference's LoDType enum (lotd_types.h:16-26) */

static inline uint32_t index_of(uint32_t type, uint32_t R, uint32_t size, uint32_t px, uint32_t py, uint32_t pz) {
    if (type == PORT_DENSE) return (px * R + py) * R + pz;                                  /* uint32 arithmetic as in the reference */
    const uint32_t h = (px * 1u) ^ (py * 2654435761u) ^ (pz * 805459861u);
    return h % size;
}

static inline void pos_fract(float x, float scale, int smooth, uint32_t* cell, float* p) {
    float v = fmaf(x, scale, 0.5f);                                                         /* nvcc contracts x * scale + 0.5f */
    const float fl = floorf(v);
    *cell = (uint32_t)fl;
    v -= fl;
    *p = smooth ? v * v * (3.0f - 2.0f * v) : v;
}

int lotd_port_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void lotd_port_fwd(int n_levels, const uint32_t* res, const uint32_t* types, const uint32_t* n_feats, const uint32_t* sizes,
                   const uint32_t* offsets, int smooth, uint64_t N, const float* x, const float* params, float* y, int n_threads) {
    uint32_t n_enc = 0;
    for (int l = 0; l < n_levels; ++l) n_enc += n_feats[l];
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
#endif
    for (int64_t i = 0; i < (int64_t)N; ++i) {
        const float* xp = x + 3 * (size_t)i;
        float* yo = y + (size_t)i * n_enc;
        for (int l = 0; l < n_levels; ++l) {
            const uint32_t R = res[l], F = n_feats[l];
            const float scale = (float)(R - 2u);
            uint32_t c[3];
            float p[3];
            for (int d = 0; d < 3; ++d) pos_fract(xp[d], scale, smooth, &c[d], &p[d]);
            const float* tbl = params + offsets[l];
            for (uint32_t f = 0; f < F; ++f) yo[f] = 0.0f;
            for (int idx = 0; idx < 8; ++idx) {
                float w = 1.0f;
                uint32_t pos[3];
                for (int d = 0; d < 3; ++d) {
                    if (idx & (1 << d)) { w *= p[d]; pos[d] = c[d] + 1u; }
                    else { w *= 1.0f - p[d]; pos[d] = c[d]; }
                }
                const float* e = tbl + (size_t)index_of(types[l], R, sizes[l], pos[0], pos[1], pos[2]) * F;
                for (uint32_t f = 0; f < F; ++f) yo[f] += w * e[f];
            }
            yo += F;
        }
    }
}

void lotd_port_bwd(int n_levels, const uint32_t* res, const uint32_t* types, const uint32_t* n_feats, const uint32_t* sizes,
                   const uint32_t* offsets, int smooth, uint64_t N, const float* x, const float* dL_dy, float* grad, int n_threads) {
    uint32_t n_enc = 0;
    for (int l = 0; l < n_levels; ++l) n_enc += n_feats[l];
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
#endif
    for (int64_t i = 0; i < (int64_t)N; ++i) {
        const float* xp = x + 3 * (size_t)i;
        const float* go = dL_dy + (size_t)i * n_enc;
        for (int l = 0; l < n_levels; ++l) {
            const uint32_t R = res[l], F = n_feats[l];
            const float scale = (float)(R - 2u);
            uint32_t c[3];
            float p[3];
            for (int d = 0; d < 3; ++d) pos_fract(xp[d], scale, smooth, &c[d], &p[d]);
            float* tbl = grad + offsets[l];
            for (int idx = 0; idx < 8; ++idx) {
                float w = 1.0f;
                uint32_t pos[3];
                for (int d = 0; d < 3; ++d) {
                    if (idx & (1 << d)) { w *= p[d]; pos[d] = c[d] + 1u; }
                    else { w *= 1.0f - p[d]; pos[d] = c[d]; }
                }
                float* e = tbl + (size_t)index_of(types[l], R, sizes[l], pos[0], pos[1], pos[2]) * F;
                for (uint32_t f = 0; f < F; ++f) {
                    const float v = w * go[f];
#ifdef _OPENMP
#pragma omp atomic
#endif
                    e[f] += v;
                }
            }
            go += F;
        }
    }
}
