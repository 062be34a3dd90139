// occ_update.cu -- occupancy value-grid maintenance for sm_100a.
//
// Behavioural contract: nr3d_lib/models/accelerations/occgrid/utils.py:18-133 (sample_pts_in_voxels, binarize,
// update_occ_val_grid[_idx]_, update_batched_occ_val_grid[_idx]_) and ema_single.py:186-205,215-218 (the EMA step and query).
// The reference composes ~15 torch passes plus torch_scatter.scatter_max (include-self max into `ema * grid`) and writes back
// only the touched cells:      grid[c] <- max(ema * grid[c], max{val_i : cell(i) = c})   for touched c,   untouched c unchanged.
// Here:   (1) one scatter kernel: point -> cell -> atomicMax of the order-preserving integer image of the value into a
//             zero-initialised uint32 scratch (0 = untouched; no float encodes to 0 except one NaN payload);
//         (2) one streaming pass over the cells: EMA-max of the touched ones, scratch reset to 0 for the next call, fused
//             threshold -> bool grid, optional double-precision sum for the mean-relative threshold.
// All float expressions use explicit _rn intrinsics in the reference's operation order (torch evaluates them as separate
// elementwise ops, so there is no FMA contraction to reproduce): cell indices are bit-exact.
#include "common.cuh"

namespace nr3d {

constexpr int kOccThreads = 256;

__device__ __forceinline__ uint32_t enc_ordered(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ((p / 2 + 0.5) * res).long().clamp(0, res - 1)       utils.py:109, ema_single.py:188,216
__device__ __forceinline__ int64_t cell_of(float p, int32_t res) {
    const float u = __fmul_rn(__fadd_rn(__fmul_rn(p, 0.5f), 0.5f), (float)res);
    int64_t c = (int64_t)u;                       // truncation toward zero, like Tensor.long()
    return c < 0 ? 0 : (c > (int64_t)res - 1 ? (int64_t)res - 1 : c);
}

struct Res3 { int32_t x, y, z; };

template <bool FROM_PTS>
__global__ void __launch_bounds__(kOccThreads)
occ_scatter_max_kernel(uint64_t N, const float* __restrict__ pts, const int64_t* __restrict__ gidx, const int64_t* __restrict__ bidx,
                       uint64_t batch_data_size, const float* __restrict__ vals, uint32_t B, Res3 res, uint32_t* __restrict__ scratch) {
    const uint64_t i = (uint64_t)blockIdx.x * kOccThreads + threadIdx.x;
    if (i >= N) return;
    int64_t cx, cy, cz;
    if (FROM_PTS) {
        cx = cell_of(pts[3 * i], res.x); cy = cell_of(pts[3 * i + 1], res.y); cz = cell_of(pts[3 * i + 2], res.z);
    } else {
        cx = gidx[3 * i]; cy = gidx[3 * i + 1]; cz = gidx[3 * i + 2];
        if (cx < 0 || cy < 0 || cz < 0 || cx >= res.x || cy >= res.y || cz >= res.z) return;   // reference: out-of-range index error
    }
    int64_t b = 0;
    if (bidx) b = bidx[i]; else if (batch_data_size) b = (int64_t)(i / batch_data_size);
    if (b < 0 || b >= (int64_t)B) return;
    const uint64_t cell = (((uint64_t)b * res.x + cx) * res.y + cy) * res.z + cz;
    atomicMax(scratch + cell, enc_ordered(vals[i]));
}

// streaming pass over all cells; VEC cells per thread iteration
__global__ void __launch_bounds__(kOccThreads)
occ_apply_kernel(uint64_t n, float* __restrict__ grid, uint32_t* __restrict__ scratch, float ema, int32_t write_occ, float thre,
                 uint8_t* __restrict__ occ, double* __restrict__ sum) {
    double local = 0.0;
    const uint64_t n4 = n / 4;
    const uint64_t stride = (uint64_t)gridDim.x * kOccThreads;
    for (uint64_t q = (uint64_t)blockIdx.x * kOccThreads + threadIdx.x; q < n4; q += stride) {
        float4 g = reinterpret_cast<float4*>(grid)[q];
        const uint4 s = reinterpret_cast<uint4*>(scratch)[q];
        if (s.x | s.y | s.z | s.w) {
            if (s.x) g.x = fmaxf(__fmul_rn(ema, g.x), dec_ordered(s.x));
            if (s.y) g.y = fmaxf(__fmul_rn(ema, g.y), dec_ordered(s.y));
            if (s.z) g.z = fmaxf(__fmul_rn(ema, g.z), dec_ordered(s.z));
            if (s.w) g.w = fmaxf(__fmul_rn(ema, g.w), dec_ordered(s.w));
            reinterpret_cast<float4*>(grid)[q] = g;
            reinterpret_cast<uint4*>(scratch)[q] = make_uint4(0, 0, 0, 0);
        }
        if (write_occ) reinterpret_cast<uchar4*>(occ)[q] = make_uchar4(g.x > thre, g.y > thre, g.z > thre, g.w > thre);
        if (sum) local += ((double)g.x + (double)g.y) + ((double)g.z + (double)g.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {    // tail cells
        const uint64_t c = n4 * 4 + threadIdx.x;
        float g = grid[c];
        const uint32_t s = scratch[c];
        if (s) { g = fmaxf(__fmul_rn(ema, g), dec_ordered(s)); grid[c] = g; scratch[c] = 0; }
        if (write_occ) occ[c] = g > thre;
        if (sum) local += (double)g;
    }
    if (sum) {
        __shared__ double part[kOccThreads / 32];
        for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kOccThreads / 32; ++w) t += part[w];
            atomicAdd(sum, t);
        }
    }
}

// binarize (utils.py:84-87): occ = val > thr,  thr = consider_mean ? min(mean - eps, occ_thre) : occ_thre
__global__ void __launch_bounds__(kOccThreads)
occ_binarize_kernel(uint64_t n, const float* __restrict__ grid, float thre, int32_t consider_mean, float eps, const double* __restrict__ sum,
                    uint8_t* __restrict__ occ) {
    float thr = thre;
    if (consider_mean) thr = fminf(__fsub_rn((float)(*sum / (double)n), eps), thre);
    const uint64_t stride = (uint64_t)gridDim.x * kOccThreads;
    for (uint64_t c = (uint64_t)blockIdx.x * kOccThreads + threadIdx.x; c < n; c += stride) occ[c] = grid[c] > thr;
}

// sample_pts_in_voxels (utils.py:18-39): pts = ((gidx[vidx] + offsets) / res) * 2 - 1 ; vidx given (random) or i / n_per_vox
__global__ void __launch_bounds__(kOccThreads)
occ_sample_kernel(uint64_t n_pts, const int64_t* __restrict__ gidx, const int64_t* __restrict__ vidx, uint64_t n_per_vox,
                  const float* __restrict__ offsets, Res3 res, float* __restrict__ pts, int64_t* __restrict__ vidx_out) {
    const uint64_t i = (uint64_t)blockIdx.x * kOccThreads + threadIdx.x;
    if (i >= n_pts) return;
    const int64_t v = vidx ? vidx[i] : (int64_t)(i / n_per_vox);
    const float r[3] = {(float)res.x, (float)res.y, (float)res.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float a = __fadd_rn((float)gidx[3 * v + d], offsets[3 * i + d]);
        pts[3 * i + d] = __fsub_rn(__fmul_rn(__fdiv_rn(a, r[d]), 2.0f), 1.0f);
    }
    if (vidx_out) vidx_out[i] = v;
}

// query (ema_single.py:214-218 / batched): out[i] = occ[b, cell(pts[i])]
__global__ void __launch_bounds__(kOccThreads)
occ_query_kernel(uint64_t N, const float* __restrict__ pts, const int64_t* __restrict__ bidx, uint64_t batch_data_size, uint32_t B, Res3 res,
                 const uint8_t* __restrict__ occ, uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * kOccThreads + threadIdx.x;
    if (i >= N) return;
    int64_t b = 0;
    if (bidx) b = bidx[i]; else if (batch_data_size) b = (int64_t)(i / batch_data_size);
    if (b < 0 || b >= (int64_t)B) { out[i] = 0; return; }
    const uint64_t cell = (((uint64_t)b * res.x + cell_of(pts[3 * i], res.x)) * res.y + cell_of(pts[3 * i + 1], res.y)) * res.z + cell_of(pts[3 * i + 2], res.z);
    out[i] = occ[cell];
}

static inline unsigned stream_grid(uint64_t work_items) {
    const uint64_t blocks = div_up<uint64_t>(work_items, kOccThreads);
    const uint64_t cap = (uint64_t)kSMs * 8;
    return (unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_occ_scatter_max(uint64_t N, const float* pts, const int64_t* gidx, const int64_t* bidx, uint64_t batch_data_size, const float* vals,
                         uint32_t B, const int32_t* res, uint32_t* scratch, void* stream) {
    if (N == 0) return 0;
    NR3D_CHECK((pts != nullptr) != (gidx != nullptr), "occ_scatter_max: give exactly one of `pts` / `gidx`");
    NR3D_CHECK(vals && res && scratch && B > 0, "occ_scatter_max: null argument");
    NR3D_CHECK(res[0] > 0 && res[1] > 0 && res[2] > 0, "occ_scatter_max: resolution must be positive, got [%d, %d, %d]", res[0], res[1], res[2]);
    const Res3 r{res[0], res[1], res[2]};
    const unsigned grid = (unsigned)div_up<uint64_t>(N, kOccThreads);
    if (pts) occ_scatter_max_kernel<true><<<grid, kOccThreads, 0, (cudaStream_t)stream>>>(N, pts, nullptr, bidx, batch_data_size, vals, B, r, scratch);
    else     occ_scatter_max_kernel<false><<<grid, kOccThreads, 0, (cudaStream_t)stream>>>(N, nullptr, gidx, bidx, batch_data_size, vals, B, r, scratch);
    NR3D_LAUNCH_CHECK("occ_scatter_max");
    return 0;
}

int nr3d_occ_apply(uint64_t n_cells, float* grid, uint32_t* scratch, float ema_decay, int32_t write_occ, float occ_thre, uint8_t* occ,
                   double* sum, void* stream) {
    if (n_cells == 0) return 0;
    NR3D_CHECK(grid && scratch, "occ_apply: null argument");
    NR3D_CHECK(!write_occ || occ, "occ_apply: `occ` output required when write_occ is set");
    NR3D_CHECK(((uintptr_t)grid & 15) == 0 && ((uintptr_t)scratch & 15) == 0 && (!occ || ((uintptr_t)occ & 3) == 0),
               "occ_apply: grid / scratch must be 16-byte aligned and occ 4-byte aligned");
    occ_apply_kernel<<<stream_grid(n_cells / 4 + 1), kOccThreads, 0, (cudaStream_t)stream>>>(n_cells, grid, scratch, ema_decay, write_occ, occ_thre, occ, sum);
    NR3D_LAUNCH_CHECK("occ_apply");
    return 0;
}

int nr3d_occ_binarize(uint64_t n_cells, const float* grid, float occ_thre, int32_t consider_mean, float eps, const double* sum, uint8_t* occ,
                      void* stream) {
    if (n_cells == 0) return 0;
    NR3D_CHECK(grid && occ, "occ_binarize: null argument");
    NR3D_CHECK(!consider_mean || sum, "occ_binarize: `sum` (from nr3d_occ_apply) required when consider_mean is set");
    occ_binarize_kernel<<<stream_grid(n_cells), kOccThreads, 0, (cudaStream_t)stream>>>(n_cells, grid, occ_thre, consider_mean, eps, sum, occ);
    NR3D_LAUNCH_CHECK("occ_binarize");
    return 0;
}

int nr3d_occ_sample_in_voxels(uint64_t n_pts, const int64_t* gidx, const int64_t* vidx, uint64_t n_per_vox, const float* offsets,
                              const int32_t* res, float* pts, int64_t* vidx_out, void* stream) {
    if (n_pts == 0) return 0;
    NR3D_CHECK(gidx && offsets && res && pts, "occ_sample_in_voxels: null argument");
    NR3D_CHECK(vidx || n_per_vox > 0, "occ_sample_in_voxels: give `vidx` or a positive `n_per_vox`");
    occ_sample_kernel<<<(unsigned)div_up<uint64_t>(n_pts, kOccThreads), kOccThreads, 0, (cudaStream_t)stream>>>(n_pts, gidx, vidx, n_per_vox, offsets,
        Res3{res[0], res[1], res[2]}, pts, vidx_out);
    NR3D_LAUNCH_CHECK("occ_sample_in_voxels");
    return 0;
}

int nr3d_occ_query(uint64_t N, const float* pts, const int64_t* bidx, uint64_t batch_data_size, uint32_t B, const int32_t* res, const uint8_t* occ,
                   uint8_t* out, void* stream) {
    if (N == 0) return 0;
    NR3D_CHECK(pts && res && occ && out && B > 0, "occ_query: null argument");
    occ_query_kernel<<<(unsigned)div_up<uint64_t>(N, kOccThreads), kOccThreads, 0, (cudaStream_t)stream>>>(N, pts, bidx, batch_data_size, B,
        Res3{res[0], res[1], res[2]}, occ, out);
    NR3D_LAUNCH_CHECK("occ_query");
    return 0;
}

}  // extern "C"
